"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the compiled reference (oracle/_ref/libnexus_ref.so).

TEST INFRASTRUCTURE: importable only from tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke().  The product package nexus_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libnexus_ref.so")

_oracle = None
_ref = None


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.orc_bvh2_cost.restype = C.c_double
        _oracle.orc_bvh8_cost.restype = C.c_double
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """The unmodified reference kernels + our headless driver (needs a GPU to *run*; loading works anywhere)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


# ------------------------------------------------------------------ oracle: builder ----
def prim_bounds(prims, prim_type):
    prims = np.ascontiguousarray(prims, dtype=np.float32)
    n = prims.shape[0]
    bounds = np.empty((n, 6), np.float32)
    scene = np.empty(6, np.float32)
    oracle().orc_prim_bounds(_p(prims), C.c_uint32(n), C.c_int(prim_type), _p(bounds), _p(scene))
    return bounds, scene


def morton(bounds, scene, bits64):
    n = bounds.shape[0]
    out = np.empty(n, np.uint64)
    oracle().orc_morton(_p(bounds), C.c_uint32(n), _p(scene), C.c_int(int(bits64)), _p(out))
    return out


def build_bvh2(bounds, codes, bits64):
    n = bounds.shape[0]
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    rc = oracle().orc_build_bvh2(_p(np.ascontiguousarray(bounds, np.float32)), _p(np.ascontiguousarray(codes, np.uint64)),
                                 C.c_uint32(n), C.c_int(int(bits64)), _p(nodes))
    assert rc == 0, rc
    return nodes


def build_bvh8(bvh2_nodes, n):
    cap = (4 * n - 1 + 6) // 7
    nodes = np.zeros((cap, 20), np.uint32)
    prim_idx = np.zeros(n, np.uint32)
    cnt = C.c_uint32(0)
    rc = oracle().orc_build_bvh8(_p(np.ascontiguousarray(bvh2_nodes)), C.c_uint32(n), _p(nodes), _p(prim_idx), C.byref(cnt))
    assert rc == 0, rc
    return nodes[:cnt.value].copy(), prim_idx


def canon_bvh8(nodes, prim_idx):
    nodes = np.ascontiguousarray(nodes).view(np.uint32).reshape(-1, 20)
    prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
    out = np.zeros_like(nodes)
    pout = np.zeros_like(prim_idx)
    rc = oracle().orc_canon_bvh8(_p(nodes), _p(prim_idx), C.c_uint32(nodes.shape[0]), C.c_uint32(prim_idx.shape[0]), _p(out), _p(pout))
    assert rc == 0, f"canon_bvh8 failed: {rc}"
    return out, pout


def canon_bvh2(nodes, n):
    nodes = np.ascontiguousarray(nodes).view(np.uint32).reshape(-1, 8)
    out = np.zeros_like(nodes)
    rc = oracle().orc_canon_bvh2(_p(nodes), C.c_uint32(n), _p(out))
    assert rc == 0, f"canon_bvh2 failed: {rc}"
    return out


def bvh2_cost(nodes, scene):
    nodes = np.ascontiguousarray(nodes)
    return oracle().orc_bvh2_cost(_p(nodes), C.c_uint32(nodes.shape[0]), _p(np.ascontiguousarray(scene, np.float32)))


def bvh8_cost(nodes, scene):
    nodes = np.ascontiguousarray(nodes)
    return oracle().orc_bvh8_cost(_p(nodes), C.c_uint32(nodes.shape[0]), _p(np.ascontiguousarray(scene, np.float32)))


def check_bvh8(nodes, prim_idx, bounds):
    nodes = np.ascontiguousarray(nodes)
    return oracle().orc_check_bvh8(_p(nodes), _p(np.ascontiguousarray(prim_idx, np.uint32)), C.c_uint32(nodes.shape[0]),
                                   C.c_uint32(prim_idx.shape[0]), _p(np.ascontiguousarray(bounds, np.float32)))
