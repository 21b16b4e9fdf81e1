"""GPU parity of the H-PLOC builder + BVH8 collapse through the C ABI.

Bar (north_star): BVH topology and node ordering bit-exact with the reference after canonical renumbering (both builders
number nodes in GPU-schedule order; tests/oracle_lib.canon_* renumbers breadth-first in slot order).  Checked against
 (1) the committed golden trees produced by the unmodified reference kernels (tests/golden/builder_ref.npz),
 (2) the CPU oracle on the same inputs, and (3) the reference itself when oracle/_ref is present on the box.
"""
import os

import numpy as np
import pytest

import nexus_b200 as nx
import oracle_lib as O
from golden_cases import builder_cases
from nexus_b200 import scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "builder_ref.npz")


def _build(ctx, prims, speed):
    n = prims.shape[0]
    b2 = nx.BuildBVH2(ctx, prims, prioritizeSpeed=speed)
    n2 = b2.ToHost(); b2.Free()
    b8 = nx.BuildBVH8(ctx, prims, prioritizeSpeed=speed)
    n8, p8 = b8.ToHost(); bounds = b8.bounds; b8.Free()
    return n, n2, n8, p8, bounds


@pytest.mark.parametrize("case", builder_cases(), ids=lambda c: c[0])
def test_trees_equal_reference_golden(ctx, case):
    name, prims, speed = case
    gold = np.load(GOLD)
    n, n2, n8, p8, bounds = _build(ctx, prims, speed)
    c8, cp8 = O.canon_bvh8(n8, p8)
    assert c8.shape == gold[name + "/bvh8"].shape
    assert (c8 == gold[name + "/bvh8"]).all(), "BVH8 nodes differ from the reference"
    assert (cp8 == gold[name + "/prim_idx"]).all()
    if n > 1:
        assert (O.canon_bvh2(n2, n) == gold[name + "/bvh2"]).all(), "BVH2 differs from the reference"
    assert (bounds == gold[name + "/bounds"]).all()


@pytest.mark.parametrize("case", builder_cases(), ids=lambda c: c[0])
def test_trees_equal_cpu_oracle(ctx, case):
    name, prims, speed = case
    n, n2, n8, p8, _ = _build(ctx, prims, speed)
    tri = 1 if prims.shape[1] == 9 else 0
    bits64 = 0 if speed else 1
    pb, sceneb = O.prim_bounds(prims, tri)
    codes = nx.debug_morton(ctx, prims, bits64)          # the device's fast-math Morton keys (div.approx), fed to the oracle
    o2 = O.build_bvh2(pb, codes, bits64)
    if n > 1:
        assert (O.canon_bvh2(n2, n) == O.canon_bvh2(o2, n)).all()
    o8, op8 = O.build_bvh8(o2, n)
    c8, cp8 = O.canon_bvh8(n8, p8)
    oc8, ocp8 = O.canon_bvh8(o8, op8)
    assert c8.shape == oc8.shape and (c8 == oc8).all() and (cp8 == ocp8).all()
    assert O.check_bvh8(n8, p8, pb) == 0                  # every primitive once, child boxes contain their subtrees


def test_structural_invariants_and_metrics(ctx):
    rng = np.random.default_rng(5)
    n = 200_000
    c = rng.uniform(-20, 20, (n, 1, 3)).astype(np.float32)
    prims = (c + rng.uniform(-0.05, 0.05, (n, 3, 3)).astype(np.float32)).reshape(n, 9)
    b2 = nx.BuildBVH2(ctx, prims, prioritizeSpeed=True)
    n2 = b2.ToHost(); b2.Free()
    assert n2.shape[0] == 2 * n - 1
    leaf = n2[:, 6] == 0xffffffff
    assert leaf[:n].all() and not leaf[n:].any()                       # leaves [0,n), inner [n,2n-1), root = 2n-2
    assert (n2[:n, 7] == np.arange(n)).all()
    kids = np.concatenate([n2[n:, 6], n2[n:, 7]])
    assert len(np.unique(kids)) == 2 * n - 2 and (2 * n - 2) not in kids  # every node but the root has exactly one parent
    b8, m = nx.BuildBVH8(ctx, prims, prioritizeSpeed=True, metrics=True)
    n8, p8 = b8.ToHost(); b8.Free()
    assert len(n8) <= (4 * n - 1 + 6) // 7
    assert (np.sort(p8) == np.arange(n)).all()
    pb, sceneb = O.prim_bounds(prims, 1)
    assert O.check_bvh8(n8, p8, pb) == 0
    # SAH cost as Eval.cu defines it: device sum vs the oracle's evaluation of the same tree (summation order differs)
    assert abs(m["bvh8_cost"] - O.bvh8_cost(n8, sceneb)) <= 1e-4 * m["bvh8_cost"]
    assert abs(m["bvh2_cost"] - O.bvh2_cost(n2, sceneb)) <= 1e-4 * m["bvh2_cost"]
    assert m["total_ms"] > 0 and 2.0 < m["avg_children_per_node"] <= 8.0


@pytest.mark.parametrize("speed", [True, False])
def test_matches_live_reference_large(ctx, have_ref, speed):
    """When the compiled reference travelled to the box: a 500k-triangle build, both key widths, tree for tree."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from nexus_b200 import scenes
    prims = scenes.test_triangles(500_000)
    _, n2, n8, p8, _ = _build(ctx, prims, speed)
    r8, rp8, _ = O.ref_build_bvh8(prims, speed)
    c8, cp8 = O.canon_bvh8(n8, p8)
    rc8, rcp8 = O.canon_bvh8(r8, rp8)
    assert c8.shape == rc8.shape and (c8 == rc8).all() and (cp8 == rcp8).all()
    r2, _ = O.ref_build_bvh2(prims, speed)
    assert (O.canon_bvh2(n2, len(prims)) == O.canon_bvh2(r2, len(prims))).all()


def test_rebuild_is_deterministic(ctx):
    """Node numbering is ours (level order), and the tree does not depend on the schedule: two builds, same bytes after canon."""
    from nexus_b200 import scenes
    prims = scenes.test_triangles(100_000, seed=9)
    a = _build(ctx, prims, True)
    b = _build(ctx, prims, True)
    assert (O.canon_bvh8(a[2], a[3])[0] == O.canon_bvh8(b[2], b[3])[0]).all()


def test_invalid_inputs_fail_loudly(ctx):
    with pytest.raises(nx.NexusError):
        nx.BuildBVH8(ctx, np.zeros((0, 9), np.float32))


def test_bench_size_build_properties(ctx, have_ref):
    """At bench size (10 M triangles of the NexusBVH benchmark mesh, BASELINE.json configs[3]'s generator) parity is checked
    through size-independent properties: every primitive appears exactly once, every child box contains its subtree, the
    node count respects the ceil((4n-1)/7) bound, rebuilding gives the same tree, and - when the compiled reference travelled
    to the box - node count and both SAH costs (Eval.cu definitions) equal the reference's (1e-4: float-atomic summation order)."""
    n = 10_000_000
    prims = scenes.test_triangles(n)
    b8, m = nx.BuildBVH8(ctx, prims, prioritizeSpeed=True, metrics=True)
    n8, p8 = b8.ToHost(); b8.Free()
    assert len(n8) <= (4 * n - 1 + 6) // 7
    assert (np.bincount(p8, minlength=n) == 1).all()
    pb, _ = O.prim_bounds(prims, 1)
    assert O.check_bvh8(n8, p8, pb) == 0
    b8b = nx.BuildBVH8(ctx, prims, prioritizeSpeed=True)
    n8b, p8b = b8b.ToHost(); b8b.Free()
    ca, cb = O.canon_bvh8(n8, p8), O.canon_bvh8(n8b, p8b)     # numbering inside a level follows the atomics; the tree does not
    assert (ca[0] == cb[0]).all() and (ca[1] == cb[1]).all()
    if have_ref:
        import ctypes as C
        mm = np.zeros(9, np.float32); cnt = C.c_uint32(0)
        assert O.ref().nxref_benchmark_bvh8(prims.ctypes.data_as(C.c_void_p), C.c_uint32(n), 1, 1, 0, 1, mm.ctypes.data_as(C.c_void_p), C.byref(cnt), None) == 0
        assert cnt.value == len(n8)
        assert abs(m["bvh2_cost"] - mm[6]) <= 1e-4 * mm[6] and abs(m["bvh8_cost"] - mm[7]) <= 1e-4 * mm[7]


@pytest.mark.parametrize("pmax", [1, 2, 3])
@pytest.mark.parametrize("case", builder_cases(), ids=lambda c: c[0])
def test_sah_optimal_collapse_equals_cpu_restatement(ctx, case, pmax):
    """nx_build_config::collapse = NX_COLLAPSE_SAH_OPTIMAL (SURVEY.md §8 row a20 on the GPU): the C(n, i) table and child
    selection of the reference's CPU BVH8Builder, run bottom-up on the device, against the CPU restatement
    (oracle_bvh.cpp: orc_build_bvh8_optimal) applied to the device's own BVH2: canonical CWBVH8 and leaf order bit for bit."""
    name, prims, speed = case
    n = prims.shape[0]
    tri = 1 if prims.shape[1] == 9 else 0
    b2 = nx.BuildBVH2(ctx, prims, prioritizeSpeed=speed)
    n2 = b2.ToHost(); b2.Free()
    b8 = nx.BuildBVH8(ctx, prims, prioritizeSpeed=speed, collapse=nx.COLLAPSE_SAH_OPTIMAL, maxLeafPrims=pmax)
    n8, p8 = b8.ToHost(); b8.Free()
    pb, _ = O.prim_bounds(prims, tri)
    assert O.check_bvh8(n8, p8, pb) == 0
    o8, op8 = O.build_bvh8_optimal(n2, n, pmax)
    c8, cp8 = O.canon_bvh8(n8, p8)
    oc8, ocp8 = O.canon_bvh8(o8, op8)
    assert c8.shape == oc8.shape and (c8 == oc8).all() and (cp8 == ocp8).all()
    meta = n8.view(np.uint8).reshape(-1, 80)[:, 24:32]
    leaf = (meta != 0) & ((meta & 0x1f) < 24)
    assert np.isin(meta[leaf] >> 5, [1, 3, 7][:pmax]).all()          # unary primitive counts 1..pmax


def test_sah_optimal_collapse_large_and_fewer_nodes(ctx):
    """500k triangles of the NexusBVH benchmark mesh: same equality at size."""
    prims = scenes.test_triangles(500_000)
    n = len(prims)
    b2 = nx.BuildBVH2(ctx, prims, prioritizeSpeed=True)
    n2 = b2.ToHost(); b2.Free()
    for pmax in (1, 3):
        b8 = nx.BuildBVH8(ctx, prims, prioritizeSpeed=True, collapse=nx.COLLAPSE_SAH_OPTIMAL, maxLeafPrims=pmax)
        n8, p8 = b8.ToHost(); b8.Free()
        o8, op8 = O.build_bvh8_optimal(n2, n, pmax)
        c8, cp8 = O.canon_bvh8(n8, p8)
        oc8, ocp8 = O.canon_bvh8(o8, op8)
        assert c8.shape == oc8.shape and (c8 == oc8).all() and (cp8 == ocp8).all()
    with pytest.raises(nx.NexusError):
        nx.BuildBVH8(ctx, prims[:100], collapse=7)


def test_refit_keeps_the_topology_and_reproduces_a_build_on_unchanged_boxes(ctx):
    """nx_bvh8_refit_aabb (TLAS refit): refitting with the boxes the tree was built from reproduces the built
    nodes bit for bit (a collapsed node's child boxes ARE the unions of what lies below them); refitting with moved boxes keeps
    topology and leaf order, and every node's decoded child box contains the boxes of the primitives below it."""
    rng = np.random.default_rng(11)
    for n in (1, 2, 9, 300, 5000):
        c = rng.uniform(-50, 50, size=(n, 3)).astype(np.float32)
        h = rng.uniform(0.1, 3.0, size=(n, 3)).astype(np.float32)
        boxes = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
        bvh = nx.BuildBVH8(ctx, boxes, prioritizeSpeed=False, collapse=nx.COLLAPSE_SAH_OPTIMAL, maxLeafPrims=1)
        built, order = bvh.ToHost()
        bvh.RefitAABB(boxes)
        again, order2 = bvh.ToHost()
        assert (again == built).all() and (order2 == order).all(), n
        moved = boxes.copy()
        sel = rng.random(n) < 0.3
        moved[sel] += np.tile(rng.uniform(-20, 20, size=(int(sel.sum()), 3)).astype(np.float32), 2)
        bvh.RefitAABB(moved)
        nodes, order3 = bvh.ToHost()
        assert (order3 == order).all()
        assert (nodes[:, 4:8] == built[:, 4:8]).all() and ((nodes[:, 3] >> 24) == (built[:, 3] >> 24)).all()     # childBase, primBase, meta, imask
        assert np.allclose(bvh.bounds, np.concatenate([moved[:, :3].min(0), moved[:, 3:].max(0)]))
        # containment, top-down: the decoded box of every leaf child holds its primitives, of every inner child the child's own frame
        nb = nodes.view(np.uint8).reshape(-1, 80)
        p = nodes[:, 0:3].copy().view(np.float32)
        e = nb[:, 12:15].astype(np.int32)
        cell = np.ldexp(1.0, e - 127)
        for i in range(len(nodes)):
            imask = int(nb[i, 15]); child_base = int(nodes[i, 4]); prim_base = int(nodes[i, 5])
            for s in range(8):
                m = int(nb[i, 24 + s])
                if not m:
                    continue
                lo = p[i] + cell[i] * nb[i, [32 + s, 40 + s, 48 + s]]
                hi = p[i] + cell[i] * nb[i, [56 + s, 64 + s, 72 + s]]
                if (m & 31) >= 24:
                    ch = child_base + bin(imask & ((1 << s) - 1)).count("1")
                    assert (p[ch] >= lo - 1e-4).all() and (p[ch] + cell[ch] * 255 >= lo - 1e-4).all(), (n, i, s)
                    # the child's frame origin is its box minimum: inside the parent's box for it
                    assert (p[ch] <= hi + 1e-4).all()
                else:
                    for t in range(bin(m >> 5).count("1")):
                        b = moved[order3[prim_base + (m & 31) + t]]
                        assert (b[:3] >= lo - 1e-4).all() and (b[3:] <= hi + 1e-4).all(), (n, i, s)
        bvh.Free()


def test_build_and_free_cycles_return_their_device_memory():
    """BuildBVH2 / BuildBVH8 (both collapses) / FreeDeviceBVH, eight times on one context: free device memory after the last cycle is
    what it was after the first, and closing the context returns the workspace pools (BVHBuilder.h:58-66 frees per handle; the
    temporaries of a build - keys, cluster tables, C(n, i) tables - are the context's here)."""
    import torch
    from nexus_b200 import scenes
    def free_mb():
        torch.cuda.synchronize()
        return torch.cuda.mem_get_info()[0] / 2 ** 20
    before = free_mb()
    ctx = nx.Context(0)
    prims = scenes.test_triangles(300_000, seed=4)
    marks = []
    for k in range(8):
        b2 = nx.BuildBVH2(ctx, prims, prioritizeSpeed=bool(k & 1))
        b8 = nx.BuildBVH8(ctx, prims, prioritizeSpeed=bool(k & 1))
        b8o = nx.BuildBVH8(ctx, prims, prioritizeSpeed=True, collapse=nx.COLLAPSE_SAH_OPTIMAL, maxLeafPrims=2)
        assert b2.h.node_count == 2 * len(prims) - 1 and b8.nodeCount > b8o.nodeCount > 0
        b2.Free(); b8.Free(); b8o.Free()
        marks.append(free_mb())
    assert abs(marks[-1] - marks[0]) <= 8.0, marks
    ctx.close()
    assert before - free_mb() <= 64.0, (before, marks)


@pytest.mark.parametrize("speed", [True, False])
def test_every_hploc_kernel_builds_the_same_tree(speed, monkeypatch):
    """The three H-PLOC organisations in the library (NX_HPLOC=0: one phase, clusters in registers; 1: block-local phase in shared memory
    + global phase; 2, the default: one phase with a per-warp shared-memory merge table) perform the same merges: the same BVH2 after
    canonical renumbering, for both key widths, on a mesh-like and on a clustered input."""
    from nexus_b200 import scenes
    inputs = [scenes.test_triangles(150_000, seed=21), np.asarray(scenes.uv_sphere(96, 80)).reshape(-1, 9).astype(np.float32)]
    canon = {}
    for mode in ("2", "0", "1"):
        monkeypatch.setenv("NX_HPLOC", mode)
        c = nx.Context(0)
        for k, prims in enumerate(inputs):
            b2 = nx.BuildBVH2(c, prims, prioritizeSpeed=speed)
            got = O.canon_bvh2(b2.ToHost(), len(prims))
            b2.Free()
            if mode == "2":
                canon[k] = got
            else:
                assert got.shape == canon[k].shape and (got == canon[k]).all(), (mode, k)
        c.close()
