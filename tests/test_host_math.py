"""The scene layer's host arithmetic pinned against the reference's own host code (no GPU needed): the D_MeshInstance record
(transform T * Rz * Ry * Rx * S from position / Euler degrees / scale, its inverse, the world box of the mesh box's eight corners) and
the D_Camera record (right, up, viewport vectors, lower-left corner, lens radius): transforms, boxes and camera records bit for bit, the
inverse to 1e-6 relative.  Golden: tests/golden/host_ref.npz, produced by the
UNMODIFIED MeshInstance::ToDevice / Camera::ToDevice compiled with g++ (scripts/make_golden_host.py); the compiled reference is also
called live when it travelled with the repository."""
import ctypes as C

import numpy as np
import pytest

import nexus_b200 as nx
import oracle_lib as O
from golden_cases import host_cases
from nexus_b200._capi import Aabb, lib

GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "host_ref.npz")
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def _ours():
    inst, cams = host_cases()
    L = lib()
    rec = np.zeros((len(inst["position"]), 160), np.uint8)
    for k in range(len(rec)):
        b = inst["mesh_bounds"][k]
        box = Aabb((C.c_float * 3)(*b[:3]), (C.c_float * 3)(*b[3:]))
        assert L.nx_host_instance_record(P(inst["position"][k]), P(inst["rotation"][k]), P(inst["scale"][k]), C.byref(box),
                                         C.c_uint32(int(inst["mesh_idx"][k])), C.c_uint32(int(inst["material_idx"][k])), P(rec[k])) == 0
    cam = np.zeros((len(cams["position"]), 88), np.uint8)
    for k in range(len(cam)):
        pod = nx.Camera(position=tuple(cams["position"][k]), forward=tuple(cams["forward"][k]), horizontalFOV=float(cams["hfov"][k]),
                        focusDistance=float(cams["focus"][k]), defocusAngle=float(cams["defocus"][k])).pod()
        assert L.nx_host_camera_record(C.byref(pod), C.c_uint32(int(cams["res"][k, 0])), C.c_uint32(int(cams["res"][k, 1])), P(cam[k])) == 0
    cam[:, 76:80] = 0
    return rec, cam


def _check(rec, cam, want_rec, want_cam):
    assert (rec[:, :8] == want_rec[:, :8]).all()                                                   # mesh and material index
    m, wm = rec[:, 8:72].copy().view(np.float32), want_rec[:, 8:72].copy().view(np.float32)         # transform
    inv, winv = rec[:, 72:136].copy().view(np.float32), want_rec[:, 72:136].copy().view(np.float32)
    box, wbox = rec[:, 136:160].copy().view(np.float32), want_rec[:, 136:160].copy().view(np.float32)
    assert (m.view(np.uint32) == wm.view(np.uint32)).all()                                         # same formulas in the same order: the same bits
    assert (box.view(np.uint32) == wbox.view(np.uint32)).all()
    # the inverse is computed by a different (cofactor) formula than Mat4::Inverted: equal to rounding, 1e-7 relative observed
    assert (np.abs(inv - winv) <= 1e-6 * np.abs(winv).max(axis=1, keepdims=True)).all()
    c, wc = cam[:, :76].copy().view(np.float32), want_cam[:, :76].copy().view(np.float32)
    assert (c.view(np.uint32) == wc.view(np.uint32)).all() and (cam[:, 80:] == want_cam[:, 80:]).all()   # frame, viewport, lens; resolution


def test_instance_and_camera_records_equal_the_reference_host_code():
    g = np.load(GOLD)
    rec, cam = _ours()
    _check(rec, cam, g["instance_records"], g["camera_records"])


@pytest.mark.skipif(not O.have_refcpu(), reason="oracle/_ref/libnexus_refcpu.so (the compiled reference host code) is not present")
def test_live_reference_host_code_agrees_with_the_golden_file():
    R = C.CDLL(O.REFCPU_SO)
    inst, cams = host_cases()
    g = np.load(GOLD)
    out = np.zeros(160, np.uint8)
    for k in (0, 5, 17, 40):
        R.ref_host_instance(P(inst["position"][k]), P(inst["rotation"][k]), P(inst["scale"][k]), P(inst["mesh_bounds"][k]),
                            C.c_uint32(int(inst["mesh_idx"][k])), C.c_uint32(int(inst["material_idx"][k])), P(out))
        assert (out == g["instance_records"][k]).all()
