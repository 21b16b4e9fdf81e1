"""The C-ABI library loads and exports every entry point include/nexus_b200.h declares; POD layouts match the reference's
device layouts.  No compute calls: runs without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nexus_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|void\*|const char\*)\s*\*?\s*(nx_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_three_surfaces():
    names = declared_functions()
    assert len(names) >= 50, names
    for must in ("nx_bvh2_build_tri", "nx_bvh8_build_aabb", "nx_bvh2_to_host", "nx_bvh8_benchmark",          # builder surface
                 "nx_scene_add_mesh", "nx_scene_add_instance", "nx_scene_update", "nx_scene_set_hdr_map",     # scene surface
                 "nx_renderer_create", "nx_renderer_render", "nx_renderer_read_accum", "nx_renderer_resize",  # render surface
                 "nx_trace_closest", "nx_trace_any", "nx_write_pfm", "nx_write_exr"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    from nexus_b200._capi import lib
    L = lib()
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    assert L.nx_abi_version() == 1


def test_no_torch_types_in_the_abi():
    text = open(HEADER).read()
    assert "torch" not in text and "at::" not in text and "#include <cuda" not in text


def test_pod_layouts_match_reference_device_layouts():
    from nexus_b200 import _capi as K
    assert C.sizeof(K.Aabb) == 24            # NXB::AABB
    assert C.sizeof(K.MaterialPod) == 92     # D_Material
    assert C.sizeof(K.Bvh8) == 56            # NXB::BVH8 handle
    assert C.sizeof(K.Bvh2) == 40
    assert C.sizeof(K.BuildMetrics) == 36
    import nexus_b200 as nx
    assert nx.RAY_DTYPE.itemsize == 32 and nx.HIT_DTYPE.itemsize == 20


def test_context_creation_fails_loudly_without_a_gpu():
    """The product has no CPU path: on a machine without a CUDA device creating a context raises."""
    import torch
    import nexus_b200 as nx
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nx.NexusError):
        nx.Context(0)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under nexus_b200/ may import, include or link it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "nexus_b200")):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle_lib|liboracle|oracle/|orc_[a-z]", src):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_image_writers_roundtrip(tmp_path):
    """nx_write_pfm / nx_write_exr are host-only: the files parse back to the same pixels, and both take the renderer's layout
    (row 0 = bottom row of the image): PFM stores rows bottom to top, so the file body IS the input; EXR scanline 0 is the top
    row, so the last scanline chunk is input row 0."""
    import numpy as np
    import nexus_b200 as nx
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 4, (5, 7, 3)).astype(np.float32)
    p = tmp_path / "a.pfm"
    nx.write_pfm(p, img)
    raw = open(p, "rb").read()
    head, dims, scale, body = raw.split(b"\n", 3)
    assert head == b"PF" and dims == b"7 5" and float(scale) < 0
    back = np.frombuffer(body, "<f4").reshape(5, 7, 3)
    assert (back == img).all()
    e = tmp_path / "a.exr"
    nx.write_exr(e, img)
    raw = open(e, "rb").read()
    assert raw[:4] == bytes([0x76, 0x2f, 0x31, 0x01])
    # last scanline chunk: y, size, then B, G, R planes
    row = np.frombuffer(raw[-7 * 12:], "<f4").reshape(3, 7)
    assert (row[2] == img[0, :, 0]).all() and (row[1] == img[0, :, 1]).all() and (row[0] == img[0, :, 2]).all()
    # first scanline chunk (behind the header and the 5-entry offset table) = top row = input row 4
    import struct
    hdr_end = raw.index(b"screenWindowWidth\x00float\x00") + len(b"screenWindowWidth\x00float\x00") + 4 + 4 + 1
    off0 = struct.unpack_from("<Q", raw, hdr_end)[0]
    y0, size0 = struct.unpack_from("<ii", raw, off0)
    assert y0 == 0 and size0 == 7 * 12
    top = np.frombuffer(raw[off0 + 8: off0 + 8 + 7 * 12], "<f4").reshape(3, 7)
    assert (top[2] == img[4, :, 0]).all()
