"""Radiance .hdr reader behind Scene.AddHDRMap(filePath, fileName) (the reference loads environment maps with stb_image's stbi_loadf,
src/Assets/IMGLoader.cpp:13-31): flat and run-length encoded files decode to the same floats stb's formula gives."""
import numpy as np
import pytest

from nexus_b200 import hdr


def _rgbe(img):
    """float RGB -> RGBE bytes (the standard encoder: shared exponent of the largest component)."""
    m = img.max(axis=-1)
    out = np.zeros(img.shape[:-1] + (4,), np.uint8)
    nz = m > 1e-32
    mant, exp = np.frexp(m[nz])
    out[nz, :3] = (img[nz] * (mant * 256.0 / m[nz])[:, None]).astype(np.uint8)
    out[nz, 3] = (exp + 128).astype(np.uint8)
    return out


def _rle_channel(row):
    out, i = bytearray(), 0
    while i < len(row):
        run = 1
        while i + run < len(row) and run < 127 and row[i + run] == row[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, int(row[i])]); i += run
        else:
            j = i
            while j < len(row) and j - i < 128 and not (j + 3 < len(row) and row[j] == row[j + 1] == row[j + 2] == row[j + 3]):
                j += 1
            out += bytes([j - i]) + bytes(int(v) for v in row[i:j]); i = j
    return bytes(out)


def _write(path, rgbe, rle):
    h, w = rgbe.shape[:2]
    body = bytearray()
    for y in range(h):
        if rle:
            body += bytes([2, 2, w >> 8, w & 255])
            for c in range(4):
                body += _rle_channel(rgbe[y, :, c])
        else:
            body += rgbe[y].tobytes()
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# made by a test\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (h, w) + bytes(body))


def test_flat_and_rle_files_decode_alike(tmp_path):
    rs = np.random.RandomState(9)
    img = rs.uniform(0, 4, (12, 40, 3)).astype(np.float32)
    img[3, 5:25] = (1000.0, 500.0, 2.0)                 # a long run (the sun) and a black stretch
    img[7, :12] = 0.0
    rgbe = _rgbe(img)
    _write(tmp_path / "flat.hdr", rgbe, rle=False)
    _write(tmp_path / "rle.hdr", rgbe, rle=True)
    a, b = hdr.load_hdr(tmp_path / "flat.hdr"), hdr.load_hdr(tmp_path / "rle.hdr")
    assert a.shape == (12, 40, 4) and a.dtype == np.float32 and (a == b).all() and (a[..., 3] == 1).all()
    want = rgbe[..., :3].astype(np.float32) * np.ldexp(np.float32(1), rgbe[..., 3].astype(np.int32) - 136)[..., None]
    want[rgbe[..., 3] == 0] = 0
    assert (a[..., :3] == want).all()
    # RGBE quantises every component in steps of (largest component's power of two) / 256
    assert (np.abs(a[..., :3] - img) <= img.max(-1, keepdims=True) / 128 + 1e-6).all() and (a[7, :12, :3] == 0).all()


def test_malformed_files_are_rejected(tmp_path):
    (tmp_path / "a.hdr").write_bytes(b"P6\n1 1\n255\n...")
    (tmp_path / "b.hdr").write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_xyze\n\n-Y 1 +X 1\n\0\0\0\0")
    (tmp_path / "c.hdr").write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 1 +X 1\n\0\0\0\0")
    (tmp_path / "d.hdr").write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 16\n" + bytes([2, 2, 0, 16, 200, 1]))
    for name, what in (("a", "not a Radiance"), ("b", "format"), ("c", "orientation"), ("d", "overflows|truncated")):
        with pytest.raises(hdr.HdrError, match=what):
            hdr.load_hdr(tmp_path / (name + ".hdr"))
