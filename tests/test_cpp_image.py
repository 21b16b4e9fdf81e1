"""include/nexus_b200_image.hpp (the C++ host layer's PNG / JPEG decoder: what stb_image does for the reference,
src/Assets/IMGLoader.cpp:13-43) against Pillow, which the Python host layer uses for the same job.  Host only, no GPU.

PNG is lossless: every pixel must be equal.  JPEG decoders legitimately differ by the rounding of the inverse DCT, the colour
transform and the chroma interpolation (libjpeg's integer IDCT and fixed-point colour tables against floating point here): 4:4:4
and greyscale files must agree to 3 grey levels everywhere and half a level on average, subsampled files to 1.5 levels on average
and a few levels at worst on a smooth image."""
import io
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "image_check")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), EXE + ".cpp", "-o", EXE])
    return EXE


def decode(exe, path, tmp_path):
    out = str(tmp_path / "out.rgba")
    r = subprocess.run([exe, str(path), out], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.strip())
    blob = open(out, "rb").read()
    head, _, rest = blob.partition(b"\n")
    w, h = (int(v) for v in head.split())
    return np.frombuffer(rest, np.uint8).reshape(h, w, 4)


def pillow(path):
    return np.asarray(Image.open(path).convert("RGBA"), np.uint8)


def picture(w, h, seed=0, smooth=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([x * 255.0 / max(w - 1, 1), y * 255.0 / max(h - 1, 1), (x + y) * 255.0 / max(w + h - 2, 1), 255.0 - x * 200.0 / max(w - 1, 1)], -1)
    if not smooth:
        base = base + rng.normal(0, 40, base.shape)
    return np.clip(base, 0, 255).astype(np.uint8)


def chunk(kind, body):
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xffffffff)


def raw_png(w, h, depth, ctype, rows, interlace=0, extra=b"", level=6):
    """rows: the filtered scanline bytes (filter byte included), already in pass order for interlaced files."""
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)) + extra +
            chunk(b"IDAT", zlib.compress(rows, level)) + chunk(b"IEND", b""))


def test_png_modes_equal_pillow(exe, tmp_path):
    px = picture(37, 23, 1)
    cases = {
        "rgb": Image.fromarray(px[..., :3], "RGB"), "rgba": Image.fromarray(px, "RGBA"), "l": Image.fromarray(px[..., 0], "L"),
        "la": Image.fromarray(np.ascontiguousarray(px[..., [0, 3]]), "LA"), "bilevel": Image.fromarray(px[..., 0], "L").convert("1"),
        "p": Image.fromarray(px[..., :3], "RGB").quantize(colors=200),
        "p16": Image.fromarray(px[..., :3], "RGB").quantize(colors=13),          # 4-bit palette indices
    }
    for name, im in cases.items():
        for opts in ({}, {"compress_level": 0}, {"optimize": True}):                   # stored blocks, fixed / dynamic Huffman
            path = tmp_path / f"{name}.png"
            im.save(path, **opts)
            assert (decode(exe, path, tmp_path) == pillow(path)).all(), (name, opts)
    # palette transparency
    p = Image.fromarray(px[..., :3], "RGB").quantize(colors=50)
    path = tmp_path / "ptrns.png"
    p.save(path, transparency=bytes((i * 5) % 256 for i in range(50)))
    assert (decode(exe, path, tmp_path) == pillow(path)).all()
    # a larger image with every filter type in use (Pillow picks filters per row) and long matches
    big = picture(640, 480, 2)
    path = tmp_path / "big.png"
    Image.fromarray(big, "RGBA").save(path)
    assert (decode(exe, path, tmp_path) == big).all()


def test_png_interlaced_sixteen_bit_and_colour_key(exe, tmp_path):
    rng = np.random.default_rng(3)
    w, h = 19, 11
    px = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # Adam7, filter type 0 on every pass row
    rows = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = px[y0::dy, x0::dx]
        if sub.size:
            rows += b"".join(b"\x00" + r.tobytes() for r in sub)
    path = tmp_path / "adam7.png"
    path.write_bytes(raw_png(w, h, 8, 2, rows, interlace=1))
    got = decode(exe, path, tmp_path)
    assert (got[..., :3] == px).all() and (got[..., 3] == 255).all()
    assert (pillow(path) == got).all()
    # 16-bit RGB with Sub / Up / Average / Paeth filters written by hand: the decoder keeps the high byte
    px16 = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
    be = px16.astype(">u2").view(np.uint8).reshape(h, w * 6).astype(np.int32)
    rows, prev = b"", np.zeros(w * 6, np.int32)
    for y in range(h):
        f = 1 + y % 4
        cur = be[y]
        left = np.concatenate([np.zeros(6, np.int32), cur[:-6]])
        upleft = np.concatenate([np.zeros(6, np.int32), prev[:-6]])
        if f == 1:
            enc = cur - left
        elif f == 2:
            enc = cur - prev
        elif f == 3:
            enc = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            enc = cur - pred
        rows += bytes([f]) + (enc & 255).astype(np.uint8).tobytes()
        prev = cur
    path = tmp_path / "rgb16.png"
    path.write_bytes(raw_png(w, h, 16, 2, rows))
    got = decode(exe, path, tmp_path)
    assert (got[..., :3] == (px16 >> 8).astype(np.uint8)).all()
    # colour key: the pixels equal to the tRNS colour become transparent
    key = px[3, 4]
    path = tmp_path / "key.png"
    path.write_bytes(raw_png(w, h, 8, 2, b"".join(b"\x00" + r.tobytes() for r in px), extra=chunk(b"tRNS", struct.pack(">HHH", *[int(v) for v in key]))))
    got = decode(exe, path, tmp_path)
    assert got[3, 4, 3] == 0 and (got[..., 3] == np.where((px == key).all(-1), 0, 255)).all()
    assert (pillow(path) == got).all()


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_jpeg_baseline_agrees_with_pillow(exe, tmp_path, subsampling):
    px = picture(163, 117, 4, smooth=True)[..., :3]
    for quality in (95, 75):
        path = tmp_path / f"s{subsampling}_{quality}.jpg"
        Image.fromarray(px, "RGB").save(path, quality=quality, subsampling=subsampling)
        got, want = decode(exe, path, tmp_path).astype(np.int32), pillow(path).astype(np.int32)
        d = np.abs(got - want)
        assert (got[..., 3] == 255).all()
        if subsampling == 0:
            assert d.max() <= 3 and d.mean() <= 0.5, (quality, d.max(), d.mean())
        else:
            assert d.mean() <= 1.5 and d.max() <= 12, (quality, d.mean(), d.max())


def test_jpeg_greyscale_restart_intervals_and_noise(exe, tmp_path):
    px = picture(100, 75, 5)
    path = tmp_path / "grey.jpg"
    Image.fromarray(px[..., 0], "L").save(path, quality=90)
    got, want = decode(exe, path, tmp_path).astype(np.int32), pillow(path).astype(np.int32)
    assert np.abs(got - want).max() <= 3 and np.abs(got - want).mean() <= 0.5
    # noisy colour image, 4:4:4, with restart markers when this Pillow can write them
    path = tmp_path / "noise.jpg"
    try:
        Image.fromarray(px[..., :3], "RGB").save(path, quality=92, subsampling=0, restart_marker_blocks=5)
    except TypeError:
        Image.fromarray(px[..., :3], "RGB").save(path, quality=92, subsampling=0)
    got, want = decode(exe, path, tmp_path).astype(np.int32), pillow(path).astype(np.int32)
    assert np.abs(got - want).max() <= 3 and np.abs(got - want).mean() <= 0.5
    if b"\xff\xdd" in path.read_bytes():
        path2 = tmp_path / "noise420.jpg"
        Image.fromarray(px[..., :3], "RGB").save(path2, quality=85, subsampling=2, restart_marker_rows=1)
        got, want = decode(exe, path2, tmp_path).astype(np.int32), pillow(path2).astype(np.int32)
        assert np.abs(got - want).mean() <= 3.0


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_jpeg_progressive_agrees_with_pillow(exe, tmp_path, subsampling):
    """Progressive files (SOF2: DC / AC scans with spectral selection and successive approximation, interleaved DC scans and
    per-component AC scans; three of the reference's seven demo scenes embed such textures): the same coefficients as the sequential
    encoding of the same image at the same quality, so the decoded pixels must equal OUR decoding of the baseline file exactly and
    Pillow's within the usual decoder rounding.  Smooth and noisy content, odd sizes (partial MCUs), greyscale, restart intervals."""
    for seed, smooth, size in ((7, True, (163, 117)), (8, False, (97, 61)), (9, False, (16, 8))):
        px = picture(size[0], size[1], seed, smooth=smooth)[..., :3]
        for quality in (90, 60):
            prog, base = tmp_path / "p.jpg", tmp_path / "b.jpg"
            Image.fromarray(px, "RGB").save(prog, quality=quality, subsampling=subsampling, progressive=True)
            Image.fromarray(px, "RGB").save(base, quality=quality, subsampling=subsampling, progressive=False)
            assert b"\xff\xc2" in prog.read_bytes()[:2000]
            got = decode(exe, prog, tmp_path).astype(np.int32)
            assert (got == decode(exe, base, tmp_path).astype(np.int32)).all(), (seed, quality)
            d = np.abs(got - pillow(prog).astype(np.int32))
            if subsampling == 0:
                assert d.max() <= 3 and d.mean() <= 0.5, (seed, quality, d.max(), d.mean())
            elif smooth:
                assert d.mean() <= 1.5 and d.max() <= 12, (seed, quality, d.mean(), d.max())
            else:
                assert d.mean() <= 3.0, (seed, quality, d.mean())
    grey = picture(75, 50, 10)[..., 0]
    path = tmp_path / "pg.jpg"
    Image.fromarray(grey, "L").save(path, quality=85, progressive=True)
    d = np.abs(decode(exe, path, tmp_path).astype(np.int32) - pillow(path).astype(np.int32))
    assert d.max() <= 3 and d.mean() <= 0.5
    path = tmp_path / "pr.jpg"
    Image.fromarray(picture(120, 90, 11)[..., :3], "RGB").save(path, quality=88, subsampling=subsampling, progressive=True, restart_marker_blocks=3)
    assert b"\xff\xdd" in path.read_bytes()
    d = np.abs(decode(exe, path, tmp_path).astype(np.int32) - pillow(path).astype(np.int32))
    assert d.mean() <= (0.5 if subsampling == 0 else 3.0)


def test_unsupported_and_broken_files_fail_loudly(exe, tmp_path):
    px = picture(40, 30, 6)[..., :3]
    (tmp_path / "x.bin").write_bytes(b"GIF89a....")
    with pytest.raises(RuntimeError, match="neither"):
        decode(exe, tmp_path / "x.bin", tmp_path)
    good = tmp_path / "ok.png"
    Image.fromarray(px, "RGB").save(good)
    blob = good.read_bytes()
    (tmp_path / "cut.png").write_bytes(blob[:len(blob) // 2])
    with pytest.raises(RuntimeError):
        decode(exe, tmp_path / "cut.png", tmp_path)
