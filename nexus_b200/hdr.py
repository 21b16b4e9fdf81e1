"""Radiance RGBE (.hdr) reader for Scene.AddHDRMap(filePath, fileName) (src/Scene/Scene.cpp:102-107 -> IMGLoader::LoadIMG,
src/Assets/IMGLoader.cpp:13-31: stbi_loadf with four channels).  Covers what stb_image's HDR loader covers: the `#?RADIANCE` /
`#?RGBE` signature, FORMAT=32-bit_rle_rgbe, the `-Y h +X w` orientation, flat and new-style run-length encoded scanlines.
Decoding is stb's: (mantissa) * 2^(exponent - 136), alpha 1, exponent 0 -> black.  Pure host code (numpy)."""
import numpy as np


class HdrError(ValueError):
    pass


def load_hdr(path):
    """Returns the image as (h, w, 4) float32, rows top to bottom like stb_image."""
    with open(path, "rb") as f:
        blob = f.read()
    pos = blob.find(b"\n")
    if pos < 0 or blob[:pos].strip() not in (b"#?RADIANCE", b"#?RGBE"):
        raise HdrError("not a Radiance HDR file")
    fmt_ok = False
    while True:
        end = blob.find(b"\n", pos + 1)
        if end < 0:
            raise HdrError("truncated header")
        line = blob[pos + 1:end].strip()
        pos = end
        if not line:
            break
        if line == b"FORMAT=32-bit_rle_rgbe":
            fmt_ok = True
    if not fmt_ok:
        raise HdrError("unsupported format (only 32-bit_rle_rgbe)")
    end = blob.find(b"\n", pos + 1)
    dims = blob[pos + 1:end].split()
    if len(dims) != 4 or dims[0] != b"-Y" or dims[2] != b"+X":
        raise HdrError("unsupported orientation (only -Y h +X w)")
    h, w = int(dims[1]), int(dims[3])
    data = np.frombuffer(blob, np.uint8, offset=end + 1)
    rgbe = np.empty((h, w, 4), np.uint8)
    p = 0
    if w < 8 or w >= 32768 or len(data) == 4 * w * h and not (data[0] == 2 and data[1] == 2 and not data[2] & 0x80):
        if len(data) < 4 * w * h:
            raise HdrError("truncated pixel data")
        rgbe[:] = data[:4 * w * h].reshape(h, w, 4)                        # flat
    else:
        for y in range(h):
            if p + 4 > len(data) or data[p] != 2 or data[p + 1] != 2 or (int(data[p + 2]) << 8 | int(data[p + 3])) != w:
                raise HdrError(f"scanline {y}: bad run-length header")
            p += 4
            for c in range(4):                                             # each channel separately
                x = 0
                while x < w:
                    if p >= len(data):
                        raise HdrError("truncated pixel data")
                    count = int(data[p]); p += 1
                    if count > 128:                                        # run
                        count -= 128
                        if x + count > w or p >= len(data):
                            raise HdrError(f"scanline {y}: run overflows")
                        rgbe[y, x:x + count, c] = data[p]; p += 1
                    else:                                                  # literal
                        if count == 0 or x + count > w or p + count > len(data):
                            raise HdrError(f"scanline {y}: dump overflows")
                        rgbe[y, x:x + count, c] = data[p:p + count]; p += count
                    x += count
    out = np.zeros((h, w, 4), np.float32)
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    out[..., :3] = rgbe[..., :3].astype(np.float32) * scale[..., None]
    out[..., 3] = 1.0
    return out
