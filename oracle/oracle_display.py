"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's display transform.

AccumulateKernel (Nexus/src/Cuda/PathTracer/PathTracer.cu:527-548): colour = accumulation * 2^exposure, one of the tone
curves of Nexus/src/Utils/ColorUtils.h (ACES :72-97, Uncharted 2 :99-117, AgX :119-212), LinearToGamma (:27-31), ToColorUInt
(:46-56).  Pinned against the reference's own kernel output in tests/golden/display_ref.npz (scripts/make_golden_display.py).
Only tests/ may import this module; the product never does.
"""
import numpy as np

NONE, ACES, UNCHARTED2, AGX_DEFAULT, AGX_GOLDEN, AGX_PUNCHY = range(6)   # ColorUtils::ToneMapping, ColorUtils.h:9-16


def _clamp(x, lo, hi):
    # clamp = fminf(fmaxf(x, lo), hi): a NaN operand is dropped by fmaxf, so NaN -> lo
    return np.fmin(np.fmax(x, lo), hi)


def _aces(c):      # ColorUtils.h:72-97 (S. Hill's RRT + ODT fit)
    m_in = np.array([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]], np.float32)
    m_out = np.array([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605], [-0.00327, -0.07276, 1.07602]], np.float32)
    v = c @ m_in.T
    v = (v * (v + np.float32(0.0245786)) - np.float32(0.000090537)) / (v * (np.float32(0.983729) * v + np.float32(0.4329510)) + np.float32(0.238081))
    return _clamp(v @ m_out.T, 0.0, 1.0)


def _hable(x):     # ColorUtils.h:99-110
    a, b, c, d, e, f = (np.float32(v) for v in (0.15, 0.50, 0.10, 0.20, 0.02, 0.30))
    return (x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f) - e / f


def _uncharted2(c):  # ColorUtils.h:112-117
    return _hable(np.float32(1.6) * c) * (np.float32(1.0) / _hable(np.float32(11.2)))


def _agx(c, mode):   # ColorUtils.h:119-212 (B. Wrensch's minimal AgX)
    inset = np.array([[0.842479062253094, 0.0784335999999992, 0.0792237451477643], [0.0423282422610123, 0.878468636469772, 0.0791661274605434],
                      [0.0423756549057051, 0.0784336, 0.879142973793104]], np.float32)
    outset = np.array([[1.19687900512017, -0.0980208811401368, -0.0990297440797205], [-0.0528968517574562, 1.15190312990417, -0.0989611768448433],
                       [-0.0529716355144438, -0.0980434501171241, 1.15107367264116]], np.float32)
    lo, hi = np.float32(-12.47393), np.float32(4.026069)
    v = c @ inset.T
    with np.errstate(divide="ignore", invalid="ignore"):
        v = (_clamp(np.log2(v), lo, hi) - lo) / (hi - lo)
    x2 = v * v; x4 = x2 * x2; x6 = x4 * x2
    v = (np.float32(-17.86) * x6 * v + np.float32(78.01) * x6 - np.float32(126.7) * x4 * v + np.float32(92.06) * x4 - np.float32(28.72) * x2 * v
         + np.float32(4.361) * x2 - np.float32(0.1718) * v + np.float32(0.002857))
    slope, power, sat = np.ones(3, np.float32), np.ones(3, np.float32), np.float32(1.0)
    if mode == AGX_GOLDEN:
        slope, power, sat = np.array([1.0, 0.9, 0.5], np.float32), np.full(3, 0.8, np.float32), np.float32(0.8)
    elif mode == AGX_PUNCHY:
        power, sat = np.full(3, 1.35, np.float32), np.float32(1.4)
    with np.errstate(invalid="ignore"):
        v = np.power(v * slope, power)
    luma = (v @ np.array([0.2126, 0.7152, 0.0722], np.float32))[..., None]
    v = luma + sat * (v - luma)
    with np.errstate(invalid="ignore"):
        return np.power(v @ outset.T, np.float32(2.2))


def display(rgb, mode, exposure=0.0):
    """(…, 3) linear float32 -> (…) packed RGBA8 (alpha 255), as AccumulateKernel writes the render buffer."""
    c = np.asarray(rgb, np.float32) * np.exp2(np.float32(exposure))
    if mode == ACES:
        c = _aces(c)
    elif mode == UNCHARTED2:
        c = _uncharted2(c)
    elif mode in (AGX_DEFAULT, AGX_GOLDEN, AGX_PUNCHY):
        c = _agx(c, mode)
    with np.errstate(invalid="ignore"):
        c = np.power(c.astype(np.float32), np.float32(1.0 / 2.2))
    q = (_clamp(c, 0.0, 1.0) * np.float32(255.0)).astype(np.uint32)
    return q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16) | np.uint32(0xff000000)


def unpack(rgba):
    rgba = np.asarray(rgba, np.uint32)
    return np.stack([rgba & 0xff, (rgba >> 8) & 0xff, (rgba >> 16) & 0xff], -1).astype(np.int32)
