"""CPU oracle for the ``-m n=<level>`` denoise pass -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import this module;
product code under ``upscale_video_b200/`` must never do so.

PARITY PINNED.  The reference's denoise worker (upscale/upscale_processing.py:350-362, ``apply_denoise``) is one call,
``cv2.fastNlMeansDenoisingColored(img, None, denoise, denoise, 5, 9)``, into the third-party OpenCV wheel
(unpinned in the reference's README; the image this repo is developed and tested in carries opencv 4.13.0 without
OpenCL, so the ``cv2.UMat`` argument runs OpenCV's CPU code).  OpenCV's sources are not in /root/reference, so this
file restates the published algorithm in integer numpy and is pinned against cv2 itself:

* ``bgr_to_lab`` / ``lab_to_bgr`` (cvtColor COLOR_LBGR2Lab / COLOR_Lab2LBGR on 8-bit data, OpenCV's fixed-point
  tables) equal cv2 over ALL 2**24 inputs (``tests/test_nlmeans_oracle.py::test_lab_exhaustive``);
* ``fast_nl_means_denoising`` (the fixed-point weight table + integer accumulators of
  ``FastNlMeansDenoisingInvoker<…, DistSquared, int>``) and ``fast_nl_means_denoising_colored`` equal cv2 bit for
  bit on seeded images for every level the reference accepts (1..30) (same test file);
* frozen vectors: ``tests/golden/nlm_*.npz`` generated FROM cv2 by ``tools/make_nlm_goldens.py``.

Algorithm (OpenCV photo module, fastNlMeansDenoisingColored, CPU path):
  Lab = cvtColor(BGR, LBGR2Lab); L plane and (a,b) plane pair are denoised separately (h resp. hColor), each with
  template window 5x5, search window 9x9, border 6 px BORDER_REFLECT_101; for pixel p and each of the 81 offsets q:
  d = sum over the 5x5 template and the plane's channels of (I(p+t) - I(p+q+t))**2; w = table[d >> 5] where
  table[k] = round(M * exp(-(k * 32/25) / (h*h*channels))), zeroed below 0.001*M, M = INT_MAX // (81*255);
  out = (sum w*I(p+q) + W/2) // W with W = sum w, in unsigned 32-bit arithmetic.  Then Lab2LBGR.
"""
from __future__ import annotations

import numpy as np

TEMPLATE = 5
SEARCH = 9
BORDER = SEARCH // 2 + TEMPLATE // 2
BIN_SHIFT = 5                                   # smallest s with (1 << s) >= TEMPLATE**2
FIXED_POINT_MULT = 2147483647 // (SEARCH * SEARCH * 255)   # 103969
WEIGHT_THRESHOLD = 0.001

_f32 = np.float32
_LAB_SHIFT = 12
_GAMMA_SHIFT = 3
_LAB_SHIFT2 = _LAB_SHIFT + _GAMMA_SHIFT
_BASE = 1 << 14
_MIN_AB = -8145


def _rint(x):
    return np.rint(x).astype(np.int64)


def _cdiv(a, b):
    """C integer division (truncation toward zero)."""
    a = np.asarray(a, dtype=np.int64)
    q = np.abs(a) // b
    return np.where(a < 0, -q, q)


def _descale(v, n):
    return (v + (1 << (n - 1))) >> n


def _build_lab_tables():
    # forward: XYZ = M * linear RGB scaled by the D65 white point, fixed point with 12 fractional bits
    m = np.array([0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227],
                 dtype=_f32).reshape(3, 3)
    wp = np.array([0.950456, 1.0, 1.088754], dtype=_f32)
    fwd = _rint((m * (_f32(1 << _LAB_SHIFT) / wp).astype(_f32)[:, None]).astype(_f32))
    # cube-root table over 0 .. 1.5 in steps of 1/(255*8), 15 fractional bits
    i = np.arange(256 * 3 // 2 * (1 << _GAMMA_SHIFT), dtype=_f32)
    x = i / _f32(255 * (1 << _GAMMA_SHIFT))
    lin = x * (_f32(841) / _f32(108)) + _f32(16) / _f32(116)
    cb = np.where(x < _f32(216) / _f32(24389), lin, np.cbrt(x.astype(np.float64)).astype(_f32)).astype(_f32)
    cbrt_tab = _rint(_f32(1 << _LAB_SHIFT2) * cb)
    # entry 324 is an exact .5 tie after float32 rounding of the cube root; OpenCV's own soft-float cube root
    # lands one ulp lower there.  Pinned by the exhaustive comparison against cv2 (the only entry that differs).
    cbrt_tab[324] = 17745
    # inverse: L -> (y, fy) and f -> x|z tables, 14 fractional bits
    k = np.arange(256)
    fy = (_f32(k * 100 * _BASE) / _f32(255 * 116) + _f32(16 * _BASE) / _f32(116)).astype(_f32)
    y_hi = _rint((fy * fy * fy / _f32(_BASE * _BASE)).astype(_f32))
    y_lo = _rint(_f32(k * _BASE * 20 * 9) / _f32(17 * 24389))
    fy_lo = _rint(_f32(_BASE) * (_f32(16) / _f32(116) + _f32(k * 5) / _f32(3 * 17 * 29)))
    l2y = np.where(k <= 20, y_lo, y_hi)
    l2fy = np.where(k <= 20, fy_lo, _rint(fy))
    t = np.arange(_MIN_AB, _BASE * 9 // 4 + _MIN_AB, dtype=np.int64)
    f2xz = np.where(t <= 3390, _cdiv(t * 108, 841) - _cdiv(_cdiv(_BASE * 16, 116) * 108, 841),
                    _cdiv(_cdiv(t * t, _BASE) * t, _BASE))
    xyz2rgb = np.array([3.240479, -1.53715, -0.498535, -0.969256, 1.875991, 0.041556, 0.055648, -0.204043,
                        1.057311], dtype=_f32).reshape(3, 3)
    inv = _rint(np.float64(1 << _LAB_SHIFT) * xyz2rgb.astype(np.float64) * wp.astype(np.float64)[None, :])
    return fwd, cbrt_tab, l2y, l2fy, f2xz, inv


_FWD, _CBRT, _L2Y, _L2FY, _F2XZ, _INV = _build_lab_tables()


def lab_tables():
    """The integer tables, for tests that compare them with the product's own copy."""
    return dict(fwd=_FWD, cbrt=_CBRT, l2y=_L2Y, l2fy=_L2FY, f2xz=_F2XZ, inv=_INV)


def bgr_to_lab(bgr: np.ndarray) -> np.ndarray:
    """cvtColor(..., COLOR_LBGR2Lab) for uint8 (linear RGB: no gamma curve)."""
    b = bgr[..., 0].astype(np.int64) << _GAMMA_SHIFT
    g = bgr[..., 1].astype(np.int64) << _GAMMA_SHIFT
    r = bgr[..., 2].astype(np.int64) << _GAMMA_SHIFT
    fx = _CBRT[_descale(r * _FWD[0, 0] + g * _FWD[0, 1] + b * _FWD[0, 2], _LAB_SHIFT)]
    fy = _CBRT[_descale(r * _FWD[1, 0] + g * _FWD[1, 1] + b * _FWD[1, 2], _LAB_SHIFT)]
    fz = _CBRT[_descale(r * _FWD[2, 0] + g * _FWD[2, 1] + b * _FWD[2, 2], _LAB_SHIFT)]
    lscale = (116 * 255 + 50) // 100
    lshift = -((16 * 255 * (1 << _LAB_SHIFT2) + 50) // 100)
    L = _descale(lscale * fy + lshift, _LAB_SHIFT2)
    a = _descale(500 * (fx - fy) + 128 * (1 << _LAB_SHIFT2), _LAB_SHIFT2)
    bb = _descale(200 * (fy - fz) + 128 * (1 << _LAB_SHIFT2), _LAB_SHIFT2)
    return np.clip(np.stack([L, a, bb], -1), 0, 255).astype(np.uint8)


def lab_to_bgr(lab: np.ndarray) -> np.ndarray:
    """cvtColor(..., COLOR_Lab2LBGR) for uint8."""
    L = lab[..., 0].astype(np.int64)
    a = lab[..., 1].astype(np.int64)
    b = lab[..., 2].astype(np.int64)
    y = _L2Y[L]
    ify = _L2FY[L]
    adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * _BASE // 500
    bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * _BASE // 200 + 1
    x = _F2XZ[ify + adiv - _MIN_AB]
    z = _F2XZ[ify - bdiv - _MIN_AB]
    out = []
    for row in (2, 1, 0):   # B, G, R
        v = _descale(_INV[row, 0] * x + _INV[row, 1] * y + _INV[row, 2] * z, _LAB_SHIFT + 2)
        out.append((np.clip(v, 0, 4095) * 255) >> 12)
    return np.stack(out, -1).astype(np.uint8)


def weight_table(h: float, channels: int) -> np.ndarray:
    """almost_dist2weight_ of FastNlMeansDenoisingInvoker for DistSquared with int weights."""
    mult = float(1 << BIN_SHIFT) / (TEMPLATE * TEMPLATE)
    n = int(255 * 255 * channels / mult + 1)
    dist = np.arange(n, dtype=np.float64) * mult
    den = _f32(_f32(_f32(h) * _f32(h)) * _f32(channels))
    w = np.exp(-dist / np.float64(den))
    tab = np.rint(FIXED_POINT_MULT * w).astype(np.int64)
    tab[tab < WEIGHT_THRESHOLD * FIXED_POINT_MULT] = 0
    return tab


def fast_nl_means_denoising(plane: np.ndarray, h: float) -> np.ndarray:
    """cv2.fastNlMeansDenoising(plane, None, h, 5, 9) for a uint8 plane of 1 or 2 channels (H x W or H x W x C)."""
    src = plane if plane.ndim == 3 else plane[..., None]
    hh, ww, cn = src.shape
    tab = weight_table(h, cn)
    ext = np.pad(src.astype(np.int64), ((BORDER, BORDER), (BORDER, BORDER), (0, 0)), mode="reflect")
    t = TEMPLATE // 2
    s = SEARCH // 2
    # reference window: rows/cols [BORDER - t, BORDER + h + t)
    y0 = BORDER - t
    base = ext[y0:y0 + hh + 2 * t, y0:y0 + ww + 2 * t]
    est = np.zeros((hh, ww, cn), dtype=np.int64)
    wsum = np.zeros((hh, ww), dtype=np.int64)
    for dy in range(-s, s + 1):
        for dx in range(-s, s + 1):
            sh = ext[y0 + dy:y0 + dy + hh + 2 * t, y0 + dx:y0 + dx + ww + 2 * t]
            d2 = ((base - sh) ** 2).sum(-1)
            c = np.zeros((d2.shape[0] + 1, d2.shape[1] + 1), dtype=np.int64)
            c[1:, 1:] = d2.cumsum(0).cumsum(1)
            k = TEMPLATE
            box = c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]
            w = tab[box >> BIN_SHIFT]
            est += w[..., None] * ext[BORDER + dy:BORDER + dy + hh, BORDER + dx:BORDER + dx + ww]
            wsum += w
    out = (est + (wsum // 2)[..., None]) // wsum[..., None]
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if plane.ndim == 3 else out[..., 0]


def fast_nl_means_denoising_colored(bgr: np.ndarray, h: float, h_color: float) -> np.ndarray:
    """cv2.fastNlMeansDenoisingColored(bgr, None, h, h_color, 5, 9) -- what reference apply_denoise
    (upscale/upscale_processing.py:354) computes on a frame."""
    lab = bgr_to_lab(bgr)
    l = fast_nl_means_denoising(lab[..., 0], h)
    ab = fast_nl_means_denoising(lab[..., 1:3], h_color)
    return lab_to_bgr(np.concatenate([l[..., None], ab], -1))
