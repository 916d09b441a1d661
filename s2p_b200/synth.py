"""Seeded synthetic rectified stereo tiles.

Texture = uniform ints in [0, 4096) blurred by a Gaussian (sigma 1.0): 12-bit-like
dynamic range as in Pleiades imagery.  Ground-truth disparity = smooth waves plus a
few step "buildings" so that occlusions exist; ``sec(x, y) = tex(y, x + pad - d)``
and ``ref = tex[:, pad:pad+W]``, so the true disparity of ref pixel x is ``d`` in the
s2p convention (x_sec = x_ref + d).  Optional NaN strips exercise the no-data
handling of the matcher (main_mgm.cc:172-173,210-216 in the reference).
"""
import numpy as np


def _gauss1d(sigma, radius):
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-x * x / (2 * sigma * sigma))
    return k / k.sum()


def _blur(a, sigma=1.0):
    k = _gauss1d(sigma, 4)
    a = np.pad(a, 4, mode="reflect")
    a = np.apply_along_axis(lambda r: np.convolve(r, k, mode="valid"), 1, a)
    a = np.apply_along_axis(lambda r: np.convolve(r, k, mode="valid"), 0, a)
    return a


def make_pair(h, w, dmin, dmax, seed=0, nan_border=0.0):
    """-> (ref, sec, gt_disp) float32 arrays of shape (h, w)."""
    rng = np.random.default_rng(seed)
    D = dmax - dmin + 1
    amp = 0.4 * D / 2.0
    mid = 0.5 * (dmin + dmax)
    pad = int(np.ceil(amp + abs(mid))) + 8
    tex = _blur(rng.integers(0, 4096, size=(h, w + 2 * pad)).astype(np.float64))
    yy, xx = np.mgrid[0:h, 0:w]
    a, b = 0.6 * amp, 0.4 * amp
    d = mid + a * np.sin(xx / 80.0 + seed) + b * np.cos(yy / 60.0 - seed)
    for _ in range(4):  # step "buildings"
        bw, bh = rng.integers(w // 10 + 1, w // 4 + 2), rng.integers(h // 10 + 1, h // 4 + 2)
        x0, y0 = rng.integers(0, max(1, w - bw)), rng.integers(0, max(1, h - bh))
        d[y0:y0 + bh, x0:x0 + bw] += rng.uniform(-0.3, 0.3) * amp
    d = np.clip(np.round(d), dmin + 2, dmax - 2)
    ref = tex[:, pad:pad + w]
    # sec(x) = tex(x + pad - d_sec(x)); build it by forward-mapping with a z-buffer-free
    # gather on the ref-frame disparity (adequate for a benchmark texture)
    xs = np.clip(xx + pad - d.astype(np.int64), 0, w + 2 * pad - 1)
    sec = tex[yy, xs]
    ref = ref.astype(np.float32).copy()
    sec = sec.astype(np.float32).copy()
    if nan_border > 0:
        n = max(1, int(round(nan_border * w)))
        ref[:, :n] = np.nan
        sec[:n // 2 + 1, :] = np.nan
        sec[:, w - n:] = np.nan
    return ref, sec, d.astype(np.float32)
