"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's elastic deformation
(reference utils/augmentations.py:4-39, called per sample by
utils/dataset.py:83-94 with alpha = 2000, sigma = 20 from scripts/train.py:34-36).

Two restatements:

``elastic_deform_scipy``  follows the reference call by call with the reference's own third-party
    routines (``scipy.ndimage.gaussian_filter`` / ``map_coordinates``; scipy is un-pinned in
    requirements.txt:8, 1.18.1 is installed here) — this is what pins the second one.

``elastic_deform_steps``  spells out the arithmetic of those two scipy routines for this call
    pattern in plain numpy, in scipy's operation ORDER, and is bit-identical to them (float64):
      * Gaussian taps: exp(-0.5 / sigma^2 * x^2) / sum, radius int(4 sigma + 0.5);
      * separable correlation, axis 0 then axis 1, zero outside the image (mode='constant'),
        symmetric accumulation  acc = f[c] w[c];  acc += (f[c-d] + f[c+d]) w[d]  for d = r … 1;
      * coordinates y + dy, x + dx; out-of-range coordinates are folded back with scipy's
        'reflect' rule (half-sample symmetric, applied for c < 0 and c > n - 1);
      * image: bilinear, weights (1 - t) and 1 - (1 - t), products (f * wy) * wx summed in row-major
        footprint order, result rounded half up and clamped into uint8;
      * mask: nearest (floor(c + 0.5)), label values copied.
    The CUDA kernels (csrc/elastic.cuh) follow exactly these steps with IEEE round-to-nearest adds /
    multiplies (no FMA contraction), so they are bit-exact against both.

Pinning: the reference holds no fixtures for this function (its output depends on a random seed);
tests/test_oracle.py compares both restatements with the live reference function on seeded inputs.
"""
from __future__ import annotations

import numpy as np


def reference_noise(seed: int, shape):
    """The two uniform [0, 1) fields the reference draws (augmentations.py:27-28):
    RandomState(seed).rand(*shape) for dx first, then for dy."""
    rs = np.random.RandomState(seed)
    return rs.rand(*shape), rs.rand(*shape)


def elastic_deform_scipy(image, mask, alpha, sigma, seed):
    from scipy.ndimage import gaussian_filter, map_coordinates

    u, v = reference_noise(seed, image.shape)
    dx = gaussian_filter(u * 2 - 1, sigma, mode="constant", cval=0) * alpha
    dy = gaussian_filter(v * 2 - 1, sigma, mode="constant", cval=0) * alpha
    rows, cols = np.indices(image.shape)
    at = ((rows + dy).reshape(-1, 1), (cols + dx).reshape(-1, 1))
    return (map_coordinates(image, at, order=1, mode="reflect").reshape(image.shape),
            map_coordinates(mask, at, order=0, mode="reflect").reshape(image.shape))


# ---- step-by-step -------------------------------------------------------------------------------
def gaussian_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    sd = float(sigma)
    r = int(truncate * sd + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum()


def _correlate_axis0_zero_padded(f: np.ndarray, w: np.ndarray) -> np.ndarray:
    r = (len(w) - 1) // 2
    n = f.shape[0]
    pad = np.zeros((r,) + f.shape[1:])
    p = np.concatenate([pad, f, pad], axis=0)
    acc = p[r:r + n] * w[r]
    for d in range(r, 0, -1):
        acc = acc + (p[r - d:r - d + n] + p[r + d:r + d + n]) * w[r - d]
    return acc


def gaussian_blur(field: np.ndarray, sigma: float) -> np.ndarray:
    w = gaussian_taps(sigma)
    once = _correlate_axis0_zero_padded(field, w)
    return _correlate_axis0_zero_padded(once.T, w).T


def fold_coordinate(c: np.ndarray, n: int) -> np.ndarray:
    c = np.asarray(c, dtype=np.float64)
    out = c.copy()
    lo, hi = c < 0, c > n - 1
    if n <= 1:
        out[lo | hi] = 0.0
        return out
    period = 2.0 * n
    x = c[lo]
    x = np.where(x < -period, period * np.trunc(-x / period) + x, x)
    out[lo] = np.where(x < -n, x + period, -x - 1)
    x = c[hi]
    x = x - period * np.trunc(x / period)
    out[hi] = np.where(x >= n, period - x - 1, x)
    return out


def fold_index(i: np.ndarray, n: int) -> np.ndarray:
    """d c b a | a b c d | d c b a for integer footprint positions."""
    if n <= 1:
        return np.zeros_like(i)
    i = np.where(i < 0, -i - 1, i) % (2 * n)
    return np.where(i >= n, 2 * n - 1 - i, i)


def sample_bilinear_u8(image: np.ndarray, cy: np.ndarray, cx: np.ndarray) -> np.ndarray:
    h, w = image.shape
    y, x = fold_coordinate(cy, h), fold_coordinate(cx, w)
    y0, x0 = np.floor(y), np.floor(x)
    wy0, wx0 = 1 - (y - y0), 1 - (x - x0)
    wy1, wx1 = 1.0 - wy0, 1.0 - wx0
    r0, r1 = fold_index(y0.astype(np.int64), h), fold_index(y0.astype(np.int64) + 1, h)
    c0, c1 = fold_index(x0.astype(np.int64), w), fold_index(x0.astype(np.int64) + 1, w)
    f = image.astype(np.float64)
    t = (f[r0, c0] * wy0) * wx0
    t = t + (f[r0, c1] * wy0) * wx1
    t = t + (f[r1, c0] * wy1) * wx0
    t = t + (f[r1, c1] * wy1) * wx1
    t = np.where(t > 0, t + 0.5, 0.0)
    return np.minimum(t, 255.0).astype(np.uint8)


def sample_nearest(mask: np.ndarray, cy: np.ndarray, cx: np.ndarray) -> np.ndarray:
    h, w = mask.shape
    r = fold_index(np.floor(fold_coordinate(cy, h) + 0.5).astype(np.int64), h)
    c = fold_index(np.floor(fold_coordinate(cx, w) + 0.5).astype(np.int64), w)
    return mask[r, c]


def elastic_deform_steps(image, mask, noise_dx, noise_dy, alpha, sigma):
    """image (H, W) uint8, mask (H, W) uint8/uint16, noise_* (H, W) float64 in [0, 1)."""
    dx = gaussian_blur(noise_dx * 2 - 1, sigma) * alpha
    dy = gaussian_blur(noise_dy * 2 - 1, sigma) * alpha
    rows, cols = np.indices(image.shape)
    cy, cx = rows + dy, cols + dx
    return sample_bilinear_u8(image, cy, cx), sample_nearest(mask, cy, cx)
