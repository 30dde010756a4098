"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's weight-map preprocessing
(reference scripts/preprocess_data.py:17-77, ``calculate_weight_map``).

``weight_map_full`` follows the reference step by step, INCLUDING the per-instance Euclidean
distance transforms (scipy, as the reference uses it), so that it does not assume what the closed
form below claims. ``weight_map_closed_form`` is what the CUDA kernel implements.

Why a closed form exists: the reference takes, per instance,
``minimum(edt(obj), edt(obj == 0))`` (preprocess_data.py:47). ``edt(obj)`` is zero outside the
object and ``edt(obj == 0)`` is zero inside it, so the minimum is identically zero for every mask;
hence d1 = d2 = 0 (:52-64), the border term is ``w0 * exp(-0.0) = w0`` (:72) and

    weight = float32(class-balance weight of the pixel's class) + w0        (float64)

for EVERY input. Parity target = the stored ``.npy`` maps (168 files, float64), not the U-Net paper.

Pinning: ``tests/test_oracle.py`` checks both functions against the reference's stored maps
(all of sequence 01 live in the authoring container; three committed in
``tests/golden/weight_map_golden.npz``) and against the live ``calculate_weight_map``.
"""
from __future__ import annotations

import numpy as np


def class_balance_weights(n_fg: int, n_total: int):
    """(wc_background, wc_foreground) as the float32 values the reference stores in ``wc_map``
    (preprocess_data.py:26-36): 1 / (class pixel fraction), 0 for an absent class; computed in
    float64 and rounded to float32 by the assignment into a float32 array."""
    n_bg = n_total - n_fg
    wc_bg = 1.0 / (n_bg / n_total) if n_bg > 0 else 0.0
    wc_fg = 1.0 / (n_fg / n_total) if n_fg > 0 else 0.0
    return np.float32(wc_bg), np.float32(wc_fg)


def weight_map_closed_form(labels: np.ndarray, w0=10, sigma=5) -> np.ndarray:
    """float64 (H, W): float32 class-balance weight + w0 (see the module docstring). ``sigma`` only
    enters through exp(-0 / (2 (sigma^2 + 1e-8))) = 1."""
    fg = np.asarray(labels) > 0
    wc_bg, wc_fg = class_balance_weights(int(fg.sum()), fg.size)
    border = float(w0) * float(np.exp(-0.0 / (2.0 * (float(sigma) ** 2 + 1e-8))))
    return np.where(fg, np.float64(wc_fg), np.float64(wc_bg)) + border


def weight_map_full(labels: np.ndarray, w0=10, sigma=5) -> np.ndarray:
    """Step-by-step restatement (preprocess_data.py:22-77) with real distance transforms.
    Returns float64, except for a mask without any instance, where the reference's d-maps are
    float32 and so is its result (:62-64)."""
    from scipy.ndimage import distance_transform_edt as edt

    labels = np.asarray(labels)
    fg = labels > 0
    wc_bg, wc_fg = class_balance_weights(int(fg.sum()), fg.size)
    wc = np.where(fg, wc_fg, wc_bg).astype(np.float32)

    ids = np.unique(labels[fg])
    if ids.size == 0:
        d1 = np.zeros(labels.shape, np.float32)
        d2 = np.zeros(labels.shape, np.float32)
    else:
        per_instance = np.stack(
            [np.minimum(edt(labels == i), edt(labels != i)) for i in ids], axis=-1)
        if ids.size >= 2:
            two_nearest = np.partition(per_instance, 1, axis=-1)
            d1, d2 = two_nearest[..., 0], two_nearest[..., 1]
        else:
            d1 = per_instance[..., 0]
            d2 = np.zeros_like(d1)
    d1 = np.where(np.isinf(d1), 0, d1).astype(d1.dtype)
    d2 = np.where(np.isinf(d2), 0, d2).astype(d2.dtype)
    border = w0 * np.exp(-((d1 + d2) ** 2) / (2 * (sigma ** 2 + 1e-8)))
    return wc + border
