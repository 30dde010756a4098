"""CPU restatement of the reference's ``get_instance_masks`` (TEST INFRASTRUCTURE ONLY).

Reference: utils/metrics.py:62-72 — ``skimage.measure.label(mask > 0, connectivity=2)`` then
``skimage.morphology.remove_small_objects(labels, min_size)`` then ``astype(uint16)``.
scikit-image (un-pinned in requirements.txt:9) is not installed in this image, so its published
algorithm is restated: ``label`` numbers 8-connected components 1..K in raster order of their first
pixel (scipy.ndimage.label with a full 3x3 structure does exactly that), ``remove_small_objects``
zeroes components whose pixel count is < min_size and leaves the other ids untouched.
Pinned bit-exactly on the reference's own shipped mask -> instance pairs (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np


def get_instance_masks(binary_mask: np.ndarray, min_size: int = 15) -> np.ndarray:
    from scipy import ndimage

    fg = np.asarray(binary_mask) > 0                       # utils/metrics.py:62
    labels, _ = ndimage.label(fg, structure=np.ones((3, 3), dtype=bool))   # :65
    areas = np.bincount(labels.ravel())
    too_small = areas < min_size                           # :69 (remove_small_objects)
    too_small[0] = False
    labels = labels.copy()
    labels[too_small[labels]] = 0
    return labels.astype(np.uint16)                        # :72


def get_instance_masks_pure(binary_mask: np.ndarray, min_size: int = 15) -> np.ndarray:
    """Dependency-free two-pass union-find version (small cases; cross-checks the scipy one)."""
    fg = np.asarray(binary_mask) > 0
    h, w = fg.shape
    parent = {}

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for y in range(h):
        for x in range(w):
            if not fg[y, x]:
                continue
            i = y * w + x
            parent[i] = i
            for dy, dx in ((0, -1), (-1, -1), (-1, 0), (-1, 1)):
                yy, xx = y + dy, x + dx
                if 0 <= yy < h and 0 <= xx < w and fg[yy, xx]:
                    ra, rb = find(i), find(yy * w + xx)
                    if ra != rb:
                        parent[max(ra, rb)] = min(ra, rb)
    out = np.zeros((h, w), dtype=np.int64)
    ids, areas = {}, {}
    for y in range(h):
        for x in range(w):
            if fg[y, x]:
                r = find(y * w + x)
                if r not in ids:
                    ids[r] = len(ids) + 1
                areas[r] = areas.get(r, 0) + 1
                out[y, x] = r + 1
    res = np.zeros((h, w), dtype=np.uint16)
    for y in range(h):
        for x in range(w):
            if out[y, x]:
                r = out[y, x] - 1
                if areas[r] >= min_size:
                    res[y, x] = ids[r] & 0xFFFF
    return res
