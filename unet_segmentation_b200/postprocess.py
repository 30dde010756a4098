"""Prediction post-processing on the GPU (reference scripts/predict.py:85-98).

``get_instance_masks`` keeps the reference signature (numpy in, numpy uint16 out) and is
bit-exact with ``skimage.measure.label(connectivity=2)`` + ``remove_small_objects`` + uint16 cast
(utils/metrics.py:62-72); it also accepts a CUDA tensor and then returns a CUDA tensor, avoiding
the ``.cpu().numpy()`` hop of the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def get_instance_masks(binary_mask, min_size: int = 15):
    if isinstance(binary_mask, torch.Tensor):
        if not binary_mask.is_cuda:
            raise RuntimeError("get_instance_masks (B200): tensor input must live on a CUDA device")
        return ops.ccl_label(binary_mask, int(min_size))
    arr = np.asarray(binary_mask)
    if arr.ndim != 2:
        raise ValueError("expected a 2-D mask")
    if not torch.cuda.is_available():
        raise RuntimeError("get_instance_masks (B200) needs a CUDA device; there is no CPU fallback")
    t = torch.from_numpy(np.ascontiguousarray((arr > 0).astype(np.uint8))).cuda()
    return ops.ccl_label(t, int(min_size)).cpu().numpy()


def instance_labels_from_logits(logits: torch.Tensor, min_size: int = 15) -> torch.Tensor:
    """(1|N, 2, H, W) logits -> per-image uint16 labels: softmax[:,1] > 0.5 == logit1 > logit0."""
    masks = (logits[:, 1] > logits[:, 0]).to(torch.uint8) * 255
    return torch.stack([ops.ccl_label(m, min_size) for m in masks])
