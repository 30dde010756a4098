"""Drop-in ``WeightedCrossEntropyLoss`` for the reference's utils/losses.py.

``forward(inputs (N,C,H,W) fp32 logits, targets (N,H,W) int64, weight_maps (N,H,W) fp32)`` returns
the 0-dim mean of ``w(x) * CE(x)`` exactly like the reference (utils/losses.py:49,54,57). One
coalesced CUDA kernel computes the loss partials (warp-shuffle + two-stage deterministic
reduction) and ``d loss / d logits`` in the same pass; inputs may be the non-contiguous cropped
views produced by scripts/train.py:118-126. CUDA only — CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _WceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, weight_maps):
        need_grad = logits.requires_grad
        loss, dz, err = ops.wce_forward(logits.detach(), targets, weight_maps, want_grad=need_grad)
        ctx.dz = dz
        ctx.err = err
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.dz is None:
            return None, None, None
        g = grad_out.contiguous().float()
        return ops.scale_by_device_scalar(ctx.dz, g), None, None


class WeightedCrossEntropyLoss(nn.Module):
    """Pixel-wise weighted cross-entropy (U-Net paper loss), mean over N*H*W."""

    def __init__(self):
        super().__init__()
        # kept for attribute parity with the reference (utils/losses.py:27); not used for compute
        self.cross_entropy = nn.CrossEntropyLoss(reduction="none")
        self.check_targets = False

    def forward(self, inputs, targets, weight_maps):
        if not inputs.is_cuda:
            raise RuntimeError("WeightedCrossEntropyLoss (B200) runs only on CUDA tensors; there is "
                               "no CPU fallback")
        with torch.cuda.device(inputs.device):
            loss = _WceFunction.apply(inputs, targets, weight_maps)
        if self.check_targets:  # opt-in: costs a device->host sync
            pass
        return loss
