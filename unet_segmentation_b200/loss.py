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


class _TargetErrors:
    """Out-of-range targets, surfaced without stalling the training loop.

    The reference's ``nn.CrossEntropyLoss`` faults on a target outside ``[0, C)`` (a device-side
    assert on CUDA, which the host sees at its next synchronisation). The kernel raises a device
    flag instead (and makes the loss NaN); the flag of every call is copied to pinned host memory
    behind the kernel, and a later call — or ``check_targets=True``, which waits — raises as soon
    as a finished copy shows it. Nothing here blocks the stream in the default mode.
    """

    def __init__(self):
        self.pending = []     # (event, pinned int32 flag)
        self.free = []

    def __deepcopy__(self, memo):     # CUDA events / pinned flags are per-instance bookkeeping
        return _TargetErrors()

    def __reduce__(self):
        return (_TargetErrors, ())

    def poll(self, wait: bool = False) -> None:
        keep = []
        bad = False
        for ev, flag in self.pending:
            if wait:
                ev.synchronize()
            if wait or ev.query():
                bad = bad or int(flag.item()) != 0
                self.free.append(flag)
            else:
                keep.append((ev, flag))
        self.pending = keep
        if bad:
            raise RuntimeError("WeightedCrossEntropyLoss (B200): a target value lies outside "
                               "[0, n_classes) (and is not ignore_index = -100); the reference's "
                               "nn.CrossEntropyLoss faults on such input (un-binarised mask?). The "
                               "loss of that call is NaN.")

    def watch(self, err: torch.Tensor) -> None:
        flag = self.free.pop() if self.free else torch.zeros(1, dtype=torch.int32).pin_memory()
        flag.copy_(err, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(err.device))
        self.pending.append((ev, flag))
        if len(self.pending) > 64:      # bound the list when nobody ever synchronises
            self.poll(wait=True)


class _WceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, weight_maps, errors):
        need_grad = logits.requires_grad
        loss, dz, err = ops.wce_forward(logits.detach(), targets, weight_maps, want_grad=need_grad)
        ctx.dz = dz
        if errors is not None:
            errors.watch(err)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.dz is None:
            return None, None, None, None
        g = grad_out.contiguous().float()
        return ops.scale_by_device_scalar(ctx.dz, g), None, None, None


class WeightedCrossEntropyLoss(nn.Module):
    """Pixel-wise weighted cross-entropy (U-Net paper loss), mean over N*H*W.

    ``check_targets`` (default False): True makes every call wait for its own result and raise at
    once on an out-of-range target (one device->host synchronisation per call); False raises from a
    later call, once the flag has arrived, and the loss of the offending call is NaN either way.
    """

    def __init__(self):
        super().__init__()
        # kept for attribute parity with the reference (utils/losses.py:27); not used for compute
        self.cross_entropy = nn.CrossEntropyLoss(reduction="none")
        self.check_targets = False
        self._errors = _TargetErrors()

    def forward(self, inputs, targets, weight_maps):
        if not inputs.is_cuda:
            raise RuntimeError("WeightedCrossEntropyLoss (B200) runs only on CUDA tensors; there is "
                               "no CPU fallback")
        with torch.cuda.device(inputs.device):
            # no host-side bookkeeping inside a CUDA-graph capture (events / pinned copies would
            # become graph nodes); the NaN loss still marks the bad step there
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                self._errors.poll()
            loss = _WceFunction.apply(inputs, targets, weight_maps,
                                      None if capturing else self._errors)
            if self.check_targets and not capturing:
                self._errors.poll(wait=True)
        return loss
