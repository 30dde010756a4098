"""Per-frame prediction loop of the reference (scripts/predict.py:73-112) as ONE CUDA-graph launch.

The reference handles one 512 x 512 frame at a time: ``ToTensor`` + ``Normalize(.5, .5)`` -> eval forward
-> ``softmax(dim=1)[:, 1] > 0.5`` -> ``get_instance_masks(mask, min_size=15)`` -> uint8 mask + uint16
instance labels back on the host. On a B200 the arithmetic of that is ~0.2 ms, so the loop is bound by
host work (kernel launches, tensor bookkeeping, the ``.cpu().numpy()`` hop). ``FramePredictor`` records
the whole device side once — host-to-device copy of the frame from pinned memory, the eval forward
with the head and the mask fused into the last conv, connected components + small-object filter,
both device-to-host copies — and replays it per frame.

Reference semantics kept: the input is whatever the caller's transform produced (float32, (1, C, H, W)
or (C, H, W) / (H, W)), the mask is 255 where ``softmax[:, 1] > 0.5`` (``logit1 > logit0``; a single-class
model thresholds ``sigmoid > 0.5``), labels follow ``skimage.measure.label(connectivity=2)`` +
``remove_small_objects`` + ``astype(uint16)`` bit for bit (utils/metrics.py:62-72).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .postprocess import get_instance_masks


class FramePredictor:
    """``mask_u8, labels_u16 = FramePredictor(model, (H, W))(frame)`` — numpy in, numpy out (views of
    pinned buffers that the next call overwrites; copy them to keep them).

    ``batch`` > 1 handles that many frames per launch (input (batch, C, H, W), outputs (batch, h, w)):
    one frame alone cannot fill 148 SMs — its bottleneck layers have 5 ... 17 tiles with K loops of
    up to 9 216 — so frames of a sequence are best predicted a few at a time."""

    def __init__(self, model, frame_hw: Tuple[int, int], min_size: int = 15, device=None,
                 batch: int = 1):
        if model.training:
            raise RuntimeError("FramePredictor needs model.eval() (scripts/predict.py:70)")
        p = model.outc.conv.weight
        if not p.is_cuda:
            raise RuntimeError("FramePredictor (B200) needs the model on a CUDA device; there is no "
                               "CPU fallback")
        self.model, self.min_size = model, int(min_size)
        self.device = torch.device(device) if device is not None else p.device
        h, w = frame_hw
        c = model.n_channels
        self.batch = int(batch)
        self.frame_h = torch.empty(self.batch, c, h, w, dtype=torch.float32).pin_memory()
        self.x = torch.empty(self.batch, c, h, w, dtype=torch.float32, device=self.device)
        self.graph = None
        self._weights_key = None
        self._pending = False
        self._record()

    def _key(self):
        m = self.model
        return (m._weights_epoch, m.outc.conv.weight._version, m.inc.double_conv[0].weight._version)

    def _pipeline(self):
        self.x.copy_(self.frame_h, non_blocking=True)
        _, mask = self.model.predict_mask(self.x)
        if self.batch == 1:
            return mask[0], get_instance_masks(mask[0], min_size=self.min_size)
        return mask, torch.stack([get_instance_masks(m, min_size=self.min_size) for m in mask])

    def _record(self):
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):            # warm-up: plans, packed weights, allocator pools
                for _ in range(2):
                    mask, labels = self._pipeline()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.mask_h = torch.empty(mask.shape, dtype=torch.uint8).pin_memory()
            self.labels_h = torch.empty(labels.shape, dtype=torch.uint16).pin_memory()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                mask, labels = self._pipeline()
                self.mask_h.copy_(mask, non_blocking=True)
                self.labels_h.copy_(labels, non_blocking=True)
            self._weights_key = self._key()

    def refresh(self) -> None:
        """Re-record after the model's weights changed (the packed bf16 operands are refreshed
        outside the graph)."""
        self._record()

    def submit(self, frame) -> None:
        """Enqueue one frame — or ``batch`` frames — (numpy or CPU tensor, any shape that reshapes to
        (batch, C, H, W))."""
        if self._pending:         # the previous replay may still be reading the pinned input
            torch.cuda.current_stream(self.device).synchronize()
            self._pending = False
        if self._key() != self._weights_key:
            self._record()
        src = frame if isinstance(frame, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frame))
        self.frame_h.copy_(src.reshape(self.frame_h.shape))
        with torch.cuda.device(self.device):
            self.graph.replay()
        self._pending = True

    def result(self) -> Tuple[np.ndarray, np.ndarray]:
        torch.cuda.current_stream(self.device).synchronize()
        self._pending = False
        return self.mask_h.numpy(), self.labels_h.numpy()

    def __call__(self, frame) -> Tuple[np.ndarray, np.ndarray]:
        self.submit(frame)
        return self.result()
