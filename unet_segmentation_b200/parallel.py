"""Data-parallel training and tile-sharded inference helpers (one process per GPU).

Training (BASELINE config 3): every rank holds a full fp32 replica and runs its own shard of the
batch; gradients are averaged with one NCCL all-reduce per *backward stage* (``ub_plan_backward_stage``:
outc+up4, up3, up2, up1, down4 ... inc), launched on a side stream as soon as the kernels of the
stage are enqueued, so the exchange over NVLink overlaps the remaining backward kernels
(SURVEY F10 / §8e). Batch-norm statistics stay local to each rank (the semantics of stock
``DistributedDataParallel`` around the reference); ``sync_bn_buffers`` broadcasts rank 0's running
statistics when a single checkpoint is wanted.

Inference (config 4): overlap-tile units are independent, so tiles are dealt round-robin to the
ranks and there is no collective on the data path; only the finished uint8 tiles are gathered.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent work units (tiles) to ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every replica start from rank ``src``'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


def sync_bn_buffers(module: torch.nn.Module, src: int = 0, group=None) -> None:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for b in module.buffers():
            dist.broadcast(b, src=src, group=group)


class StageGradAllReducer:
    """Averages gradients across ranks, one collective per backward stage, overlapped with backward.

    Attach to a ``unet_segmentation_b200.unet.UNet``::

        reducer = StageGradAllReducer(model)      # after dist.init_process_group
        loss.backward()                           # all-reduces are issued from inside backward
        optimizer.step()

    Works with the ``nccl`` backend on CUDA tensors (side stream + events) and with ``gloo`` on CPU
    tensors (synchronous; used by the CPU tests of the host logic).
    """

    def __init__(self, model: Optional[torch.nn.Module] = None, group=None,
                 wire_dtype: Optional[torch.dtype] = None):
        """``wire_dtype=torch.bfloat16`` sends the gradients over NVLink as bf16 (62 MB instead of
        124 MB per step for the reference net). The averaged gradient is rounded to bf16 once more than
        with the default fp32 payload (masters, momentum and the update stay fp32). Measured on an
        8 x B200 NVSwitch box it is SLOWER than fp32 (14.72-14.75 vs 14.56-14.62 ms per step,
        profiles/r02_grad_wire_ab.txt): the pack / unpack kernels cost more than the shorter
        collective saves. Kept for links slower than NVLink 5; fp32 is the default."""
        self.group = group
        self.wire_dtype = wire_dtype
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.comm_stream = None
        self.n_collectives = 0
        self.bytes_reduced = 0
        self.model = model
        if model is not None:
            model.set_backward_hooks(self.on_stage, self.on_done)

    # The library computes weight gradients on its own side stream; this hook orders the
    # communication stream after it itself (plan.join_side), so the backward stream never stalls.
    @property
    def joins_side_stream(self) -> bool:
        return self.model is not None and hasattr(self.model, "_latest_training_plan")

    def on_stage(self, stage: int, flat_grads: torch.Tensor) -> None:
        if self.world == 1:
            return
        self.n_collectives += 1
        self.bytes_reduced += flat_grads.numel() * flat_grads.element_size()
        if flat_grads.is_cuda:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(device=flat_grads.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(flat_grads.device))
            self.comm_stream.wait_event(ready)
            if self.joins_side_stream:
                plan = self.model._latest_training_plan()
                if plan is not None:
                    plan.join_side(self.comm_stream)
            with torch.cuda.stream(self.comm_stream):
                if self.wire_dtype is not None and self.wire_dtype != flat_grads.dtype:
                    wire = flat_grads.to(self.wire_dtype)
                    dist.all_reduce(wire, op=dist.ReduceOp.AVG, group=self.group)
                    flat_grads.copy_(wire)
                else:
                    dist.all_reduce(flat_grads, op=dist.ReduceOp.AVG, group=self.group)
            flat_grads.record_stream(self.comm_stream)
        elif self.wire_dtype is not None and self.wire_dtype != flat_grads.dtype:
            wire = flat_grads.to(self.wire_dtype)          # same rounding points as the CUDA path
            dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group)
            flat_grads.copy_(wire.to(flat_grads.dtype).div_(self.world).to(self.wire_dtype))
        else:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)
            flat_grads.div_(self.world)

    def on_done(self) -> None:
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


def gather_tiles(local: Sequence[torch.Tensor], n_total: int, rank: int, world: int,
                 group=None) -> List[torch.Tensor]:
    """Collect equally-shaped per-rank tile results (dealt by ``shard_indices``) on every rank, in
    tile order. No reduction: the payload is the disjoint set of finished tiles."""
    if world == 1:
        return list(local)
    per_rank = (n_total + world - 1) // world
    proto = local[0] if len(local) else None
    if proto is None:
        raise ValueError("every rank needs at least one tile (world <= number of tiles)")
    buf = torch.zeros((per_rank,) + tuple(proto.shape), dtype=proto.dtype, device=proto.device)
    for i, t in enumerate(local):
        buf[i] = t
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    tiles: List[Optional[torch.Tensor]] = [None] * n_total
    for r in range(world):
        for i, idx in enumerate(shard_indices(n_total, r, world)):
            tiles[idx] = out[r][i]
    return tiles  # type: ignore[return-value]
