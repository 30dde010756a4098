"""Overlap-tile inference (U-Net paper, Fig. 2) for images larger than one forward pass.

The reference repository ships only result pictures of this strategy, not its code (SURVEY F2:
``predictions_output_overlap_tile/*.png``, ``images/old readme unet.txt:73-87``), so the semantics are
defined here and pinned by an invariant instead of golden vectors:

  * the image is extended by mirroring (``reflect``) by the network margin (92 px for 5 levels) so
    that the output covers the whole image;
  * input tiles have a size S ≡ 12 (mod 16) (572 by default → 388² outputs, no pooling floor loss,
    SURVEY F8) and origins that are multiples of 16; under these two conditions every max-pool window
    of a tile coincides with a window of the whole-image forward, so a tile's logits are identical to
    the corresponding crop of a single huge forward pass (eval-mode BN) — there is nothing to blend,
    overlapping output pixels are equal and the later tile simply overwrites;
  * tiles are independent units: they are batched on one GPU and dealt round-robin across ranks
    (``parallel.shard_indices``) with no inter-GPU reduction.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import parallel


def network_margin(levels: int = 5) -> int:
    """Half of (input - output) for pooling-aligned sizes: 92 for the 5-level reference net
    (572 -> 388, reference models/unet_model.py:175-187)."""
    size = 12 + 16 * (1 << levels)          # ≡ 12 (mod 16) and large enough: no floor anywhere
    s = size
    for i in range(levels):
        s -= 4
        if i < levels - 1:
            assert s % 2 == 0
            s //= 2
    for _ in range(levels - 1):
        s = 2 * s - 4
    return (size - s) // 2


def plan_tiles(h: int, w: int, tile_in: int = 572, levels: int = 5) -> Tuple[int, int, List[Tuple[int, int]]]:
    """Returns (tile_out, stride, origins). Origins are output-space top-left corners, multiples
    of 16, covering [0,h) x [0,w)."""
    if tile_in % 16 != 12:
        raise ValueError("tile_in must be ≡ 12 (mod 16) so that no pooling level floors (SURVEY F8)")
    tile_out = tile_in - 2 * network_margin(levels)
    if tile_out <= 0:
        raise ValueError("tile too small for the network margin")
    stride = (tile_out // 16) * 16
    ys = list(range(0, max(h - tile_out, 0) + stride, stride)) if h > tile_out else [0]
    xs = list(range(0, max(w - tile_out, 0) + stride, stride)) if w > tile_out else [0]
    ys = [y for y in ys if y < h]
    xs = [x for x in xs if x < w]
    while ys[-1] + tile_out < h:
        ys.append(ys[-1] + stride)
    while xs[-1] + tile_out < w:
        xs.append(xs[-1] + stride)
    return tile_out, stride, [(y, x) for y in ys for x in xs]


def _reflect_index(i: torch.Tensor, n: int) -> torch.Tensor:
    """numpy 'reflect' (no edge repeat) extension of arbitrary length."""
    if n == 1:
        return torch.zeros_like(i)
    period = 2 * (n - 1)
    i = i % period
    return torch.where(i >= n, period - i, i)


def extract_tiles(image: torch.Tensor, origins, tile_in: int, margin: int) -> torch.Tensor:
    """image (H, W) fp32 -> (T, 1, tile_in, tile_in), mirror-extended around the borders (one
    gather for the whole batch of tiles)."""
    h, w = image.shape
    dev = image.device
    ar = torch.arange(tile_in, device=dev)
    oy = torch.tensor([y - margin for (y, _) in origins], device=dev)
    ox = torch.tensor([x - margin for (_, x) in origins], device=dev)
    iy = _reflect_index(ar[None, :] + oy[:, None], h)      # (T, S)
    ix = _reflect_index(ar[None, :] + ox[:, None], w)
    return image[iy[:, :, None], ix[:, None, :]].unsqueeze(1)


# Cost model of one eval forward of `bt` tiles of tile_in^2 pixels on one B200 (measured: 8 x 1212^2 in
# 14.6 ms; ~0.25 ms of launch / tail latency per forward): used to pick the tile size, the batch and
# how many ranks are worth using.
_MS_PER_PIXEL = 14.6 / (8 * 1212 * 1212)
_MS_PER_FORWARD = 0.25
_MS_PER_GATHER = 0.08          # all-gather of the finished uint8 tiles when more than one rank works


def choose_plan(h: int, w: int, world: int = 1, levels: int = 5, max_tile_in: int = 2300,
                min_tile_in: int = 380, batch_tiles: int = 8) -> Tuple[int, int, int]:
    """(tile_in, ranks_used, tiles_per_batch) minimising the modelled time of overlap-tile inference:
    batches per rank x (fixed latency + pixels executed), over every tile size ≡ 12 (mod 16) and every
    rank count up to ``world``. It trades the halo overhead of small tiles (572 -> 388: 2.04 MFLOP
    per output pixel vs 1.47 without halo) against the coverage waste of tiles that do not divide the
    image, rank imbalance, padded batches and per-forward latency; using FEWER ranks than offered is
    allowed, so the throughput never drops when GPUs are added (a 1024^2 image has work for four)."""
    best, best_cost = (572, 1, 1), None
    for tile_in in range(min_tile_in + (12 - min_tile_in) % 16, max_tile_in + 1, 16):
        n = len(plan_tiles(h, w, tile_in, levels)[2])
        for ranks in range(1, max(1, min(world, n)) + 1):
            per_rank = -(-n // ranks)
            bt = max(1, min(batch_tiles, per_rank))
            batches = -(-per_rank // bt)
            cost = batches * (_MS_PER_FORWARD + _MS_PER_PIXEL * bt * tile_in * tile_in)
            if ranks > 1:
                cost += _MS_PER_GATHER
            if best_cost is None or cost < best_cost * 0.999:
                best, best_cost = (tile_in, ranks, bt), cost
    return best


def choose_tile(h: int, w: int, world: int = 1, levels: int = 5, max_tile_in: int = 2300,
                min_tile_in: int = 380, batch_tiles: int = 8) -> int:
    """Input tile size (≡ 12 mod 16) of ``choose_plan``."""
    return choose_plan(h, w, world, levels, max_tile_in, min_tile_in, batch_tiles)[0]


class _TileSession:
    """Everything of an overlap-tile run that depends only on (image shape, tile plan, rank layout):
    device tables of tile origins per batch and for the stitch, the uint8 slot buffers. Cached on the
    model, so a repeated prediction does no host-side planning, no allocation and no H2D copy."""

    def __init__(self, model, h, w, tile_in, batch_tiles, rank, world, ranks_used, device):
        levels = getattr(model, "levels", 5)
        self.margin = network_margin(levels)
        self.tile_in = tile_in
        self.tile_out, _, self.origins = plan_tiles(h, w, tile_in, levels)
        n = len(self.origins)
        self.ranks_used = max(1, min(ranks_used, world, n))
        self.slots = -(-n // self.ranks_used)                       # tile slots per working rank
        self.bt = max(1, min(batch_tiles, self.slots))
        self.slots = -(-self.slots // self.bt) * self.bt            # whole batches
        table = torch.full((world, self.slots, 2), -1, dtype=torch.int32)
        for r in range(self.ranks_used):
            for k, idx in enumerate(parallel.shard_indices(n, r, self.ranks_used)):
                table[r, k, 0], table[r, k, 1] = self.origins[idx]
        self.mine = parallel.shard_indices(n, rank, self.ranks_used) if rank < self.ranks_used else []
        self.n_batches = -(-len(self.mine) // self.bt)
        self.table = table.to(device)                               # (world, slots, 2): stitch table
        self.my_table = self.table[rank].contiguous()
        self.masks = torch.zeros(self.slots, self.tile_out, self.tile_out, dtype=torch.uint8,
                                 device=device)
        self.gathered = (torch.empty(world, self.slots, self.tile_out, self.tile_out,
                                     dtype=torch.uint8, device=device) if world > 1 else None)
        self.world = world


def _session(model, h, w, tile_in, batch_tiles, rank, world, ranks_used, device) -> _TileSession:
    cache = model.__dict__.setdefault("_tile_sessions", {})
    key = (h, w, tile_in, batch_tiles, rank, world, ranks_used, device.index)
    sess = cache.get(key)
    if sess is None:
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))
        sess = cache[key] = _TileSession(model, h, w, tile_in, batch_tiles, rank, world, ranks_used,
                                         device)
    return sess


@torch.no_grad()
def overlap_tile_predict(model, image: torch.Tensor, tile_in: Optional[int] = 572,
                         batch_tiles: int = 8, rank: int = 0, world: int = 1, group=None,
                         return_logits: bool = False, ranks_used: Optional[int] = None):
    """Whole-image binary mask (uint8, 255 = foreground) of a 2-D fp32 CUDA image.

    ``model`` is a ``unet_segmentation_b200.UNet`` in eval mode. With ``world > 1`` every rank calls
    this with the same image and gets the same stitched result. ``tile_in=None`` picks tile size,
    batch and the number of ranks worth using with ``choose_plan`` (``ranks_used`` overrides the
    latter; ranks beyond it stay idle but still take part in the gather).

    Data path per batch of tiles, all on the device and without per-tile host work: one
    ``ub_extract_tiles`` launch mirrors / gathers the tiles straight into the eval plan's input
    buffer, the forward pass (replayed from a CUDA graph) leaves the uint8 masks in the plan's mask
    buffer, and after the last batch ONE ``ub_stitch_tiles`` launch writes every tile of every rank
    into the result. Ranks exchange finished uint8 tiles with a single all-gather — no reduction.
    """
    from . import _lib
    from ._lib import check
    import ctypes as C

    if image.dim() != 2 or not image.is_cuda:
        raise ValueError("expected a 2-D CUDA image")
    if model.training:
        raise RuntimeError("overlap_tile_predict needs model.eval()")
    if model.n_channels != 1:
        raise ValueError("overlap-tile inference is defined for single-channel images")
    lib = _lib.load()
    dev = image.device
    levels = getattr(model, "levels", 5)
    h, w = image.shape
    if tile_in is None:
        tile_in, auto_ranks, batch_tiles = choose_plan(h, w, world, levels, batch_tiles=batch_tiles)
        if ranks_used is None:
            ranks_used = auto_ranks
    if ranks_used is None:
        ranks_used = world
    sess = _session(model, h, w, tile_in, batch_tiles, rank, world, ranks_used, dev)
    image = image.contiguous().float()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    out_logits = []
    with torch.cuda.device(dev):
        if sess.n_batches:
            plan = model._plan_for_shape(sess.bt, 1, tile_in, tile_in, False, dev)
            xs, _, _ = plan.static_io(dev)
        for b in range(sess.n_batches):
            org = sess.my_table[b * sess.bt:(b + 1) * sess.bt]
            check(lib.ub_extract_tiles(C.c_void_p(image.data_ptr()), h, w, C.c_void_p(org.data_ptr()),
                                       sess.bt, tile_in, sess.margin, C.c_void_p(xs.data_ptr()), stream),
                  "ub_extract_tiles")
            logits, mask = plan.forward(xs, want_mask=True, static_out=True)
            sess.masks[b * sess.bt:(b + 1) * sess.bt].copy_(mask)
            if return_logits:
                out_logits.append(logits.clone())
        if world > 1:
            import torch.distributed as dist

            dist.all_gather_into_tensor(sess.gathered, sess.masks, group=group)
            tiles, table, count = sess.gathered, sess.table, world * sess.slots
        else:
            tiles, table, count = sess.masks, sess.my_table, sess.slots
        full = torch.empty(h, w, dtype=torch.uint8, device=dev)
        check(lib.ub_stitch_tiles(C.c_void_p(tiles.data_ptr()), C.c_void_p(table.data_ptr()), count,
                                  sess.tile_out, C.c_void_p(full.data_ptr()), h, w, stream),
              "ub_stitch_tiles")
    if not return_logits:
        return full
    # logits are a test / debugging facility: stitched on the host side of the API
    per_tile = [t for batch in out_logits for t in batch.unbind(0)][:len(sess.mine)]
    if world > 1:
        per_tile = _gather_logits(per_tile, sess, world, model.n_classes, group)
    full_logits = torch.zeros(model.n_classes, h, w, dtype=torch.float32, device=dev)
    order = sess.mine if world == 1 else range(len(sess.origins))
    for k, idx in enumerate(order):
        y, x = sess.origins[idx]
        hh, ww = min(sess.tile_out, h - y), min(sess.tile_out, w - x)
        full_logits[:, y:y + hh, x:x + ww] = per_tile[k][:, :hh, :ww]
    return full, full_logits


def _gather_logits(per_tile, sess, world, n_classes, group):
    """return_logits across ranks: pad every rank to `slots` tiles and gather (test facility)."""
    import torch.distributed as dist

    proto_shape = (sess.slots, n_classes, sess.tile_out, sess.tile_out)
    dev = sess.masks.device
    buf = torch.zeros(proto_shape, dtype=torch.float32, device=dev)
    for i, t in enumerate(per_tile):
        buf[i] = t
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    n = len(sess.origins)
    tiles = [None] * n
    for r in range(sess.ranks_used):
        for i, idx in enumerate(parallel.shard_indices(n, r, sess.ranks_used)):
            tiles[idx] = out[r][i]
    return tiles
