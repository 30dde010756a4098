"""Host-side logic of the multi-GPU paths, exercised with world_size-2 gloo process groups on CPU:
stage-wise gradient averaging (data-parallel training) and round-robin tile sharding + gather
(overlap-tile inference). No CUDA needed."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unet_segmentation_b200 import parallel, tiling


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _stage_reduce(rank, world):
    red = parallel.StageGradAllReducer(model=None)
    flat = torch.arange(40, dtype=torch.float32) * (rank + 1)
    # three "stages" covering disjoint contiguous slices, like ub_plan_stage_params ranges
    for s, (a, b) in enumerate([(25, 40), (10, 25), (0, 10)]):
        red.on_stage(s, flat[a:b])
    red.on_done()
    return flat.tolist(), red.n_collectives, red.bytes_reduced


def test_stage_grad_allreducer_averages_across_ranks():
    out = _run(_stage_reduce)
    expect = (torch.arange(40, dtype=torch.float32) * 1.5).tolist()   # mean of x1 and x2
    for flat, n_coll, nbytes in out:
        assert flat == expect
        assert n_coll == 3 and nbytes == 40 * 4


def _stage_reduce_bf16_wire(rank, world):
    red = parallel.StageGradAllReducer(model=None, wire_dtype=torch.bfloat16)
    flat = (torch.arange(40, dtype=torch.float32) + 0.123) * (rank + 1)
    for s, (a, b) in enumerate([(25, 40), (0, 25)]):
        red.on_stage(s, flat[a:b])
    red.on_done()
    return flat.tolist(), flat.dtype == torch.float32


def test_stage_grad_allreducer_bf16_payload_keeps_fp32_masters():
    """wire_dtype=bf16: each rank's slice is rounded to bf16, summed, averaged and rounded once more;
    the caller's gradient buffer stays fp32 and every rank ends with identical values."""
    out = _run(_stage_reduce_bf16_wire)
    base = torch.arange(40, dtype=torch.float32) + 0.123
    w = base.bfloat16() + (base * 2).bfloat16()                 # bf16 accumulation on the wire
    expect = (w.float() / 2).bfloat16().float()
    for flat, is_f32 in out:
        assert is_f32
        got = torch.tensor(flat)
        assert torch.equal(got, expect)
        assert float((got - base * 1.5).abs().max() / (base * 1.5).abs().max()) < 1e-2


def _tile_gather(rank, world):
    n_tiles = 5
    mine = parallel.shard_indices(n_tiles, rank, world)
    local = [torch.full((3, 4), float(i), dtype=torch.float32) for i in mine]
    tiles = parallel.gather_tiles(local, n_tiles, rank, world)
    return [float(t[0, 0]) for t in tiles], mine


def test_tile_sharding_and_gather_cover_every_tile_once():
    out = _run(_tile_gather)
    assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]
    for vals, _ in out:
        assert vals == [0.0, 1.0, 2.0, 3.0, 4.0]


def test_shard_indices_partition():
    for n, world in [(9, 8), (484, 8), (4, 4), (1, 1), (0, 2)]:
        seen = sorted(i for r in range(world) for i in parallel.shard_indices(n, r, world))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        parallel.shard_indices(4, 2, 2)


def test_tile_plan_geometry():
    assert tiling.network_margin(5) == 92
    for (h, w), n in [((1024, 1024), 9), ((8192, 8192), 484), ((324, 324), 1), ((388, 388), 1),
                      ((389, 388), 2)]:
        tout, stride, origins = tiling.plan_tiles(h, w, 572)
        assert tout == 388 and stride == 384 and len(origins) == n
        assert all(y % 16 == 0 and x % 16 == 0 for y, x in origins)
        cover = torch.zeros(h, w, dtype=torch.bool)
        for y, x in origins:
            cover[y:y + tout, x:x + tout] = True
        assert bool(cover.all())
    with pytest.raises(ValueError):
        tiling.plan_tiles(1024, 1024, 512)     # 512 ≢ 12 (mod 16): pooling would floor


def test_reflect_extension_matches_numpy():
    import numpy as np

    img = torch.arange(7 * 5, dtype=torch.float32).reshape(7, 5)
    tiles = tiling.extract_tiles(img, [(0, 0)], tile_in=7 + 2 * 9, margin=9)
    ref = np.pad(img.numpy(), ((9, 9), (9, 11)), mode="reflect")[:25, :25]
    assert np.array_equal(tiles[0, 0].numpy(), ref)


def test_choose_plan_minimises_modelled_time_and_never_loses_with_more_ranks():
    """tiling.choose_plan: aligned tile sizes (≡ 12 mod 16), full coverage, modelled time no worse
    than the 572-tile default and monotone in the number of ranks offered (it may leave ranks idle:
    a 1024^2 image has work for four GPUs, not eight)."""
    from unet_segmentation_b200 import tiling

    def cost(size, tile_in, ranks, bt):
        n = len(tiling.plan_tiles(size, size, tile_in)[2])
        per_rank = -(-n // ranks)
        bt = max(1, min(bt, per_rank))
        c = -(-per_rank // bt) * (tiling._MS_PER_FORWARD + tiling._MS_PER_PIXEL * bt * tile_in * tile_in)
        return c + (tiling._MS_PER_GATHER if ranks > 1 else 0.0)

    for size in (600, 1024, 2048, 8192):
        prev = None
        for world in (1, 2, 4, 8):
            t, ranks, bt = tiling.choose_plan(size, size, world)
            assert t == tiling.choose_tile(size, size, world)
            assert t % 16 == 12 and 1 <= ranks <= world and 1 <= bt <= 8
            tile_out, stride, origins = tiling.plan_tiles(size, size, t)
            assert stride % 16 == 0 and all(y % 16 == 0 and x % 16 == 0 for y, x in origins)
            assert max(y for y, _ in origins) + tile_out >= size and ranks <= len(origins)
            c = cost(size, t, ranks, bt)
            assert c <= cost(size, 572, min(world, len(tiling.plan_tiles(size, size, 572)[2])), 8) + 1e-9
            assert prev is None or c <= prev + 1e-9          # more GPUs never model slower
            prev = c
    # 8192^2: 16 tiles of 2236 -> 2052 (stride 2048, 1.19 input pixels per output pixel; 64 tiles of
    # 1212 -> 1028 execute 1.40), two per rank on eight GPUs
    assert tiling.choose_plan(8192, 8192, 8) == (2236, 8, 2)
    assert tiling.choose_plan(8192, 8192, 1) == (2236, 1, 8)
    assert tiling.choose_plan(1024, 1024, 1)[:2] == (1212, 1)


def test_tile_session_tables_cover_every_tile_once_and_leave_idle_ranks_empty():
    """Host logic of the sharded overlap-tile path (tiling._TileSession): the (world, slots, 2) origin
    table that drives ub_extract_tiles / ub_stitch_tiles holds every tile exactly once, round-robin
    over the ranks in use, unused slots and idle ranks are marked with a negative row, slots come in
    whole batches, and every rank derives the same table."""
    import types

    import torch

    from unet_segmentation_b200 import tiling

    model = types.SimpleNamespace(levels=5)
    dev = torch.device("cpu")
    for (h, w, tile_in, bt, world, used) in [(1024, 1024, 700, 8, 8, 4), (8192, 8192, 2236, 8, 8, 8),
                                             (500, 420, 572, 2, 1, 1), (800, 800, 572, 4, 3, 3),
                                             (1024, 1024, 444, 2, 8, 8), (300, 300, 252, 2, 4, 9)]:
        tables = []
        for rank in range(world):
            s = tiling._TileSession(model, h, w, tile_in, bt, rank, world, used, dev)
            tables.append(s.table)
            assert s.slots % s.bt == 0 and s.table.shape == (world, s.slots, 2)
            assert s.n_batches * s.bt >= len(s.mine)
            assert torch.equal(s.my_table, s.table[rank])
            if rank >= s.ranks_used:
                assert s.mine == [] and s.n_batches == 0 and bool((s.my_table[:, 0] < 0).all())
        assert all(torch.equal(t, tables[0]) for t in tables)
        t = tables[0]
        valid = t[t[:, :, 0] >= 0]
        got = sorted((int(y), int(x)) for y, x in valid.tolist())
        assert got == sorted(s.origins)                      # every tile once, none invented
        assert s.ranks_used <= min(world, len(s.origins))
        per_rank = (t[:, :, 0] >= 0).sum(1)
        assert int(per_rank[:s.ranks_used].max() - per_rank[:s.ranks_used].min()) <= 1   # balanced
        assert int(per_rank[s.ranks_used:].sum()) == 0


def test_bf16_rounding_oracle_is_a_small_perturbation_of_the_fp32_oracle():
    """T2 oracle plumbing (oracle/unet_ref.py emulate_bf16): same graph, bf16 rounding at the conv
    operands except the first conv — close to, but not identical with, the fp32 restatement, and
    gradients still reach every parameter."""
    import torch

    from oracle import unet_ref

    sd = unet_ref.make_state_dict(1, 2, seed=3)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    img, t, w = unet_ref.synthetic_batch(1, size=188, seed=5)
    ref = unet_ref.unet_forward(full, img, training=True)
    emu = unet_ref.unet_forward(full, img, training=True, emulate_bf16=True)
    rel = float((emu - ref).norm() / ref.norm())
    assert 1e-5 < rel < 5e-2
    unet_ref.weighted_cross_entropy(emu, t, w).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params.values())
