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
