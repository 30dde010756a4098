"""CPU tests of the overlap-tile semantics (BASELINE config 4; parity unpinned): the
oracle restatement satisfies the aligned-tile invariant that stands in for golden vectors, and the
product's host-side tiling logic (tile plan, mirror extension, tile extraction — plain torch, no
kernels) agrees with the independently written oracle."""
import numpy as np
import pytest
import torch

from oracle import overlap_tile_ref, unet_ref


def _eval_state_dict(levels, seed=3):
    sd = unet_ref.make_state_dict(1, 2, seed=seed, levels=levels)
    gen = torch.Generator().manual_seed(17)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = torch.randn(nf, generator=gen) * 0.1
        sd[k.replace("running_mean", "running_var")] = 0.5 + torch.rand(nf, generator=gen)
    return sd


@pytest.mark.parametrize("levels,tile_in,h,w", [(3, 76, 70, 50), (5, 204, 37, 30)])
def test_oracle_tiles_equal_one_whole_image_forward(levels, tile_in, h, w):
    """Tiles of size ≡ 12 (mod 16) at origins ≡ 0 (mod 16): stitched logits == one forward over the
    mirror-extended image (SURVEY §8c invariant), up to fp32 summation order inside oneDNN."""
    sd = _eval_state_dict(levels)
    g = torch.Generator().manual_seed(9)
    img = 0.4 + 0.2 * torch.rand(h, w, generator=g)
    tile_out, stride, origins = overlap_tile_ref.tile_origins(h, w, tile_in, levels)
    assert len(origins) >= 4 and stride % 16 == 0
    tiled = overlap_tile_ref.overlap_tile_logits(sd, img, tile_in, levels)
    whole = overlap_tile_ref.whole_image_logits(sd, img, tile_in, levels)
    assert tiled.shape == whole.shape == (2, h, w)
    scale = float(whole.abs().max())
    assert float((tiled - whole).abs().max()) <= 2e-5 * scale
    margin_px = (whole[1] - whole[0]).abs() > 1e-4 * scale
    assert torch.equal(overlap_tile_ref.mask_from_logits(tiled)[margin_px],
                       overlap_tile_ref.mask_from_logits(whole)[margin_px])


def test_misaligned_tiles_do_differ():
    """The alignment conditions are what makes the invariant hold: a tile whose origin is odd sees
    different pooling windows and its logits differ visibly (~1e-3 relative)."""
    levels, tile_in = 3, 76
    sd = _eval_state_dict(levels)
    g = torch.Generator().manual_seed(4)
    img = 0.4 + 0.2 * torch.rand(140, 140, generator=g)
    m = overlap_tile_ref.margin(levels)
    whole = unet_ref.unet_forward(sd, img[None, None], training=False, levels=levels)[0]
    t_out = tile_in - 2 * m

    def tile_at(y, x):
        z = unet_ref.unet_forward(sd, img[y:y + tile_in, x:x + tile_in][None, None], training=False,
                                  levels=levels)[0]
        return z, whole[:, y:y + t_out, x:x + t_out]

    z, ref = tile_at(16, 32)
    assert float((z - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    z, ref = tile_at(17, 32)
    assert float((z - ref).abs().max()) > 1e-4 * float(ref.abs().max())


def test_product_tile_plan_and_extraction_match_the_oracle():
    from unet_segmentation_b200 import tiling

    for levels in (3, 4, 5):
        assert tiling.network_margin(levels) == overlap_tile_ref.margin(levels)
    assert overlap_tile_ref.margin(5) == 92          # 572 -> 388, models/unet_model.py:175-187
    rng = np.random.default_rng(0)
    for tile_in in (204, 220, 572):
        for h, w in [(1, 1), (20, 36), (37, 35), (388, 389), (400, 770), (800, 500), (1024, 1024)]:
            t_out, stride, origins = overlap_tile_ref.tile_origins(h, w, tile_in)
            assert (t_out, stride, origins) == tiling.plan_tiles(h, w, tile_in), (tile_in, h, w)
            assert all(y % 16 == 0 and x % 16 == 0 for y, x in origins)
            assert max(y for y, _ in origins) + t_out >= h and max(x for _, x in origins) + t_out >= w
    for (h, w), tile_in in [((37, 30), 204), ((61, 90), 220), ((5, 3), 204)]:
        img = rng.random((h, w)).astype(np.float32)
        ext = overlap_tile_ref.mirror_extend(img, tile_in)
        _, _, origins = overlap_tile_ref.tile_origins(h, w, tile_in)
        tiles = tiling.extract_tiles(torch.from_numpy(img), origins, tile_in, overlap_tile_ref.margin())
        assert tiles.shape == (len(origins), 1, tile_in, tile_in)
        for k, (y, x) in enumerate(origins):
            assert np.array_equal(tiles[k, 0].numpy(), ext[y:y + tile_in, x:x + tile_in]), (h, w, k)
