"""Overlap-tile inference (BASELINE config 4) on the GPU. The reference ships no tiling code
(SURVEY F2), so parity is the invariant the semantics were designed around: with eval-mode BN, tile
size ≡ 12 (mod 16) and tile origins ≡ 0 (mod 16), the stitched tile-wise logits equal one forward pass
over the whole mirror-extended image; masks then feed the bit-exact connected-component step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ccl_ref, unet_ref  # noqa: E402


def _eval_model(seed=3):
    from unet_segmentation_b200.unet import UNet

    sd = unet_ref.make_state_dict(1, 2, seed=seed)
    gen = torch.Generator().manual_seed(17)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = torch.randn(nf, generator=gen) * 0.1
        sd[k.replace("running_mean", "running_var")] = 0.5 + torch.rand(nf, generator=gen)
    m = UNet(1, 2)
    m.load_state_dict(sd)
    return m.cuda().eval(), {k: v.cuda() for k, v in sd.items()}


def test_tiles_equal_whole_image_forward():
    from unet_segmentation_b200 import tiling

    model, sd = _eval_model()
    h, w = 500, 420                      # 2 x 2 tiles of 388 outputs, ragged right / bottom edges
    g = torch.Generator().manual_seed(5)
    img = (0.4 + 0.2 * torch.rand(h, w, generator=g)).cuda()
    mask, logits = tiling.overlap_tile_predict(model, img, tile_in=572, batch_tiles=2,
                                               return_logits=True)
    assert mask.shape == (h, w) and logits.shape == (2, h, w)
    # whole-image pass through the same kernels: mirror-extend so that the (aligned) output covers
    # the image; 764 + 184 = 948 ≡ 12 (mod 16)
    tout, stride, origins = tiling.plan_tiles(h, w, 572)
    full_out = origins[-1][0] + tout      # 772: rows covered by the tiling
    big = tiling.extract_tiles(img, [(0, 0)], tile_in=full_out + 184, margin=92)
    whole, _ = model.predict_mask(big.contiguous())
    torch.cuda.synchronize()
    assert whole.shape[-1] == full_out
    assert torch.equal(logits, whole[0, :, :h, :w])          # bit-exact (SURVEY §8c invariant)
    assert torch.equal(mask > 0, logits[1] > logits[0])
    # and against the fp32 oracle on the same mirror-extended image (bf16 tolerance)
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, big, training=False)[0, :, :h, :w]
    rel = float((logits.double() - ref.double()).norm() / ref.double().norm())
    agree = float(((logits[1] > logits[0]) == (ref[1] > ref[0])).float().mean())
    print(f"\n[overlap-tile 500x420] logits rel-L2 vs fp32 oracle {rel:.3e}, mask agreement {agree:.5f}")
    assert rel < 2e-2 and agree > 0.995


def test_tile_sharding_is_rank_invariant():
    """Dealing tiles to 'ranks' and stitching gives the same mask as one rank doing everything."""
    from unet_segmentation_b200 import parallel, tiling

    model, _ = _eval_model(seed=4)
    img = (0.4 + 0.2 * torch.rand(800, 800, generator=torch.Generator().manual_seed(1))).cuda()
    full = tiling.overlap_tile_predict(model, img, batch_tiles=4)
    tout, _, origins = tiling.plan_tiles(800, 800, 572)
    assert len(origins) == 9
    stitched = torch.zeros_like(full)
    for rank in range(3):                     # emulate world = 3 on one device
        idx = parallel.shard_indices(len(origins), rank, 3)
        tiles = tiling.extract_tiles(img, [origins[i] for i in idx], 572, 92)
        _, m = model.predict_mask(tiles.contiguous())
        for k, i in enumerate(idx):
            y, x = origins[i]
            hh, ww = min(tout, 800 - y), min(tout, 800 - x)
            stitched[y:y + hh, x:x + ww] = m[k][:hh, :ww]
    torch.cuda.synchronize()
    assert torch.equal(full, stitched)


def test_predict_postprocess_matches_reference_restatement():
    """scripts/predict.py:85-98 on the GPU: mask = softmax[:,1] > 0.5, then get_instance_masks."""
    from unet_segmentation_b200.postprocess import get_instance_masks, instance_labels_from_logits

    model, _ = _eval_model(seed=6)
    img = (0.4 + 0.2 * torch.rand(1, 1, 512, 512, generator=torch.Generator().manual_seed(2))).cuda()
    logits, mask = model.predict_mask(img)
    prob = torch.softmax(logits, dim=1)[:, 1]
    ref_mask = ((prob > 0.5).cpu().numpy()[0] * 255).astype(np.uint8)
    got_mask = mask[0].cpu().numpy()
    # softmax > 0.5 and z1 > z0 may differ only at exact fp ties
    assert (ref_mask != got_mask).mean() < 1e-5
    labels = get_instance_masks(got_mask, min_size=15)          # numpy in / numpy out, as the reference
    assert labels.dtype == np.uint16 and labels.shape == (324, 324)
    assert np.array_equal(labels, ccl_ref.get_instance_masks(got_mask, 15))
    dev = instance_labels_from_logits(logits, 15)
    assert np.array_equal(dev[0].cpu().numpy(), labels)


def test_ccl_on_reference_golden_pairs():
    import os

    from unet_segmentation_b200.postprocess import get_instance_masks

    blob = np.load(os.path.join(os.path.dirname(__file__), "golden", "ccl_golden.npz"))
    for i in sorted(k[4:] for k in blob.files if k.startswith("mask")):
        assert np.array_equal(get_instance_masks(blob[f"mask{i}"], 15), blob[f"inst{i}"]), i
