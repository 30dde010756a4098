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


def _native_extract(img, origins, tile_in, margin):
    import ctypes as C

    from unet_segmentation_b200 import _lib

    lib = _lib.load()
    table = torch.tensor([[y, x] for y, x in origins], dtype=torch.int32, device=img.device)
    out = torch.empty(len(origins), 1, tile_in, tile_in, dtype=torch.float32, device=img.device)
    _lib.check(lib.ub_extract_tiles(C.c_void_p(img.data_ptr()), img.shape[0], img.shape[1],
                                    C.c_void_p(table.data_ptr()), len(origins), tile_in, margin,
                                    C.c_void_p(out.data_ptr()),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "extract")
    return out


@pytest.mark.parametrize("h,w,tile_in", [(500, 420, 572), (97, 131, 252), (1024, 1024, 700), (7, 5, 28)])
def test_native_tile_gather_and_stitch_match_the_host_reference(h, w, tile_in):
    """ub_extract_tiles (mirror rule, any extension length: the 7 x 5 image is reflected several
    times) against tiling.extract_tiles (torch indexing, itself checked against numpy.pad and the
    oracle on the CPU), and ub_stitch_tiles against a slice-by-slice stitch; unused slots skipped."""
    import ctypes as C

    from unet_segmentation_b200 import _lib, tiling

    g = torch.Generator().manual_seed(h * w)
    img = torch.rand(h, w, generator=g).cuda()
    margin = 92 if tile_in > 200 else 9
    tile_out = tile_in - 2 * margin
    stride = (tile_out // 16) * 16 if tile_out >= 16 else tile_out
    origins = [(y, x) for y in range(0, h, max(stride, 1)) for x in range(0, w, max(stride, 1))]
    got = _native_extract(img, origins + [(-1, 0)], tile_in, margin)
    ref = tiling.extract_tiles(img, origins, tile_in, margin)
    torch.cuda.synchronize()
    assert torch.equal(got[:-1], ref)
    assert float(got[-1].abs().max()) == 0.0                    # unused slot: zeros
    if tile_out % 4:
        return
    lib = _lib.load()
    tiles = (torch.rand(len(origins) + 1, tile_out, tile_out, generator=g) * 255).to(torch.uint8).cuda()
    # make overlapping regions agree (the invariant the real pipeline guarantees): tiles = crops of one image
    canvas = (torch.rand(h + tile_out, w + tile_out, generator=g) * 255).to(torch.uint8).cuda()
    for k, (y, x) in enumerate(origins):
        tiles[k] = canvas[y:y + tile_out, x:x + tile_out]
    table = torch.tensor([[y, x] for y, x in origins] + [[-1, 0]], dtype=torch.int32, device="cuda")
    full = torch.full((h, w), 7, dtype=torch.uint8, device="cuda")
    _lib.check(lib.ub_stitch_tiles(C.c_void_p(tiles.data_ptr()), C.c_void_p(table.data_ptr()),
                                   len(origins) + 1, tile_out, C.c_void_p(full.data_ptr()), h, w,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "stitch")
    torch.cuda.synchronize()
    assert torch.equal(full, canvas[:h, :w])


def test_eval_forward_replays_from_a_cuda_graph_bit_identically():
    """An eval plan called again with its own static buffers is captured on the second call and
    replayed afterwards (ub_plan_forward); the replays equal the plain launches bit for bit, follow
    weight and BN-buffer updates (the graph reads them through the bound pointers), and a new input
    shape gets its own plan."""
    model, _ = _eval_model(seed=8)
    img = (0.4 + 0.2 * torch.rand(2, 1, 252, 252, generator=torch.Generator().manual_seed(3))).cuda()
    first, m1 = model.predict_mask(img)                       # plain launches
    plan = next(p for k, p in model._plans.items() if not k[4])
    assert plan.graph_replays() == 0
    second, m2 = model.predict_mask(img)                      # capture + first replay
    third, m3 = model.predict_mask(img)
    torch.cuda.synchronize()
    assert plan.graph_replays() == 2
    assert torch.equal(first, second) and torch.equal(first, third)
    assert torch.equal(m1, m2) and torch.equal(m1, m3)
    other = (0.4 + 0.2 * torch.rand(2, 1, 252, 252, generator=torch.Generator().manual_seed(4))).cuda()
    o1, _ = model.predict_mask(other)                         # same plan, new data: still the graph
    assert plan.graph_replays() == 3 and not torch.equal(o1, first)
    assert torch.equal(model.predict_mask(img)[0], first)
    with torch.no_grad():                                     # weights change -> operands re-packed
        model.up4.conv.double_conv[3].weight.mul_(0.5)
        model.inc.double_conv[1].running_mean.add_(0.05)
    changed, _ = model.predict_mask(img)
    torch.cuda.synchronize()
    assert not torch.equal(changed, first)
    from unet_segmentation_b200.unet import UNet

    twin = UNet(1, 2).cuda().eval()
    twin.load_state_dict(model.state_dict())
    fresh, _ = twin.predict_mask(img)                         # plain launches on a fresh plan
    torch.cuda.synchronize()
    assert torch.equal(changed, fresh)


def _tiled_golden_mask(size):
    import os

    blob = np.load(os.path.join(os.path.dirname(__file__), "golden", "ccl_golden.npz"))
    keys = sorted(k for k in blob.files if k.startswith("mask"))
    reps = -(-size // 324)
    rows = []
    for r in range(reps):
        rows.append(np.concatenate([blob[keys[(r + c) % len(keys)]] for c in range(reps)], axis=1))
    return np.ascontiguousarray(np.concatenate(rows, axis=0)[:size, :size])


@pytest.mark.parametrize("size", [2048, 8192])
def test_ccl_on_a_stitched_whole_image_mask(size):
    """Whole-image post-processing (scripts/predict.py:85-98 on the stitched mask):
    the reference's shipped binary masks (~1 700 speckly components per 324^2 frame) tiled to the
    full image — about 10^6 components at 8192^2, far beyond uint16, so the wrap-around of the
    reference's astype(uint16) is exercised too — labelled on the GPU, bit-exact against the oracle."""
    from unet_segmentation_b200.postprocess import get_instance_masks

    mask = _tiled_golden_mask(size)
    ref = ccl_ref.get_instance_masks(mask, 15)
    dev = get_instance_masks(torch.from_numpy(mask).cuda(), min_size=15)
    torch.cuda.synchronize()
    got = dev.cpu().numpy()
    assert got.dtype == np.uint16 and got.shape == (size, size)
    assert np.array_equal(got, ref)
    # one giant component (a solid mask with a few holes): the contended-atomics case
    solid = np.full((size, size), 255, dtype=np.uint8)
    solid[::97, ::89] = 0
    solid[size // 2, :] = 0                                   # split into two components
    ref2 = ccl_ref.get_instance_masks(solid, 15)
    got2 = get_instance_masks(torch.from_numpy(solid).cuda(), min_size=15).cpu().numpy()
    assert np.array_equal(got2, ref2) and int(ref2.max()) == 2


def test_overlap_tile_to_instance_labels_end_to_end():
    """predict path on a large image entirely on the device: overlap-tile mask -> stitched-mask
    connected components, equal to running the oracle's labelling on the same stitched mask, and
    identical when the tiles are dealt to 3 emulated ranks of which only 2 are used."""
    from unet_segmentation_b200 import tiling
    from unet_segmentation_b200.postprocess import get_instance_masks

    model, _ = _eval_model(seed=4)
    img = (0.4 + 0.2 * torch.rand(1100, 900, generator=torch.Generator().manual_seed(1))).cuda()
    full = tiling.overlap_tile_predict(model, img, tile_in=572, batch_tiles=4)
    again = tiling.overlap_tile_predict(model, img, tile_in=572, batch_tiles=4)   # graph replays
    auto = tiling.overlap_tile_predict(model, img, tile_in=None)                  # choose_plan
    torch.cuda.synchronize()
    assert torch.equal(full, again) and torch.equal(full, auto)
    labels = get_instance_masks(full, min_size=15)
    torch.cuda.synchronize()
    assert np.array_equal(labels.cpu().numpy(), ccl_ref.get_instance_masks(full.cpu().numpy(), 15))


def test_frame_predictor_graph_equals_the_eager_predict_loop():
    """unet_segmentation_b200.predict.FramePredictor (scripts/predict.py:73-112 as one CUDA-graph launch per
    frame) against the same steps issued one by one, and its labels against the oracle."""
    from unet_segmentation_b200.postprocess import get_instance_masks
    from unet_segmentation_b200.predict import FramePredictor

    model, _ = _eval_model(seed=6)
    with torch.no_grad():          # move the decision boundary into the logits' range: a speckly mask
        img0 = (0.4 + 0.2 * torch.rand(1, 1, 512, 512, generator=torch.Generator().manual_seed(2))).cuda()
        z = model(img0)
        model.outc.conv.bias[1] += (z[:, 0] - z[:, 1]).median()
    fp = FramePredictor(model, (512, 512), min_size=15)
    for seed in (2, 3, 4):
        frame = (0.4 + 0.2 * torch.rand(512, 512, generator=torch.Generator().manual_seed(seed))).numpy()
        mask, labels = fp(frame)
        mask, labels = mask.copy(), labels.copy()
        x = torch.from_numpy(frame).reshape(1, 1, 512, 512).cuda()
        _, ref_mask = model.predict_mask(x)
        ref_labels = get_instance_masks(ref_mask[0], 15)
        torch.cuda.synchronize()
        assert mask.shape == (324, 324) and mask.dtype == np.uint8 and labels.dtype == np.uint16
        assert np.array_equal(mask, ref_mask[0].cpu().numpy())
        assert np.array_equal(labels, ref_labels.cpu().numpy())
        assert np.array_equal(labels, ccl_ref.get_instance_masks(mask, 15))
        assert 0.02 < (mask > 0).mean() < 0.98 and labels.max() > 1
    # weights change: the predictor re-records (packed operands are refreshed outside the graph)
    with torch.no_grad():
        model.outc.conv.weight.mul_(-1.0)
    mask2, _ = fp(frame)
    _, ref2 = model.predict_mask(x)
    torch.cuda.synchronize()
    assert np.array_equal(mask2, ref2[0].cpu().numpy())
