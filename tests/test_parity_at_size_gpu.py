"""Parity at the sizes BASELINE.json benchmarks (VERDICT r1 "what's missing" 1): the CUDA path through
the C ABI against the fp32 oracle on the same GPU (TF32 off, tests/conftest.py) at

  * configs[1]: batch 16 x 1 x 512 x 512 training step (the shape bench.py times),
  * configs[4]: the wide U-Net (base 128, depth 5) on a 1024 x 1024 crop,
  * configs[3]: the 8192 x 8192 overlap-tile plan bench.py runs (16 tiles 2236 -> 2052, batch 8),
  * configs[0]: reference-held golden vectors at 512 x 512 (recorded from the unmodified reference by
    oracle/make_golden.py) at the north-star tolerance itself,

with the `torch.autocast(bf16)` control of SURVEY F3 / §8c COMPUTED next to the T1 numbers (the gate is
"no worse than the control on every metric"; the north-star bars that bf16 can meet are asserted as
such: logits / loss 2e-2).
Reference path: /root/reference/models/unet_model.py:105-146, utils/losses.py:49-57.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import unet_ref  # noqa: E402

ZERO_GRAD_BIASES = (".double_conv.0.bias", ".double_conv.3.bias", ".up.bias")


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def mask_agreement(a, b):
    return float(((a[:, 1] > a[:, 0]) == (b[:, 1] > b[:, 0])).float().mean())


def build(base=64, levels=5, seed=0):
    from unet_segmentation_b200.unet import UNet

    sd = unet_ref.make_state_dict(1, 2, seed=seed, base=base, levels=levels)
    m = UNet(1, 2, base_channels=base, levels=levels)
    m.load_state_dict(sd)
    return m.cuda(), {k: v.cuda() for k, v in sd.items()}


def oracle_step(sd, img, t, w, levels=5, autocast=False):
    """One fp32 step of the oracle on the GPU; autocast=True is the control of SURVEY F3: the same
    network under torch.autocast(bf16) (what stock PyTorch mixed precision computes)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = unet_ref.unet_forward(full, img, training=True, levels=levels, buffers_out={})
    logits = logits.float()
    loss = unet_ref.weighted_cross_entropy(logits, t, w)
    loss.backward()
    grads = {k: p.grad.float() for k, p in params.items()}
    return logits.detach(), loss.detach(), grads


def t1_metrics(logits, loss, grads, ref_logits, ref_loss, ref_grads):
    cos = {k: cosine(g, ref_grads[k]) for k, g in grads.items() if not k.endswith(ZERO_GRAD_BIASES)}
    vals = np.array(list(cos.values()))
    conf = (ref_logits[:, 1] - ref_logits[:, 0]).abs() > 0.05
    same = (logits[:, 1] > logits[:, 0]) == (ref_logits[:, 1] > ref_logits[:, 0])
    return {"logits": rel_l2(logits, ref_logits),
            "mask_conf": float(same[conf].float().mean()), "conf_frac": float(conf.float().mean()),
            "loss": abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)),
            "mask": mask_agreement(logits, ref_logits),
            "cos_min": float(vals.min()), "cos_median": float(np.median(vals)),
            "n999": int((vals >= 0.999).sum()), "n": len(vals), "cos": cos}


def fmt(tag, m):
    return (f"{tag}: logits rel-L2 {m['logits']:.3e}  loss rel {m['loss']:.3e}  mask agree "
            f"{m['mask']:.5f} ({m['mask_conf']:.5f} on the {m['conf_frac']:.3f} with |z1-z0|>0.05)  grad cos min {m['cos_min']:.4f} median {m['cos_median']:.4f}  "
            f">=0.999: {m['n999']}/{m['n']}")


def library_step(model, img, t, w):
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    model.train()
    model.zero_grad(set_to_none=True)
    logits = model(img)
    loss = WeightedCrossEntropyLoss()(logits, t, w)
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad for n, p in model.named_parameters()}
    return logits.detach(), loss.detach(), grads


def assert_no_worse_than_control(ours, ctrl, tag):
    """SURVEY §8c gate: on every T1 metric the library must be at least as close to the fp32
    reference as stock torch.autocast(bf16) is (small slack for run-to-run noise of the control)."""
    assert ours["logits"] <= ctrl["logits"] * 1.02, (tag, ours["logits"], ctrl["logits"])
    assert ours["loss"] <= max(ctrl["loss"] * 1.5, 5e-4), (tag, ours["loss"], ctrl["loss"])
    assert ours["mask"] >= ctrl["mask"] - 1e-4, (tag, ours["mask"], ctrl["mask"])
    assert ours["cos_min"] >= ctrl["cos_min"] - 0.02, (tag, ours["cos_min"], ctrl["cos_min"])
    assert ours["cos_median"] >= ctrl["cos_median"] - 0.01, (tag, ours["cos_median"], ctrl["cos_median"])


@pytest.mark.parametrize("n", [2, 16])
def test_training_step_vs_fp32_oracle_and_autocast_control_512(n):
    """N=2 is the survey's experiment (Appendix B); N=16 x 512^2 is BASELINE configs[1], the exact
    shape bench.py times (different tile counts, split-K factors and CTA-pair decisions than N=2)."""
    model, sd = build(seed=0)
    img, t, w = unet_ref.synthetic_batch(n, size=512, seed=1234, device="cuda")
    ref = oracle_step(sd, img, t, w)
    ctrl = t1_metrics(*oracle_step(sd, img, t, w, autocast=True), *ref)
    logits, loss, grads = library_step(model, img, t, w)
    assert logits.shape == (n, 2, 324, 324) and logits.dtype == torch.float32
    for name, g in grads.items():
        assert g is not None and torch.isfinite(g).all(), name
        if name.endswith(ZERO_GRAD_BIASES):
            wname = name.replace(".bias", ".weight")
            assert float(g.norm()) <= 1e-3 * float(ref[2][wname].norm()) + 1e-12, name
    ours = t1_metrics(logits, loss, grads, *ref)
    print("\n" + fmt(f"[T1 N={n} x 512^2] libunetb200          ", ours))
    print(fmt(f"[T1 N={n} x 512^2] torch.autocast(bf16) ctl", ctrl))
    # north-star bars a bf16 implementation can meet end to end
    assert ours["logits"] < 2e-2 and ours["loss"] < 2e-3
    # measured (N=2 / N=16): masks 0.99568 / 0.99574 overall, 0.99745 / 0.99751 where |z1-z0| > 0.05;
    # the torch.autocast control: 0.9823 / 0.9842. 99.9 % is met in eval mode only.
    assert ours["mask"] > 0.995 and ours["mask_conf"] > 0.997
    for name in ("outc.conv.weight", "outc.conv.bias", "up4.conv.double_conv.4.weight",
                 "up4.conv.double_conv.3.weight"):
        assert ours["cos"][name] > 0.999, (name, ours["cos"][name])
    assert ours["cos_min"] > 0.85 and ours["cos_median"] > 0.94
    assert_no_worse_than_control(ours, ctrl, f"N={n}")


def test_wide_unet_base128_1024_vs_fp32_oracle():
    """BASELINE configs[4]: wide U-Net (base 128 channels, depth 5) on a 1024 x 1024 crop — every
    3x3 layer runs the 128/256-column tiles and CTA pairs, the skip concat is 1020^2 x 256 channels."""
    model, sd = build(base=128, seed=2)
    img, t, w = unet_ref.synthetic_batch(1, size=1024, seed=31, device="cuda")
    ref = oracle_step(sd, img, t, w)
    ctrl = t1_metrics(*oracle_step(sd, img, t, w, autocast=True), *ref)
    logits, loss, grads = library_step(model, img, t, w)
    assert logits.shape == (1, 2, 836, 836)
    ours = t1_metrics(logits, loss, grads, *ref)
    print("\n" + fmt("[T1 base128 1 x 1024^2] libunetb200          ", ours))
    print(fmt("[T1 base128 1 x 1024^2] torch.autocast(bf16) ctl", ctrl))
    assert ours["logits"] < 2e-2 and ours["loss"] < 2e-3 and ours["mask"] > 0.996   # measured 0.99676
    assert ours["cos"]["outc.conv.weight"] > 0.999 and ours["cos"]["outc.conv.bias"] > 0.999
    assert_no_worse_than_control(ours, ctrl, "base128")


def _eval_sd(seed):
    sd = unet_ref.make_state_dict(1, 2, seed=seed)
    gen = torch.Generator().manual_seed(17)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = torch.randn(nf, generator=gen) * 0.1
        sd[k.replace("running_mean", "running_var")] = 0.5 + torch.rand(nf, generator=gen)
    return sd


def test_overlap_tile_8192_benchmarked_plan_vs_whole_image_forward():
    """BASELINE configs[3] at full size with the plan bench.py uses (choose_plan -> 16 tiles of
    2236 -> 2052, batch 8): (a) the stitched logits equal ONE forward pass of the library over a
    2068 x 2068-output window that straddles 2 x 2 tiles (the aligned-tile invariant, bit-exact);
    (b) the same window against the fp32 oracle at the bf16 tolerance; (c) the mask is the
    thresholded stitched logits everywhere."""
    from unet_segmentation_b200 import tiling
    from unet_segmentation_b200.unet import UNet

    sd = _eval_sd(seed=3)
    model = UNet(1, 2)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    sd = {k: v.cuda() for k, v in sd.items()}
    size = 8192
    g = torch.Generator().manual_seed(99)
    frame = 0.4 + 0.2 * torch.rand(512, 512, generator=g)
    # mosaic of 512^2 frames with a smooth illumination ramp so tiles are not identical
    ramp = torch.linspace(-0.05, 0.05, size)
    img = (frame.repeat(size // 512, size // 512) + ramp[None, :] + ramp[:, None]).cuda()
    tile_in, _, bt = tiling.choose_plan(size, size, 1)
    tile_out, stride, origins = tiling.plan_tiles(size, size, tile_in)
    assert (tile_in, tile_out, len(origins), bt) == (2236, 2052, 16, 8)
    mask, logits = tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=8,
                                               return_logits=True)
    torch.cuda.synchronize()
    assert mask.shape == (size, size) and logits.shape == (2, size, size)
    assert torch.equal(mask > 0, logits[1] > logits[0])
    # one whole-window forward: output rows/cols [a, a + S), a = 1008 (multiple of 16),
    # S = 2068 (S + 184 = 2252 = 12 mod 16) covers parts of tiles 0 and 1 in both directions
    a, S = 1008, 2068
    margin = tiling.network_margin(5)
    big = tiling.extract_tiles(img, [(a, a)], tile_in=S + 2 * margin, margin=margin).contiguous()
    whole, wmask = model.predict_mask(big)
    torch.cuda.synchronize()
    assert whole.shape == (1, 2, S, S)
    assert torch.equal(logits[:, a:a + S, a:a + S], whole[0])        # bit-exact
    assert torch.equal(mask[a:a + S, a:a + S], wmask[0])
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, big, training=False)[0]
    rel = rel_l2(whole[0], ref)
    agree = float(((whole[0, 1] > whole[0, 0]) == (ref[1] > ref[0])).float().mean())
    conf = (ref[1] - ref[0]).abs() > 0.05
    agree_conf = float(((whole[0, 1] > whole[0, 0]) == (ref[1] > ref[0]))[conf].float().mean())
    print(f"\n[overlap-tile 8192^2, 16 tiles 2236->2052] window {S}^2 vs fp32 oracle: logits rel-L2 "
          f"{rel:.3e}, mask agreement {agree:.5f} ({agree_conf:.5f} on the "
          f"{float(conf.float().mean()):.3f} of pixels with |z1-z0| > 0.05)")
    assert rel < 2e-2 and agree_conf >= 0.999


@pytest.mark.parametrize("name", ["train_n1_s512", "eval_n1_s512"])
def test_against_reference_golden_512(name):
    """Vectors recorded from the UNMODIFIED reference on CPU (oracle/make_golden.py) at the real
    operating point, checked at the north-star tolerance (2e-2 on logits and loss) — the small
    fixtures of test_unet_gpu.py need looser bars because 4 x 4 logits amplify bf16 noise."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    blob = np.load(os.path.join(GOLD, "unet_golden_512.npz"))
    c = {k[len(name) + 1:]: blob[k] for k in blob.files if k.startswith(name + "/")}
    n, size, sw, sx, training = [int(v) for v in c["meta"]]
    model, _ = build(seed=sw)
    img, t, w = unet_ref.synthetic_batch(n, size=size, seed=sx, device="cuda")
    ref_logits = torch.from_numpy(c["logits"]).cuda()
    if training:
        model.train()
        logits = model(img)
        loss = WeightedCrossEntropyLoss()(logits, t, w)
        loss.backward()
        torch.cuda.synchronize()
        e_loss = abs(float(loss) - float(c["loss"])) / float(c["loss"])
        e_logits = rel_l2(logits, ref_logits)
        agree = mask_agreement(logits, ref_logits)
        print(f"\n[reference golden {name}] logits rel-L2 {e_logits:.3e}  loss rel {e_loss:.3e}  "
              f"mask agree {agree:.5f}")
        assert e_logits < 2e-2 and e_loss < 2e-3 and agree > 0.995     # measured 0.99566
        grads = dict(model.named_parameters())
        for k in ("outc.conv.weight", "outc.conv.bias", "up4.conv.double_conv.4.weight",
                  "up4.conv.double_conv.4.bias"):
            g = grads[k].grad.flatten().cpu()
            ref_norm = float(c[f"gnorm/{k}"])
            assert abs(float(g.double().norm()) - ref_norm) < 2e-2 * ref_norm, k
            got = g[torch.from_numpy(c[f"gidx/{k}"])].numpy()
            assert np.abs(got - c[f"gval/{k}"]).max() < 3e-2 * np.abs(c[f"gval/{k}"]).max() + 1e-9, k
        new = model.state_dict()
        for k in [k for k in c if k.startswith("buf/")]:
            ref_b = torch.from_numpy(c[k])
            if k.endswith("num_batches_tracked"):
                assert int(new[k[4:]]) == int(ref_b), k
            else:
                assert rel_l2(new[k[4:]].cpu(), ref_b) < 2e-2, k
    else:
        gen = torch.Generator().manual_seed(99)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
        model.eval()
        with torch.no_grad():
            logits, mask = model.predict_mask(img)
        torch.cuda.synchronize()
        e_logits = rel_l2(logits, ref_logits)
        ref_mask = ref_logits[:, 1] > ref_logits[:, 0]
        conf = (ref_logits[:, 1] - ref_logits[:, 0]).abs() > 0.05
        agree = float(((mask > 0) == ref_mask).float().mean())
        agree_conf = float(((mask > 0) == ref_mask)[conf].float().mean())
        print(f"\n[reference golden {name}] logits rel-L2 {e_logits:.3e}  mask agree {agree:.5f} "
              f"({agree_conf:.5f} where |z1-z0| > 0.05)")
        assert e_logits < 2e-2 and agree_conf >= 0.999
