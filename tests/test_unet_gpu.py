"""End-to-end parity of the drop-in UNet + WeightedCrossEntropyLoss on the GPU (through the C ABI)
against (a) golden vectors recorded from the reference itself and (b) the fp32 oracle restatement
run on the same device — SURVEY §8c tiers T1. Bars from BASELINE.json north_star: logits and loss
within 2e-2 relative; per-layer gradient cosine is reported and bounded by the measured bf16 floor
(0.999 end-to-end is unattainable for any bf16 implementation, incl. torch.autocast;
it is enforced teacher-forced per layer in test_ops_gpu.py)."""
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


def make_model(seed, n_classes=2):
    from unet_segmentation_b200.unet import UNet

    sd = unet_ref.make_state_dict(1, n_classes, seed=seed)
    m = UNet(1, n_classes)
    m.load_state_dict(sd)
    return m.cuda(), {k: v.cuda() for k, v in sd.items()}


def oracle_step(sd, img, t, w, training=True, emulate_bf16=False):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    bufs = {}
    logits = unet_ref.unet_forward(full, img, training=training, buffers_out=bufs,
                                   emulate_bf16=emulate_bf16)
    loss = unet_ref.weighted_cross_entropy(logits, t, w)
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad for k, p in params.items()}, bufs


def test_training_step_vs_fp32_oracle_512():
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    model, sd = make_model(seed=0)
    img, t, w = unet_ref.synthetic_batch(2, size=512, seed=1234, device="cuda")
    ref_logits, ref_loss, ref_grads, ref_bufs = oracle_step(sd, img, t, w)

    model.train()
    logits = model(img)
    assert logits.shape == (2, 2, 324, 324) and logits.dtype == torch.float32
    loss = WeightedCrossEntropyLoss()(logits, t, w)
    loss.backward()
    torch.cuda.synchronize()

    e_logits = rel_l2(logits, ref_logits)
    e_loss = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    agree = float(((logits[:, 1] > logits[:, 0]) == (ref_logits[:, 1] > ref_logits[:, 0]))
                  .float().mean())
    cos = {}
    for name, p in model.named_parameters():
        g = p.grad
        assert g is not None and torch.isfinite(g).all(), name
        if name.endswith(ZERO_GRAD_BIASES):
            wname = name.replace(".bias", ".weight")
            assert float(g.norm()) <= 1e-3 * float(ref_grads[wname].norm()) + 1e-12, name
            continue
        cos[name] = cosine(g, ref_grads[name])
    vals = np.array(list(cos.values()))
    print(f"\n[T1 512^2 N=2] logits rel-L2 {e_logits:.3e}  loss rel {e_loss:.3e}  mask agree "
          f"{agree:.5f}  grad cos min {vals.min():.4f} median {np.median(vals):.4f}  "
          f">=0.999: {(vals >= 0.999).sum()}/{len(vals)}")
    worst = sorted(cos.items(), key=lambda kv: kv[1])[:5]
    print("   worst:", ", ".join(f"{k}={v:.3f}" for k, v in worst))
    assert e_logits < 2e-2
    assert e_loss < 2e-3
    assert agree > 0.99
    # head / last decoder block are short paths from the loss: must be essentially exact
    for name in ("outc.conv.weight", "outc.conv.bias", "up4.conv.double_conv.4.weight",
                 "up4.conv.double_conv.3.weight"):
        assert cos[name] > 0.999, (name, cos[name])
    # measured: min 0.889, median 0.953 (torch.autocast(bf16) on the reference: 0.58 / 0.82)
    assert vals.min() > 0.85 and np.median(vals) > 0.94
    # BN buffers follow torch semantics (momentum 0.1, unbiased running_var, counter += 1)
    new = model.state_dict()
    for k, v in ref_bufs.items():
        if k.endswith("num_batches_tracked"):
            assert int(new[k]) == int(v), k
        else:
            assert rel_l2(new[k], v) < 2e-2, (k, rel_l2(new[k], v))


def test_training_step_vs_bf16_rounding_oracle_512():
    """T2: against the reference arithmetic with bf16 rounding at the operand points of
    a bf16 implementation (conv inputs / weights except the first conv, conv-output gradients). The
    result must be at least as close to this oracle as to the fp32 one; far outside the band means
    a kernel bug, not number-format drift."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    model, sd = make_model(seed=0)
    img, t, w = unet_ref.synthetic_batch(2, size=512, seed=1234, device="cuda")
    ref_logits, ref_loss, ref_grads, _ = oracle_step(sd, img, t, w, emulate_bf16=True)
    fp_logits, _, fp_grads, _ = oracle_step(sd, img, t, w)
    model.train()
    logits = model(img)
    loss = WeightedCrossEntropyLoss()(logits, t, w)
    loss.backward()
    torch.cuda.synchronize()
    e_logits = rel_l2(logits, ref_logits)
    e_loss = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    agree = float(((logits[:, 1] > logits[:, 0]) == (ref_logits[:, 1] > ref_logits[:, 0]))
                  .float().mean())
    cos, cos_fp = {}, {}
    for name, p in model.named_parameters():
        if name.endswith(ZERO_GRAD_BIASES):
            continue
        cos[name] = cosine(p.grad, ref_grads[name])
        cos_fp[name] = cosine(p.grad, fp_grads[name])
    vals, vals_fp = np.array(list(cos.values())), np.array(list(cos_fp.values()))
    print(f"\n[T2 512^2 N=2] vs bf16-rounding oracle: logits rel-L2 {e_logits:.3e}  loss rel "
          f"{e_loss:.3e}  mask agree {agree:.5f}  grad cos min {vals.min():.4f} median "
          f"{np.median(vals):.4f}  >=0.999: {(vals >= 0.999).sum()}/{len(vals)}   "
          f"(vs fp32: logits {rel_l2(logits, fp_logits):.3e}, cos min {vals_fp.min():.4f} median "
          f"{np.median(vals_fp):.4f})")
    # Measured: logits 1.39e-2 (vs 1.46e-2 against fp32), masks 99.60 %, cos min 0.895 / median 0.958.
    # The library has one rounding point more than this oracle (the conv output y is stored in bf16
    # before BatchNorm), and runs with identical rounding points already differ by ~5e-3 through
    # ReLU / max-pool mask flips, so the band is "no worse than against fp32".
    assert e_logits < 1.8e-2 and e_loss < 1e-3 and agree > 0.995
    assert e_logits <= 1.02 * rel_l2(logits, fp_logits)
    assert vals.min() > 0.87 and np.median(vals) > 0.95
    assert np.median(vals) >= np.median(vals_fp) - 1e-3   # closer to the same-rounding oracle


@pytest.mark.parametrize("name", ["train_n2_s188", "train_n1_s220", "eval_n1_s252"])
def test_against_reference_golden(name):
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    blob = np.load(os.path.join(GOLD, "unet_golden.npz"))
    c = {k[len(name) + 1:]: blob[k] for k in blob.files if k.startswith(name + "/")}
    n, size, sw, sx, training = [int(v) for v in c["meta"]]
    model, _ = make_model(seed=sw)
    img, t, w = unet_ref.synthetic_batch(n, size=size, seed=sx, device="cuda")
    ref_logits = torch.from_numpy(c["logits"]).cuda()
    if training:
        model.train()
        logits = model(img)
        loss = WeightedCrossEntropyLoss()(logits, t, w)
        loss.backward()
        torch.cuda.synchronize()
        # tiny maps (4x4 logits, 32-sample batch statistics at the bottleneck) amplify bf16 noise
        assert abs(float(loss) - float(c["loss"])) / float(c["loss"]) < 2e-2
        assert rel_l2(logits, ref_logits) < 8e-2
        g = model.outc.conv.bias.grad.cpu().double()
        assert abs(float(g.norm()) - float(c["gnorm/outc.conv.bias"])) < 5e-2 * float(
            c["gnorm/outc.conv.bias"])
    else:
        gen = torch.Generator().manual_seed(99)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
        model.eval()
        with torch.no_grad():
            logits = model(img)
        torch.cuda.synchronize()
        assert rel_l2(logits, ref_logits) < 2e-2


def test_eval_forward_and_fused_mask_vs_oracle():
    model, sd = make_model(seed=1)
    gen = torch.Generator().manual_seed(7)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = (torch.randn(nf, generator=gen) * 0.1).cuda()
        sd[k.replace("running_mean", "running_var")] = (0.5 + torch.rand(nf, generator=gen)).cuda()
    model.load_state_dict(sd)
    model.eval()
    img, _, _ = unet_ref.synthetic_batch(1, size=572, seed=5, device="cuda")
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, img, training=False)
        logits, mask = model.predict_mask(img)
        again = model(img)
    torch.cuda.synchronize()
    assert logits.shape == (1, 2, 388, 388)
    assert torch.equal(logits, again)                      # deterministic
    assert torch.equal(mask > 0, logits[:, 1] > logits[:, 0])
    e = rel_l2(logits, ref)
    agree = float(((logits[:, 1] > logits[:, 0]) == (ref[:, 1] > ref[:, 0])).float().mean())
    margin = (ref[:, 1] - ref[:, 0]).abs() > 0.05
    agree_conf = float(((logits[:, 1] > logits[:, 0]) == (ref[:, 1] > ref[:, 0]))[margin]
                       .float().mean())
    print(f"\n[eval 572^2] logits rel-L2 {e:.3e}  mask agree {agree:.5f}  (margin>0.05: "
          f"{agree_conf:.5f}, {float(margin.float().mean()):.3f} of pixels)")
    assert e < 2e-2
    assert agree_conf >= 0.999


@pytest.mark.parametrize("size", [230, 231])
def test_eval_forward_with_odd_feature_maps_vs_oracle(size):
    """Eval path on sizes whose level-0 / level-1 maps are odd (230: 226 -> 113 -> 109 -> 54, 231: 227 -> 113; floor-mode
    pooling drops the last row / column): the 2x2 max-pool fused into the two-row conv
    epilogue must floor exactly like nn.MaxPool2d(2)."""
    model, sd = make_model(seed=5)
    gen = torch.Generator().manual_seed(11)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = (torch.randn(nf, generator=gen) * 0.1).cuda()
        sd[k.replace("running_mean", "running_var")] = (0.5 + torch.rand(nf, generator=gen)).cuda()
    model.load_state_dict(sd)
    model.eval()
    img, _, _ = unet_ref.synthetic_batch(2, size=size, seed=8, device="cuda")
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, img, training=False)
        logits, mask = model.predict_mask(img)
    torch.cuda.synchronize()
    assert logits.shape == ref.shape
    e = rel_l2(logits, ref)
    print(f"\n[eval {size}^2, odd maps] logits rel-L2 {e:.3e}")
    assert e < 2e-2
    assert torch.equal(mask > 0, logits[:, 1] > logits[:, 0])


def test_reference_style_training_loop_runs_and_learns():
    """The hot loop of scripts/train.py:104-135 with the drop-in classes (SGD lr 1e-4 -> 1e-2 here
    so that three steps show progress)."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    model, _ = make_model(seed=2)
    model.train()
    crit = WeightedCrossEntropyLoss()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2, momentum=0.9)
    g = torch.Generator().manual_seed(0)
    images = 0.4 + 0.2 * torch.rand(2, 1, 252, 252, generator=g)
    masks = (images > 0.5).long()            # learnable from a single pixel
    wmaps = torch.full((2, 1, 252, 252), 12.0)
    images, masks, wmaps = images.cuda(), masks.cuda(), wmaps.cuda()
    losses = []
    for _ in range(4):
        opt.zero_grad()
        out = model(images)
        th, tw = out.shape[2:]
        s = (252 - th) // 2
        tgt = masks[:, :, s:s + th, s:s + tw].squeeze(1)
        wm = wmaps[:, :, s:s + th, s:s + tw].squeeze(1)
        loss = crit(out, tgt, wm)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0]
    assert int(model.inc.double_conv[1].num_batches_tracked) == 4


def test_sgd_trajectory_tracks_fp32_oracle():
    """Six optimizer steps with the reference's SGD(lr=1e-4, momentum=0.99) (scripts/train.py:97):
    the bf16 path must follow the fp32 oracle's loss curve (also proves the packed bf16 weight
    caches are refreshed after every optimizer.step())."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    model, sd = make_model(seed=0)
    model.train()
    img, t, w = unet_ref.synthetic_batch(2, size=252, seed=21, device="cuda")
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    opt_ref = torch.optim.SGD(list(params.values()), lr=1e-4, momentum=0.99)
    opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.99)
    crit = WeightedCrossEntropyLoss()
    ours, ref = [], []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(img), t, w)
        loss.backward()
        opt.step()
        ours.append(float(loss.detach()))
        opt_ref.zero_grad(set_to_none=True)
        bufs = {}
        rl = unet_ref.weighted_cross_entropy(
            unet_ref.unet_forward(full, img, training=True, buffers_out=bufs), t, w)
        rl.backward()
        opt_ref.step()
        for k, v in bufs.items():
            full[k] = v
        ref.append(float(rl.detach()))
    print("\n[SGD trajectory] ours", [f"{v:.3f}" for v in ours], " oracle", [f"{v:.3f}" for v in ref])
    assert ref[-1] < ref[0]
    for a, b in zip(ours, ref):
        assert abs(a - b) / b < 0.08, (ours, ref)


def test_fused_sgd_matches_torch_sgd():
    """FusedSGD (one multi-tensor kernel: SGD update + bf16 operand refresh) against
    torch.optim.SGD with the same hyper-parameters, including momentum, weight decay and Nesterov."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
    from unet_segmentation_b200.optim import FusedSGD

    img, t, w = unet_ref.synthetic_batch(2, size=252, seed=31, device="cuda")
    crit = WeightedCrossEntropyLoss()
    for kw in (dict(lr=1e-3, momentum=0.9, weight_decay=1e-4),
               dict(lr=1e-3, momentum=0.9, nesterov=True),
               dict(lr=1e-2)):
        m1, _ = make_model(seed=0)
        m2, _ = make_model(seed=0)
        m1.train(); m2.train()
        o1 = torch.optim.SGD(m1.parameters(), **kw)
        o2 = FusedSGD(m2, **kw)
        for it in range(3):
            for m, o in ((m1, o1), (m2, o2)):
                o.zero_grad(set_to_none=True)
                crit(m(img), t, w).backward()
                o.step()
            torch.cuda.synchronize()
            # step 1 starts from identical weights and identical (deterministic) gradients: the two
            # optimizers may differ only by fp32 rounding (FMA contraction). Later steps amplify that
            # through bf16 rounding / ReLU-mask flips, so the bar is loose there.
            tol = 2e-6 if it == 0 else 0.15
            for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
                assert rel_l2(p2, p1) < tol, (kw, it, n1, rel_l2(p2, p1))
            if kw.get("momentum"):
                b1 = o1.state[m1.outc.conv.weight]["momentum_buffer"]
                b2 = o2.state[m2.outc.conv.weight]["momentum_buffer"]
                assert rel_l2(b2, b1) < (1e-6 if it == 0 else 0.15)
        # the operand caches written by the fused kernel == a fresh re-pack of the same weights
        m3, _ = make_model(seed=0)
        m3.load_state_dict(m2.state_dict())
        m3.train()
        with torch.no_grad():
            m2.eval(); m3.eval()
            assert torch.equal(m2(img), m3(img))          # eval plan re-packed lazily (epoch bump)
        m2.train(); m3.train()
        l2 = m2(img); l3 = m3(img)
        torch.cuda.synchronize()
        assert torch.equal(l2, l3), kw                      # training plan uses the fused-written caches


def test_input_gradient_request_fails_loudly():
    """The library never computes d loss / d image (the first layer needs no data
    gradient); asking for it must raise instead of leaving x.grad silently empty."""
    model, _ = make_model(seed=0)
    model.train()
    img, _, _ = unet_ref.synthetic_batch(1, size=188, seed=1, device="cuda")
    with pytest.raises(RuntimeError, match="gradient with respect to the input"):
        model(img.clone().requires_grad_(True))
    with torch.no_grad():                      # fine without autograd, and in eval mode
        assert model(img.clone().requires_grad_(True)).shape == (1, 2, 4, 4)
