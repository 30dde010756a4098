"""End-to-end parity of the drop-in UNet for the configurations beside the reference default:
BASELINE config 5's wide net (base 128), other depths, multi-channel inputs, other class counts
(the reference's stale callers build ``UNet(n_channels=1, n_classes=1)``, scripts/inference.py:39).
Same oracle and metrics as tests/test_unet_gpu.py (tier T1); sizes small enough for seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import unet_ref  # noqa: E402

ZERO_GRAD_BIASES = (".double_conv.0.bias", ".double_conv.3.bias", ".up.bias")


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def build(n_channels, n_classes, base, levels, seed, bilinear=False):
    from unet_segmentation_b200.unet import UNet

    sd = unet_ref.make_state_dict(n_channels, n_classes, seed=seed, base=base, levels=levels,
                                  bilinear=bilinear)
    m = UNet(n_channels, n_classes, bilinear, base_channels=base, levels=levels)
    m.load_state_dict(sd)                      # same keys / shapes as the reference-style tree
    return m.cuda(), {k: v.cuda() for k, v in sd.items()}


def oracle_step(sd, img, t, w, levels):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    logits = unet_ref.unet_forward(full, img, training=True, levels=levels, buffers_out={})
    loss = unet_ref.weighted_cross_entropy(logits, t, w)
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad for k, p in params.items()}


@pytest.mark.parametrize("n_channels,n_classes,base,levels,n,size,bilinear", [
    (1, 2, 128, 5, 2, 252, False),      # BASELINE configs[4]: wide U-Net (base 128, depth 5)
    (1, 2, 64, 4, 2, 196, False),       # shallower net
    (3, 3, 64, 3, 2, 132, False),       # RGB input (generic first conv), 3 classes, depth 3
    (1, 2, 64, 6, 2, 444, False),       # deeper net (2048 channels at the bottleneck)
    (1, 2, 64, 5, 2, 252, True),        # UNet(1, 2, bilinear=True): nn.Upsample branch (:40-43)
    (1, 2, 64, 3, 2, 132, True),
])
def test_training_step_other_configurations(n_channels, n_classes, base, levels, n, size, bilinear):
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    model, sd = build(n_channels, n_classes, base, levels, seed=2, bilinear=bilinear)
    img, t, w = unet_ref.synthetic_batch(n, size=size, seed=31, levels=levels, device="cuda")
    if n_channels > 1:
        g = torch.Generator(device="cuda").manual_seed(5)
        img = (img + 0.1 * torch.rand(n, n_channels, size, size, device="cuda", generator=g)).contiguous()
    ref_logits, ref_loss, ref_grads = oracle_step(sd, img, t, w, levels)
    model.train()
    logits = model(img)
    assert logits.shape == ref_logits.shape and logits.dtype == torch.float32
    loss = WeightedCrossEntropyLoss()(logits, t, w)
    loss.backward()
    torch.cuda.synchronize()
    e_logits = rel_l2(logits, ref_logits)
    e_loss = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    cos = {}
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        if name.endswith(ZERO_GRAD_BIASES):
            continue
        cos[name] = cosine(p.grad, ref_grads[name])
    vals = np.array(list(cos.values()))
    print(f"\n[C={n_channels} K={n_classes} base={base} levels={levels} bilinear={bilinear} {n}x{size}^2] logits rel-L2 "
          f"{e_logits:.3e}  loss rel {e_loss:.3e}  grad cos min {vals.min():.4f} median "
          f"{np.median(vals):.4f}  out {tuple(logits.shape)}")
    # measured: logits 1.1e-2 … 1.7e-2, loss 2e-6 … 4e-4, cosine median 0.91 … 0.99 (the deeper the
    # net, the more ReLU / pool mask flips); the head is a short path from the loss.
    # bilinear=True at random init is 4-5x more sensitive to bf16 in TRAIN mode, for any
    # implementation: rounding only the conv operands of the fp32 oracle to bf16 already moves its
    # logits by 4.3e-2, plus the bf16-stored pre-BN output 5.8e-2 (CPU control) — the
    # library measures 5.8e-2 at every size; its eval-mode forward is at 4.3e-3 like the default net
    # (scripts_dev/bilinear_diag.py). Smooth up-sampled inputs carry a large per-channel mean that
    # the following BatchNorm removes, which amplifies the rounding (the mechanism of SURVEY F4).
    assert e_logits < (8e-2 if bilinear else 3e-2) and e_loss < (1e-2 if bilinear else 5e-3)
    assert cos["outc.conv.weight"] > 0.995 and cos["outc.conv.bias"] > 0.995
    assert np.median(vals) > 0.8


def test_single_class_eval_forward():
    """UNet(1, 1) as built by scripts/inference.py:39 / scripts/predict1.py:32 (logit -> sigmoid)."""
    model, sd = build(1, 1, 64, 5, seed=4)
    gen = torch.Generator().manual_seed(7)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = (torch.randn(nf, generator=gen) * 0.1).cuda()
        sd[k.replace("running_mean", "running_var")] = (0.5 + torch.rand(nf, generator=gen)).cuda()
    model.load_state_dict(sd)
    model.eval()
    img, _, _ = unet_ref.synthetic_batch(1, size=316, seed=5, device="cuda")
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, img, training=False)
        logits = model(img)
    torch.cuda.synchronize()
    assert logits.shape == ref.shape == (1, 1, 132, 132)
    assert rel_l2(logits, ref) < 2e-2
    p, pr = torch.sigmoid(logits), torch.sigmoid(ref)
    conf = (pr - 0.5).abs() > 0.02
    assert float(((p > 0.5) == (pr > 0.5))[conf].float().mean()) >= 0.999


def test_bilinear_unet_against_reference_golden_and_sgd_steps():
    """UNet(1, 2, bilinear=True) against the golden vector recorded from the reference itself, then
    three optimizer steps (FusedSGD must cope with a parameter table without up.weight / up.bias)."""
    import os

    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
    from unet_segmentation_b200.optim import FusedSGD

    blob = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_bilinear_golden.npz"))
    name = "train_n1_s220_bilinear"
    c = {k[len(name) + 1:]: blob[k] for k in blob.files if k.startswith(name + "/")}
    n, size, sw, sx, _ = [int(v) for v in c["meta"]]
    model, _ = build(1, 2, 64, 5, seed=sw, bilinear=True)
    img, t, w = unet_ref.synthetic_batch(n, size=size, seed=sx, device="cuda")
    model.train()
    crit = WeightedCrossEntropyLoss()
    logits = model(img)
    loss = crit(logits, t, w)
    loss.backward()
    torch.cuda.synchronize()
    e_loss = abs(float(loss) - float(c["loss"])) / float(c["loss"])
    e_logits = rel_l2(logits.detach(), torch.from_numpy(c["logits"]).cuda())
    g = model.outc.conv.bias.grad.cpu().double()
    e_g = abs(float(g.norm()) - float(c["gnorm/outc.conv.bias"])) / float(c["gnorm/outc.conv.bias"])
    print(f"\n[bilinear golden 1x220^2] loss rel {e_loss:.3e}  logits rel-L2 {e_logits:.3e}  "
          f"outc.bias grad-norm rel {e_g:.3e}")
    # 36x36 logits from N = 1: bf16 noise of the bilinear net (see above) on tiny batch statistics
    assert e_loss < 5e-3 and e_logits < 0.15 and e_g < 1e-2   # measured 3.7e-4 / 9.2e-2 / 8e-5
    opt = FusedSGD(model, lr=1e-3, momentum=0.9)
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(img), t, w)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    model.eval()
    with torch.no_grad():
        z = model(img)
    assert z.shape == logits.shape and torch.isfinite(z).all()
