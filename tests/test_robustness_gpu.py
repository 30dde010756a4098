"""Loud failures and module-protocol behaviour of the drop-in classes on the GPU (round-1 advisor
findings): out-of-range loss targets, the single-class (sigmoid) mask, BatchNorm hyper-parameters,
stale operand caches after `.data` writes, deepcopy / pickling after a forward pass."""
import copy
import io

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import unet_ref  # noqa: E402


def _model(n_classes=2, seed=0):
    from unet_segmentation_b200.unet import UNet

    m = UNet(1, n_classes)
    m.load_state_dict(unet_ref.make_state_dict(1, n_classes, seed=seed))
    return m.cuda()


def test_wce_out_of_range_target_is_loud():
    """The reference's nn.CrossEntropyLoss faults on a target outside [0, C) (utils/losses.py:49 — e.g.
    an un-binarised 255 mask). Here: NaN loss, a DEFINED (zero) gradient at the bad pixels, finite
    gradients elsewhere, and a RuntimeError — at once with check_targets, else from a later call."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss

    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(2, 2, 40, 36, device="cuda", generator=g, requires_grad=True)
    t = (torch.rand(2, 40, 36, device="cuda", generator=g) > 0.5).long()
    w = torch.full((2, 40, 36), 12.0, device="cuda")
    bad = t.clone()
    bad[0, 3, 5] = 255
    bad[1, 7, 1] = -3
    crit = WeightedCrossEntropyLoss()
    crit.check_targets = True
    with pytest.raises(RuntimeError, match="outside"):
        crit(logits, bad, w)
    # default mode: the call itself returns (NaN loss), the error arrives with a later call
    crit = WeightedCrossEntropyLoss()
    loss = crit(logits, bad, w)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isnan(loss)
    # gradient: NaN * finite upstream poisons everything through the loss scalar by design (dz is scaled
    # by d loss = 1 here, so the stored dz itself must be finite and zero at the two bad pixels)
    dz = logits.grad
    assert torch.isfinite(dz).all()
    assert float(dz[0, :, 3, 5].abs().max()) == 0.0 and float(dz[1, :, 7, 1].abs().max()) == 0.0
    assert float(dz.abs().sum()) > 0
    with pytest.raises(RuntimeError, match="outside"):
        crit(logits.detach(), t, w)
    # after raising once the module is usable again; ignore_index (-100) stays legal
    ok = t.clone()
    ok[0, 0, 0] = -100
    assert torch.isfinite(crit(logits.detach(), ok, w))
    crit.check_targets = True
    assert torch.isfinite(crit(logits.detach(), ok, w))


def test_single_class_mask_is_sigmoid_threshold():
    """UNet(1, 1) as built by scripts/inference.py:39: mask = sigmoid(logit) > 0.5 = logit > 0
    (scripts/inference.py:85), from the fused conv epilogue AND from the stand-alone head kernels."""
    from unet_segmentation_b200 import ops, tiling

    model = _model(n_classes=1, seed=4)
    gen = torch.Generator().manual_seed(7)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
    model.eval()
    img, _, _ = unet_ref.synthetic_batch(1, size=252, seed=5, device="cuda")
    with torch.no_grad():       # centre the logits so that both classes occur
        model.outc.conv.bias.sub_(model(img).median())
    logits, mask = model.predict_mask(img)
    torch.cuda.synchronize()
    assert logits.shape == (1, 1, 68, 68)
    frac = float((mask > 0).float().mean())
    assert 0.05 < frac < 0.95                      # both classes present: the old code returned all zeros
    assert torch.equal(mask > 0, logits[:, 0] > 0)
    # stand-alone head kernels (vectorised K = 64 and generic K)
    for k in (64, 72):
        a = torch.randn(1, 9, 11, k, device="cuda").to(torch.bfloat16)
        wt = torch.randn(1, k, device="cuda")
        lg, mk = ops.head_forward(a, wt, None, want_mask=True)
        assert torch.equal(mk > 0, lg[:, 0] > 0) and 0 < int((mk > 0).sum()) < mk.numel()
    # and through overlap-tile inference
    big = (0.4 + 0.2 * torch.rand(300, 300, generator=torch.Generator().manual_seed(1))).cuda()
    full, full_logits = tiling.overlap_tile_predict(model, big, tile_in=252, batch_tiles=2,
                                                    return_logits=True)
    assert torch.equal(full > 0, full_logits[0] > 0)


def test_batchnorm_hyperparameters_are_honoured_or_rejected():
    model = _model()
    img, _, _ = unet_ref.synthetic_batch(2, size=188, seed=3, device="cuda")
    ref = _model()
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.3
    model.train(); ref.train()
    with torch.no_grad():
        a, b = model(img), ref(img)
    torch.cuda.synchronize()
    assert torch.equal(a, b)                       # momentum only steers the running statistics
    for (k, v), (_, r) in zip(model.state_dict().items(), ref.state_dict().items()):
        if k.endswith("running_mean"):             # started at 0: new = momentum * batch mean
            assert torch.allclose(v, 3.0 * r, rtol=1e-5, atol=1e-7), k
        if k.endswith("running_var"):              # started at 1: new - 1 = momentum * (var - 1)
            assert torch.allclose(v - 1, 3.0 * (r - 1), rtol=1e-4, atol=1e-6), k
    # eps enters the normalisation itself
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps = 1e-1
    with torch.no_grad():
        c = model(img)
    assert not torch.equal(a, c)
    # unsupported settings fail loudly instead of running with default semantics
    bad = _model().train()
    bad.inc.double_conv[1].momentum = None
    with pytest.raises(RuntimeError, match="momentum=None"):
        bad(img)
    bad = _model().train()
    bad.down2.maxpool_conv[1].double_conv[4].eval()          # frozen-BN fine-tuning
    with pytest.raises(RuntimeError, match="frozen"):
        bad(img)
    bad = _model().train()
    bad.up1.conv.double_conv[1].track_running_stats = False
    with pytest.raises(RuntimeError, match="track_running_stats"):
        bad(img)


def test_data_writes_need_invalidate_weights_and_get_it():
    model = _model().eval()
    img, _, _ = unet_ref.synthetic_batch(1, size=188, seed=9, device="cuda")
    with torch.no_grad():
        before = model(img)
        model.down1.maxpool_conv[1].double_conv[0].weight.data.mul_(1.5)   # no version bump
        model.invalidate_weights()
        after = model(img)
        # an ordinary in-place op bumps the version counter: picked up without help
        model.down1.maxpool_conv[1].double_conv[0].weight.div_(1.5)
        back = model(img)
    torch.cuda.synchronize()
    assert not torch.equal(before, after)
    assert float((back - before).abs().max()) <= 2e-2 * float(before.abs().max())


def test_deepcopy_and_pickle_after_forward():
    model = _model().train()
    img, t, w = unet_ref.synthetic_batch(1, size=188, seed=2, device="cuda")
    out = model(img)
    assert model._plans
    twin = copy.deepcopy(model)                    # plans (ctypes handles of device arenas) are dropped
    assert not twin._plans and twin is not model
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)
    for m in (twin, loaded):
        for (k, a), (_, b) in zip(m.state_dict().items(), model.state_dict().items()):
            assert torch.equal(a, b), k
    # BN buffers moved on in `model` after the copy was taken only if it runs again: run all three
    model.eval(); twin.eval(); loaded.eval()
    with torch.no_grad():
        r0, r1, r2 = model(img), twin(img), loaded(img)
    torch.cuda.synchronize()
    assert torch.equal(r0, r1) and torch.equal(r0, r2)
    assert out.shape == (1, 2, 4, 4)
