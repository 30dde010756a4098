"""Teacher-forced per-layer parity (tier T0): every CUDA kernel, called through the C
ABI, against the single torch op it replaces, in fp32 on the same bf16-rounded inputs. Only the
accumulation order differs, so the bars are tight: rel-L2 <= 4e-3 for bf16 outputs (one bf16
rounding), <= 1e-3 and cosine >= 0.999 for fp32 gradients."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3
F32_TOL = 1e-3


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def bf(x):
    return x.to(torch.bfloat16).float()


@pytest.fixture(scope="module")
def ops():
    from unet_segmentation_b200 import _lib, ops as _ops

    _lib.require_cuda()
    return _ops


def rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("n,h,w,ci,co", [
    (1, 12, 12, 64, 64),       # less than one M tile
    (2, 37, 29, 64, 128),      # odd sizes, tiles straddle rows and images
    (2, 23, 23, 128, 256),
    (1, 19, 21, 256, 512),     # two N tiles
    (3, 30, 30, 64, 64),
])
def test_conv3x3_forward(ops, n, h, w, ci, co):
    x = bf(rand(n, ci, h, w))
    wt = bf(rand(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=1))
    b = rand(co, seed=2)
    wf, _ = ops.pack_conv3x3(wt)
    y, _, _ = ops.conv3x3_forward(ops.nhwc(x), None, wf, b, epilogue=1)
    ref = F.conv2d(x, wt, b)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(y), ref) < BF16_TOL


def test_conv3x3_forward_stats_and_bn(ops):
    n, h, w, ci, co = 2, 41, 33, 64, 128
    x = bf(rand(n, ci, h, w)) + 0.3
    wt = bf(rand(co, ci, 3, 3, scale=0.05, seed=1))
    b = rand(co, seed=2)
    gamma, beta = rand(co, seed=3) + 1.5, rand(co, seed=4)
    rm, rv = torch.zeros(co, device="cuda"), torch.ones(co, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    wf, _ = ops.pack_conv3x3(wt)
    y, stats, info = ops.conv3x3_forward(ops.nhwc(x), None, wf, b, epilogue=0)
    scale, shift, mean, rstd = ops.bn_finalize(stats, info, gamma, beta, rm, rv, nbt)
    a, p, _ = ops.bn_apply_relu(y, scale, shift, pool=True)
    ref_y = F.conv2d(x, wt, b)
    bn = torch.nn.BatchNorm2d(co).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    ref_a = F.relu(bn(ref_y))
    torch.cuda.synchronize()
    assert rel_l2(mean, ref_y.mean((0, 2, 3))) < 1e-3
    assert rel_l2(1 / rstd ** 2, ref_y.var((0, 2, 3), unbiased=False) + 1e-5) < 2e-3
    assert rel_l2(rm, bn.running_mean) < 1e-3 and rel_l2(rv, bn.running_var) < 2e-3
    assert int(nbt) == 1
    assert rel_l2(ops.nchw(a), ref_a) < 6e-3   # two bf16 roundings (y and a)
    # pooling is exact given a
    assert torch.equal(ops.nchw(p), F.max_pool2d(ops.nchw(a), 2))


@pytest.mark.parametrize("cs,cu,co", [(64, 64, 128), (64, 128, 64), (128, 256, 128), (256, 512, 256)])
def test_conv3x3_concat_sources(ops, cs, cu, co):
    """Zero-copy crop + concat: src0 = centre crop of a larger skip tensor, src1 = up-sampled.
    Unequal channel counts are the bilinear=True decoder (C_skip + 2 C_skip -> C_skip)."""
    n, hs, ws_, h, w = 2, 30, 28, 20, 18
    skip = bf(rand(n, cs, hs, ws_))
    up = bf(rand(n, cu, h, w, seed=5))
    ch, cw = (hs - h) // 2, (ws_ - w) // 2
    wt = bf(rand(co, cs + cu, 3, 3, scale=(2.0 / (9 * (cs + cu))) ** 0.5, seed=1))
    wf, wd = ops.pack_conv3x3(wt)
    skip_nhwc = ops.nhwc(skip)
    y, _, _ = ops.conv3x3_forward(skip_nhwc[:, ch:ch + h, cw:cw + w, :], ops.nhwc(up), wf, None)
    ref = F.conv2d(torch.cat([skip[:, :, ch:ch + h, cw:cw + w], up], 1), wt)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(y), ref) < BF16_TOL
    # weight gradient over the same two sources
    dy = bf(rand(n, co, h - 2, w - 2, seed=7))
    dw = ops.conv3x3_wgrad(skip_nhwc[:, ch:ch + h, cw:cw + w, :], ops.nhwc(up), ops.nhwc(dy))
    ref_dw = torch.nn.grad.conv2d_weight(
        torch.cat([skip[:, :, ch:ch + h, cw:cw + w], up], 1), wt.shape, dy)
    torch.cuda.synchronize()
    assert rel_l2(dw, ref_dw) < F32_TOL and cosine(dw, ref_dw) > 0.9999
    # data gradient of the concatenated input: one tensor whose channel ranges are the two sources'
    dx = ops.conv3x3_dgrad(ops.nhwc(dy), wd)
    ref_dx = torch.nn.grad.conv2d_input((n, cs + cu, h, w), wt, dy)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(dx), ref_dx) < BF16_TOL


def test_conv3x3_eval_epilogue(ops):
    n, h, w, ci, co = 1, 25, 25, 64, 64
    x = bf(rand(n, ci, h, w))
    wt = bf(rand(co, ci, 3, 3, scale=0.06, seed=1))
    scale, shift = rand(co, seed=2).abs() + 0.5, rand(co, seed=3)
    wf, _ = ops.pack_conv3x3(wt)
    y, _, _ = ops.conv3x3_forward(ops.nhwc(x), None, wf, None, epilogue=2, scale=scale, shift=shift)
    ref = F.relu(F.conv2d(x, wt) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(y), ref) < BF16_TOL


@pytest.mark.parametrize("n,h,w,ci,co", [(1, 10, 10, 64, 64), (2, 21, 17, 128, 64), (2, 14, 14, 256, 512)])
def test_conv3x3_dgrad(ops, n, h, w, ci, co):
    dy = bf(rand(n, co, h, w))
    wt = bf(rand(co, ci, 3, 3, scale=0.05, seed=1))
    _, wd = ops.pack_conv3x3(wt)
    dx = ops.conv3x3_dgrad(ops.nhwc(dy), wd)
    ref = torch.nn.grad.conv2d_input((n, ci, h + 2, w + 2), wt, dy)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(dx), ref) < BF16_TOL


@pytest.mark.parametrize("n,h,w,ci,co", [(1, 12, 12, 64, 64), (2, 35, 27, 64, 128), (2, 18, 18, 256, 256),
                                         (1, 50, 50, 128, 64)])
def test_conv3x3_wgrad(ops, n, h, w, ci, co):
    x = bf(rand(n, ci, h, w))
    dy = bf(rand(n, co, h - 2, w - 2, seed=3))
    dw = ops.conv3x3_wgrad(ops.nhwc(x), None, ops.nhwc(dy))
    ref = torch.nn.grad.conv2d_weight(x, (co, ci, 3, 3), dy)
    torch.cuda.synchronize()
    assert rel_l2(dw, ref) < F32_TOL and cosine(dw, ref) > 0.9999


@pytest.mark.parametrize("n,h,w,ci", [(1, 6, 6, 128), (2, 11, 9, 256), (1, 24, 24, 1024)])
def test_conv_transpose(ops, n, h, w, ci):
    co = ci // 2
    x = bf(rand(n, ci, h, w))
    wt = bf(rand(ci, co, 2, 2, scale=0.05, seed=1))
    b = rand(co, seed=2)
    wf, wb, b4 = ops.pack_convT(wt, b)
    # write into the second channel range of a wider (concat-like) buffer
    buf = torch.zeros(n, 2 * h, 2 * w, 2 * co, dtype=torch.bfloat16, device="cuda")
    ops.convT_forward(ops.nhwc(x), wf, b4, buf[..., co:])
    ref = F.conv_transpose2d(x, wt, b, stride=2)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(buf[..., co:]), ref) < BF16_TOL
    assert float(buf[..., :co].float().abs().max()) == 0.0
    # backward: gradient arrives as a channel slice of d(concat)
    dbuf = ops.nhwc(bf(rand(n, 2 * co, 2 * h, 2 * w, seed=4)))
    dup = dbuf[..., co:]
    dx = ops.convT_dgrad(dup, wb)
    dw = ops.convT_wgrad(dup, ops.nhwc(x))
    dup_nchw = ops.nchw(dup)
    ref_dx = F.conv2d(dup_nchw, wt, stride=2)
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, None, stride=2).backward(dup_nchw)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(dx), xr.grad) < BF16_TOL
    assert rel_l2(ops.nchw(dx), ref_dx) < BF16_TOL
    assert rel_l2(dw, wr.grad) < F32_TOL and cosine(dw, wr.grad) > 0.9999


def _bn_relu_ref(y, gamma, beta, g_fn):
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a = F.relu(F.batch_norm(yr, None, None, gr, br, True, 0.1, 1e-5))
    g_fn(a)
    return yr.grad, gr.grad, br.grad


def test_bn_relu_backward_direct(ops):
    n, h, w, c = 2, 19, 23, 128
    y = bf(rand(n, c, h, w) * 1.3 + 0.2)
    gamma, beta = rand(c, seed=1) + 1.2, rand(c, seed=2) * 0.3
    g = bf(rand(n, c, h, w, seed=3))
    mean = y.mean((0, 2, 3)); var = y.var((0, 2, 3), unbiased=False)
    rstd = (var + 1e-5).rsqrt(); scale = gamma * rstd; shift = beta - mean * scale
    dy, dgamma, dbeta = ops.bn_relu_backward(ops.nhwc(y), scale, shift, mean, rstd, g=ops.nhwc(g))
    ry, rg, rb = _bn_relu_ref(y, gamma, beta, lambda a: a.backward(g))
    torch.cuda.synchronize()
    assert rel_l2(dgamma, rg) < F32_TOL and rel_l2(dbeta, rb) < F32_TOL
    assert rel_l2(ops.nchw(dy), ry) < BF16_TOL


@pytest.mark.parametrize("h,w", [(20, 20), (21, 19)])
def test_bn_relu_backward_pool_skip(ops, h, w):
    """Upstream = max-pool backward (first arg-max) + centre-cropped skip gradient (channel slice)."""
    n, c, th, tw = 2, 64, 8, 10
    y = bf(rand(n, c, h, w) + 0.1)
    gamma, beta = rand(c, seed=1) + 1.0, rand(c, seed=2) * 0.2
    mean = y.mean((0, 2, 3)); var = y.var((0, 2, 3), unbiased=False)
    rstd = (var + 1e-5).rsqrt(); scale = gamma * rstd; shift = beta - mean * scale
    gp = bf(rand(n, c, h // 2, w // 2, seed=3))
    dcat = bf(rand(n, 2 * c, th, tw, seed=4))          # d(concat); first c channels -> skip
    ch, cw = (h - th) // 2, (w - tw) // 2
    dcat_nhwc = ops.nhwc(dcat)
    dy, dgamma, dbeta = ops.bn_relu_backward(ops.nhwc(y), scale, shift, mean, rstd, g=None,
                                             gp=ops.nhwc(gp), gs=dcat_nhwc[..., :c], crop=(ch, cw))
    # same through the per-pixel kernel that uses the arg-max saved by the forward pass
    _, _, amax = ops.bn_apply_relu(ops.nhwc(y), scale, shift, pool=True)
    dy2, dgamma2, dbeta2 = ops.bn_relu_backward(ops.nhwc(y), scale, shift, mean, rstd, g=None,
                                                gp=ops.nhwc(gp), gs=dcat_nhwc[..., :c],
                                                crop=(ch, cw), argmax=amax)
    torch.cuda.synchronize()
    # (the two kernels reduce in different orders, so dy may differ in the last bf16 bit)
    assert rel_l2(dy2.float(), dy.float()) < 2e-3
    assert rel_l2(dgamma2, dgamma) < 1e-4 and rel_l2(dbeta2, dbeta) < 1e-4
    dy_pix, dgamma_pix, dbeta_pix = dy2, dgamma2, dbeta2

    def g_fn(a):
        # forward rounding of a to bf16 decides the arg-max, as in the kernel
        a_r = a + (bf(a.detach()) - a.detach())
        pooled = F.max_pool2d(a_r, 2)
        skip = a_r[:, :, ch:ch + th, cw:cw + tw]
        ((pooled * gp).sum() + (skip * dcat[:, :c]).sum()).backward()

    ry, rg, rb = _bn_relu_ref(y, gamma, beta, g_fn)
    torch.cuda.synchronize()
    assert rel_l2(dgamma, rg) < F32_TOL and rel_l2(dbeta, rb) < F32_TOL
    assert rel_l2(ops.nchw(dy), ry) < BF16_TOL
    assert rel_l2(dgamma_pix, rg) < F32_TOL and rel_l2(dbeta_pix, rb) < F32_TOL
    assert rel_l2(ops.nchw(dy_pix), ry) < BF16_TOL


@pytest.mark.parametrize("ci", [1, 3])
def test_first_conv(ops, ci):
    n, h, w, co = 2, 40, 36, 64
    x = 0.4 + 0.2 * torch.rand(n, ci, h, w, device="cuda",
                               generator=torch.Generator(device="cuda").manual_seed(ci))
    wt = rand(co, ci, 3, 3, scale=0.3, seed=1)
    b = rand(co, seed=2) * 0.1
    gamma, beta = rand(co, seed=3) + 1.0, rand(co, seed=4) * 0.2
    rm, rv = torch.zeros(co, device="cuda"), torch.ones(co, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    a, st = ops.first_conv_forward(x, wt, b, gamma, beta, rm, rv, nbt)
    wr = wt.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn_rm, bn_rv = torch.zeros(co, device="cuda"), torch.ones(co, device="cuda")
    ref_a = F.relu(F.batch_norm(F.conv2d(x, wr, b), bn_rm, bn_rv, gr, br, True, 0.1, 1e-5))
    g = bf(rand(n, co, h - 2, w - 2, seed=5))
    ref_a.backward(g)
    dw, dgamma, dbeta = ops.first_conv_backward(x, wt, b, st, ops.nhwc(g), a)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(a), ref_a) < BF16_TOL
    # near-zero channel means: compare against the scale of the data, not of the mean itself
    assert torch.allclose(rm, bn_rm, rtol=1e-3, atol=1e-5) and rel_l2(rv, bn_rv) < 1e-3
    assert int(nbt) == 1
    assert rel_l2(dgamma, gr.grad) < F32_TOL and rel_l2(dbeta, br.grad) < F32_TOL
    assert rel_l2(dw, wr.grad) < F32_TOL and cosine(dw, wr.grad) > 0.9999


@pytest.mark.parametrize("nc", [1, 2, 3])
def test_head(ops, nc):
    n, h, w, k = 2, 17, 13, 64
    a = bf(rand(n, k, h, w).relu())
    wt, b = rand(nc, k, scale=0.2, seed=1), rand(nc, seed=2)
    logits, mask = ops.head_forward(ops.nhwc(a), wt, b, want_mask=True)
    ar = a.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.conv2d(ar, wr.view(nc, k, 1, 1), br)
    dl = rand(n, nc, h, w, seed=3)
    ref.backward(dl)
    da, dw, db = ops.head_backward(dl, ops.nhwc(a), wt)
    torch.cuda.synchronize()
    assert rel_l2(logits, ref) < 1e-5
    if nc == 2:
        assert torch.equal(mask > 0, logits[:, 1] > logits[:, 0])
    assert rel_l2(ops.nchw(da), ar.grad) < BF16_TOL
    assert rel_l2(dw, wr.grad) < F32_TOL and rel_l2(db, br.grad) < F32_TOL


@pytest.mark.parametrize("nc", [2, 4])
def test_weighted_cross_entropy(ops, nc):
    """Reference formula (utils/losses.py:49,54,57) on non-contiguous cropped views
    (scripts/train.py:118-126)."""
    n, hh, th = 3, 40, 24
    logits = rand(n, nc, th, th, scale=2.0).requires_grad_(True)
    full_t = torch.randint(0, nc, (n, 1, hh, hh), device="cuda")
    full_w = torch.rand(n, 1, hh, hh, device="cuda") * 3 + 10
    s = (hh - th) // 2
    t = full_t[:, :, s:s + th, s:s + th].squeeze(1)
    wm = full_w[:, :, s:s + th, s:s + th].squeeze(1)
    assert not t.is_contiguous()
    loss, dz, err = ops.wce_forward(logits.detach(), t, wm)
    ref = (torch.nn.CrossEntropyLoss(reduction="none")(logits, t) * wm).mean()
    ref.backward()
    torch.cuda.synchronize()
    assert int(err) == 0
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < 1e-5
    assert rel_l2(dz, logits.grad) < 1e-5


def test_ccl_matches_scipy(ops):
    from scipy import ndimage

    rng = np.random.default_rng(0)
    for h, w, p in [(64, 64, 0.45), (324, 324, 0.55), (37, 1050, 0.5), (5, 5, 1.0), (16, 16, 0.0)]:
        m = (rng.random((h, w)) < p).astype(np.uint8) * 255
        lab, _ = ndimage.label(m > 0, structure=np.ones((3, 3)))
        area = np.bincount(lab.ravel())
        small = area < 15
        small[0] = False
        ref = lab.copy()
        ref[small[lab]] = 0
        ref = ref.astype(np.uint16)
        out = ops.ccl_label(torch.from_numpy(m).cuda(), 15).cpu().numpy()
        assert np.array_equal(out, ref), (h, w, p)


# ---- wide feature maps: the "row-run" kernel (smem-side im2col) takes over when a 128-pixel tile of
# ---- one output row is >= 80 % valid (igemm_rr.cuh)
@pytest.mark.parametrize("n,h,w,ci,co", [
    (2, 7, 130, 64, 64),        # exactly one full tile per row
    (1, 9, 254, 64, 128),       # two tiles per row, second partially filled
    (2, 6, 125, 128, 256),      # 123 valid of 128, two channel chunks, BN = 256
    (1, 5, 510, 64, 64),        # four tiles per row (inc.b geometry)
])
def test_conv3x3_forward_rowrun(ops, n, h, w, ci, co):
    x = bf(rand(n, ci, h, w))
    wt = bf(rand(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=1))
    b = rand(co, seed=2)
    gamma, beta = rand(co, seed=3) + 1.5, rand(co, seed=4)
    wf, _ = ops.pack_conv3x3(wt)
    y, stats, info = ops.conv3x3_forward(ops.nhwc(x), None, wf, b, epilogue=0)
    _, _, mean, rstd = ops.bn_finalize(stats, info, gamma, beta)
    ref = F.conv2d(x, wt, b)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(y), ref) < BF16_TOL
    assert torch.allclose(mean, ref.mean((0, 2, 3)), rtol=2e-3, atol=2e-4)
    assert rel_l2(1 / rstd ** 2, ref.var((0, 2, 3), unbiased=False) + 1e-5) < 2e-3
    # eval epilogue through the same kernel
    scale, shift = rand(co, seed=5).abs() + 0.5, rand(co, seed=6)
    y2, _, _ = ops.conv3x3_forward(ops.nhwc(x), None, wf, None, epilogue=2, scale=scale, shift=shift)
    ref2 = F.relu(F.conv2d(x, wt) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(y2), ref2) < BF16_TOL


def test_conv3x3_rowrun_concat_sources(ops):
    n, hs, ws_, cs, h, w, cu, co = 2, 12, 150, 64, 8, 130, 64, 64
    skip = bf(rand(n, cs, hs, ws_))
    up = bf(rand(n, cu, h, w, seed=5))
    ch, cw = (hs - h) // 2, (ws_ - w) // 2
    wt = bf(rand(co, cs + cu, 3, 3, scale=0.04, seed=1))
    wf, _ = ops.pack_conv3x3(wt)
    y, _, _ = ops.conv3x3_forward(ops.nhwc(skip)[:, ch:ch + h, cw:cw + w, :], ops.nhwc(up), wf, None)
    ref = F.conv2d(torch.cat([skip[:, :, ch:ch + h, cw:cw + w], up], 1), wt)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(y), ref) < BF16_TOL


@pytest.mark.parametrize("n,h,w,ci,co", [(1, 6, 126, 64, 64), (2, 5, 252, 64, 128), (1, 4, 123, 128, 256)])
def test_conv3x3_dgrad_rowrun(ops, n, h, w, ci, co):
    dy = bf(rand(n, co, h, w))
    wt = bf(rand(co, ci, 3, 3, scale=0.05, seed=1))
    _, wd = ops.pack_conv3x3(wt)
    dx = ops.conv3x3_dgrad(ops.nhwc(dy), wd)
    ref = torch.nn.grad.conv2d_input((n, ci, h + 2, w + 2), wt, dy)
    torch.cuda.synchronize()
    assert rel_l2(ops.nchw(dx), ref) < BF16_TOL


@pytest.mark.parametrize("n,h,w,c,nc", [(2, 17, 13, 64, 2), (1, 40, 33, 64, 3), (2, 9, 21, 128, 2),
                                        (1, 324, 324, 64, 2)])
def test_bn_apply_relu_fused_head(ops, n, h, w, c, nc):
    """Last conv unit: BN-apply + ReLU with the 1x1 OutConv fused (reference
    models/unet_model.py:56-63,145) == separate BN-apply kernel followed by the head kernel."""
    y = bf(rand(n, c, h, w))
    scale, shift = rand(c, seed=1) * 0.3 + 1.0, rand(c, seed=2) * 0.2
    wt, b = rand(nc, c, scale=0.2, seed=3), rand(nc, seed=4)
    a, logits = ops.bn_apply_relu_head(ops.nhwc(y), scale, shift, wt, b)
    a2, _ = ops.bn_apply_relu(ops.nhwc(y), scale, shift)
    logits2, _ = ops.head_forward(a2, wt, b)
    ref_a = bf(F.relu(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)))
    ref = F.conv2d(ref_a, wt.view(nc, c, 1, 1), b)
    torch.cuda.synchronize()
    assert torch.equal(a, a2)
    assert rel_l2(logits, logits2) < 1e-6
    assert rel_l2(logits, ref) < 1e-4
    if c == 64:
        assert torch.equal(logits, logits2)      # same summation tree as the stand-alone head


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(1, 6, 128, 64, 0, 64), (2, 5, 254, 128, 0, 64),
                                            (1, 7, 128, 64, 64, 64), (2, 40, 230, 64, 0, 64),
                                            (3, 33, 330, 64, 64, 64)])
def test_conv3x3_wgrad_wide_maps(ops, n, h, w, c0, c1, co):
    """Weight gradient with Cout = 64 on wide maps (the shapes of inc.b / up4.a / up4.b), incl. the
    zero-copy concat of two sources and ragged row ends."""
    x0 = bf(rand(n, c0, h, w))
    x1 = bf(rand(n, c1, h, w, seed=9)) if c1 else None
    dy = bf(rand(n, co, h - 2, w - 2, seed=3))
    dw = ops.conv3x3_wgrad(ops.nhwc(x0), ops.nhwc(x1) if c1 else None, ops.nhwc(dy))
    x = torch.cat([x0, x1], 1) if c1 else x0
    ref = torch.nn.grad.conv2d_weight(x, (co, c0 + c1, 3, 3), dy)
    torch.cuda.synchronize()
    assert rel_l2(dw, ref) < F32_TOL and cosine(dw, ref) > 0.9999


_WGRAD_FORMS_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
from unet_segmentation_b200 import ops
g = torch.Generator(device="cuda").manual_seed(11)
outs = []
for (n, h, w, c0, c1) in [(2, 21, 77, 64, 0), (1, 9, 300, 64, 64), (3, 12, 12, 128, 0)]:
    x0 = torch.randn(n, h, w, c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(n, h, w, c1, device="cuda", generator=g).bfloat16() if c1 else None
    dy = torch.randn(n, h - 2, w - 2, 64, device="cuda", generator=g).bfloat16()
    outs.append(ops.conv3x3_wgrad(x0, x1, dy).cpu())
torch.save(outs, sys.argv[2])
"""


def test_conv3x3_wgrad_shifted_form_matches_tap_major_form(ops, tmp_path):
    """The 64-output-channel weight gradient has two forms (igemm.cuh: 128 x 192 tiles with dY displaced
    along w — the default — and the tap-major 128 x 64 tiles, UB_WGRAD_SHIFT=0). The switch is read once
    per process, so each form runs in its own interpreter on the same seeded inputs; they differ only
    in summation order."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "forms.py"
    script.write_text(_WGRAD_FORMS_SCRIPT)
    res = {}
    for form in ("1", "0"):
        out = tmp_path / f"dw_{form}.pt"
        env = dict(os.environ, UB_WGRAD_SHIFT=form)
        subprocess.run([sys.executable, str(script), root, str(out)], check=True, env=env, timeout=300)
        res[form] = torch.load(out)
    for a, b in zip(res["1"], res["0"]):
        assert a.shape == b.shape
        assert rel_l2(a, b) < 1e-5 and not torch.equal(a, torch.zeros_like(a))


@pytest.mark.parametrize("n,h,w,ci,nc", [(1, 12, 12, 64, 2), (2, 20, 30, 64, 3), (1, 6, 130, 64, 2),
                                         (2, 9, 260, 128, 2)])
def test_conv3x3_eval_fused_head(ops, n, h, w, ci, nc):
    """Eval last unit: conv + folded BN + ReLU + 1x1 OutConv + mask in ONE epilogue (im2col and
    row-run / resident-weight kernels) == the unfused conv followed by the head kernel."""
    x = bf(rand(n, ci, h, w))
    wt = bf(rand(64, ci, 3, 3, scale=0.05, seed=1))
    wf, _ = ops.pack_conv3x3(wt)
    scale, shift = rand(64, seed=2) * 0.2 + 1.0, rand(64, seed=3) * 0.3
    hw_, hb = rand(nc, 64, scale=0.2, seed=4), rand(nc, seed=5)
    logits, mask = ops.conv3x3_affine_relu_head(ops.nhwc(x), None, wf, scale, shift, hw_, hb)
    a, _, _ = ops.conv3x3_forward(ops.nhwc(x), None, wf, None, epilogue=2, scale=scale, shift=shift)
    logits2, mask2 = ops.head_forward(a, hw_, hb, want_mask=True)
    ref_a = bf(F.relu(F.conv2d(x, wt) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)))
    ref = F.conv2d(ref_a, hw_.view(nc, 64, 1, 1), hb)
    torch.cuda.synchronize()
    assert rel_l2(logits, logits2) < 1e-5          # same bf16 activations, different summation order
    assert rel_l2(logits, ref) < BF16_TOL
    margin = (logits2[:, 1] - logits2[:, 0]).abs() if nc >= 2 else None
    if nc >= 2:
        assert torch.equal(mask[margin > 1e-4], mask2[margin > 1e-4])


@pytest.mark.parametrize("n,h,w,oh,lbl_dtype,w_dtype", [(2, 36, 40, 20, torch.uint8, torch.float32),
                                                         (3, 64, 52, 40, torch.uint16, torch.float64),
                                                         (1, 512, 512, 324, torch.uint16, torch.float64)])
def test_device_input_pipeline(n, h, w, oh, lbl_dtype, w_dtype):
    """uint8 frame / instance mask / stored weight map -> (image, target, weights)
    bit-exactly as the reference builds them (utils/dataset.py:92-110, scripts/train.py:39-51,118-126)."""
    from unet_segmentation_b200 import input_pipeline
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (n, h, w), generator=g, dtype=torch.uint8)
    lbl = (torch.randint(0, 700, (n, h, w), generator=g) * (torch.rand(n, h, w, generator=g) > 0.5))
    lbl = lbl.to(torch.int32).clamp(0, 255 if lbl_dtype == torch.uint8 else 65535).to(lbl_dtype)
    wm = (10 + torch.rand(n, h, w, generator=g, dtype=torch.float64) * 3).to(w_dtype)
    image, target, weight = input_pipeline.prepare_batch(img.cuda(), lbl.cuda(), wm.cuda(), (oh, oh))
    # the reference's host pipeline + device crop
    ref_img = img.float().div(255).unsqueeze(1)
    ref_t = (lbl.to(torch.int32) > 0).long().unsqueeze(1)
    ref_w = wm.float().unsqueeze(1)
    s0h, s0w = max(0, (h - oh) // 2), max(0, (w - oh) // 2)
    ref_t = ref_t[:, :, s0h:s0h + oh, s0w:s0w + oh].squeeze(1)
    ref_w = ref_w[:, :, s0h:s0h + oh, s0w:s0w + oh].squeeze(1)
    torch.cuda.synchronize()
    assert torch.equal(image.cpu(), ref_img)
    assert target.dtype == torch.int64 and torch.equal(target.cpu(), ref_t)
    assert weight.dtype == torch.float32 and torch.equal(weight.cpu(), ref_w)
    # double-buffered staging from pinned host memory gives the same tensors
    prep = input_pipeline.DeviceBatchPreparer("cuda", (oh, oh))
    k = prep.submit(img.pin_memory(), lbl.pin_memory(), wm.pin_memory())
    image2, target2, weight2 = prep.get(k)
    torch.cuda.synchronize()
    assert torch.equal(image2, image) and torch.equal(target2, target) and torch.equal(weight2, weight)


@pytest.mark.parametrize("lbl_dtype,out_dtype", [(torch.uint16, torch.float64), (torch.uint8, torch.float32),
                                                 (torch.uint16, torch.float32)])
def test_device_weight_map_bit_exact(lbl_dtype, out_dtype):
    """calculate_weight_map (scripts/preprocess_data.py:17-77) on the device — the
    reference's stored maps (golden) and the oracle on random / empty / full / odd-sized masks."""
    from oracle import weight_map_ref
    from unet_segmentation_b200 import input_pipeline
    np_out = np.float64 if out_dtype == torch.float64 else np.float32
    if lbl_dtype == torch.uint16:
        blob = np.load(os.path.join(os.path.dirname(__file__), "golden", "weight_map_golden.npz"))
        ids = sorted(k[len("labels"):] for k in blob.files if k.startswith("labels"))
        lab = torch.from_numpy(np.stack([blob[f"labels{i}"] for i in ids]).astype(np.int32)).to(torch.uint16)
        got = input_pipeline.weight_maps_from_labels(lab.cuda(), dtype=out_dtype).cpu().numpy()
        for k, i in enumerate(ids):
            assert np.array_equal(got[k], blob[f"wmap{i}"].astype(np_out)), i
    rng = np.random.default_rng(21)
    top = 255 if lbl_dtype == torch.uint8 else 40000
    for (n, h, w), w0, sigma in [((3, 36, 40), 10, 5), ((2, 33, 17), 10, 5), ((1, 7, 9), 2.5, 0.0),
                                 ((4, 512, 512), 10, 5)]:
        lab = (rng.integers(1, top + 1, (n, h, w)) * (rng.random((n, h, w)) < rng.random((n, 1, 1)))).astype(np.int32)
        lab[0] = 0                      # an image without any instance
        if n > 1:
            lab[1] = 9                  # an image without background
        got = input_pipeline.weight_maps_from_labels(torch.from_numpy(lab).to(lbl_dtype).cuda(), w0, sigma,
                                                     out_dtype)
        assert got.dtype == out_dtype and tuple(got.shape) == (n, h, w)
        got = got.cpu().numpy()
        for k in range(n):
            ref = weight_map_ref.weight_map_closed_form(lab[k], w0, sigma).astype(np_out)
            assert np.array_equal(got[k], ref), (n, h, w, k)
    one = torch.from_numpy(lab[-1]).to(lbl_dtype).cuda()     # (H, W) input keeps its shape
    assert tuple(input_pipeline.weight_maps_from_labels(one).shape) == tuple(one.shape)


def test_device_batch_preparer_computes_missing_weight_maps():
    """Labels without stored maps: DeviceBatchPreparer derives the weights on the device (N4 -> N3)
    and yields what the reference's dataset + train-loop crop would."""
    from oracle import weight_map_ref
    from unet_segmentation_b200 import input_pipeline
    g = torch.Generator().manual_seed(5)
    n, h, w, oh = 2, 64, 64, 40
    img = torch.randint(0, 256, (n, h, w), generator=g, dtype=torch.uint8)
    lbl = (torch.randint(1, 30, (n, h, w), generator=g) * (torch.rand(n, h, w, generator=g) > 0.6)).to(torch.uint8)
    prep = input_pipeline.DeviceBatchPreparer("cuda", (oh, oh))
    image, target, weight = prep.get(prep.submit(img.pin_memory(), lbl.pin_memory(), None))
    torch.cuda.synchronize()
    s0 = (h - oh) // 2
    ref_w = torch.stack([torch.from_numpy(weight_map_ref.weight_map_closed_form(lbl[k].numpy())).float()
                         for k in range(n)])[:, s0:s0 + oh, s0:s0 + oh]
    assert torch.equal(weight.cpu(), ref_w)
    assert torch.equal(target.cpu(), (lbl > 0).long()[:, s0:s0 + oh, s0:s0 + oh])
    assert torch.equal(image.cpu(), img.float().div(255).unsqueeze(1))


@pytest.mark.parametrize("n,h,w,alpha,sigma,lbl_dtype", [(2, 96, 80, 2000, 20, torch.uint16),
                                                         (3, 33, 17, 2000, 2, torch.uint16),
                                                         (2, 5, 7, 50, 0.8, torch.uint8),
                                                         (1, 1, 9, 20, 1, torch.uint8),
                                                         (2, 512, 512, 2000, 20, torch.uint16)])
def test_device_elastic_deformation_bit_exact(n, h, w, alpha, sigma, lbl_dtype):
    """elastic_deform_image_and_mask (utils/augmentations.py:4-39) on the device
    with the reference's own per-sample seeds: image (bilinear, uint8) and instance mask (nearest)
    bit-exact against the scipy restatement of the reference."""
    from oracle import elastic_ref
    from unet_segmentation_b200 import input_pipeline
    rng = np.random.default_rng(h * w + n)
    img = rng.integers(0, 256, (n, h, w)).astype(np.uint8)
    top = 256 if lbl_dtype == torch.uint8 else 60000
    lab = (rng.integers(0, top, (n, h, w)) * (rng.random((n, h, w)) < 0.5)).astype(np.int32)
    np_lab = lab.astype(np.uint8 if lbl_dtype == torch.uint8 else np.uint16)
    seeds = [1000 + 7 * k for k in range(n)]
    noise = input_pipeline.reference_noise(seeds, (h, w))
    d_img = torch.from_numpy(img).cuda()
    d_lab = torch.from_numpy(lab).to(lbl_dtype).cuda()
    out_img, out_lab = input_pipeline.elastic_deform(d_img, d_lab, alpha, sigma, noise=noise.cuda())
    assert out_img.dtype == torch.uint8 and out_lab.dtype == lbl_dtype
    out_img, out_lab = out_img.cpu().numpy(), out_lab.cpu().numpy().astype(np.int32)
    for k in range(n):
        ri, rm = elastic_ref.elastic_deform_scipy(img[k], np_lab[k], alpha, sigma, seeds[k])
        assert np.array_equal(out_img[k], ri), (k, int((out_img[k] != ri).sum()))
        assert np.array_equal(out_lab[k], rm.astype(np.int32)), (k, int((out_lab[k] != rm).sum()))
    # uint8 label output wraps like the reference's mask.astype(np.uint8) (utils/dataset.py:93);
    # images / labels alone; device-drawn noise is reproducible from a generator
    _, lab8 = input_pipeline.elastic_deform(None, d_lab, alpha, sigma, noise=noise.cuda(), labels_as_uint8=True)
    assert lab8.dtype == torch.uint8 and np.array_equal(lab8.cpu().numpy(), (out_lab & 0xFF).astype(np.uint8))
    only_img, none = input_pipeline.elastic_deform(d_img, None, alpha, sigma, noise=noise.cuda())
    assert none is None and np.array_equal(only_img.cpu().numpy(), out_img)
    g1 = torch.Generator(device="cuda").manual_seed(3)
    g2 = torch.Generator(device="cuda").manual_seed(3)
    a = input_pipeline.elastic_deform(d_img, d_lab, alpha, sigma, generator=g1)
    b = input_pipeline.elastic_deform(d_img, d_lab, alpha, sigma, generator=g2)
    assert torch.equal(a[0], b[0]) and np.array_equal(a[1].cpu().numpy(), b[1].cpu().numpy())


def test_device_batch_preparer_with_augmentation_matches_the_manual_sequence():
    """augment=(alpha, sigma): weight maps of the UNDEFORMED labels (the reference keeps the stored
    map, utils/dataset.py:81-94), elastic deformation of frame + labels, uint8 cast of the deformed
    mask, then tensor conversion and crop."""
    from unet_segmentation_b200 import input_pipeline as ip
    g = torch.Generator().manual_seed(8)
    n, h, w, oh = 2, 96, 96, 52
    img = torch.randint(0, 256, (n, h, w), generator=g, dtype=torch.uint8)
    lbl = (torch.randint(1, 600, (n, h, w), generator=g) * (torch.rand(n, h, w, generator=g) > 0.5)).to(torch.int32).to(torch.uint16)
    prep = ip.DeviceBatchPreparer("cuda", (oh, oh), augment=(300, 4),
                                  generator=torch.Generator(device="cuda").manual_seed(77))
    image, target, weight = prep.get(prep.submit(img.pin_memory(), lbl.pin_memory(), None))
    wm = ip.weight_maps_from_labels(lbl.cuda(), dtype=torch.float32)
    d_img, d_lbl = ip.elastic_deform(img.cuda(), lbl.cuda(), 300, 4, labels_as_uint8=True,
                                     generator=torch.Generator(device="cuda").manual_seed(77))
    image2, target2, weight2 = ip.prepare_batch(d_img, d_lbl, wm, (oh, oh))
    torch.cuda.synchronize()
    assert torch.equal(image, image2) and torch.equal(target, target2) and torch.equal(weight, weight2)
    assert not torch.equal(image[:, 0], img.cuda().float().div(255))      # it did deform


@pytest.mark.parametrize("n,h,w,c,coff,ctot", [(2, 24, 24, 128, 0, 128), (1, 7, 5, 64, 64, 192),
                                               (2, 1, 3, 8, 0, 8), (1, 41, 82, 256, 128, 384)])
def test_bilinear_upsample_forward_backward(ops, n, h, w, c, coff, ctot):
    """nn.Upsample(scale_factor=2, bilinear, align_corners=True) (Up with bilinear=True,
    models/unet_model.py:40-43) and its adjoint, on a channel slice of a wider NHWC buffer (the
    up-sampled range of d(concat)), against torch in fp32 on the same bf16-rounded inputs."""
    x_full = bf(rand(n, h, w, ctot, seed=1)).to(torch.bfloat16)
    x = x_full[..., coff:coff + c]
    up = ops.upsample2x(x)
    ref_in = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = F.interpolate(ref_in, scale_factor=2, mode="bilinear", align_corners=True)
    assert up.shape == (n, 2 * h, 2 * w, c)
    assert rel_l2(up.float(), ref.detach().permute(0, 2, 3, 1)) < BF16_TOL
    g_full = bf(rand(n, 2 * h, 2 * w, ctot, seed=2)).to(torch.bfloat16)
    g = g_full[..., coff:coff + c]
    dx = ops.upsample2x_backward(g)
    ref.backward(g.float().permute(0, 3, 1, 2))
    assert dx.shape == (n, h, w, c)
    assert rel_l2(dx.float(), ref_in.grad.permute(0, 2, 3, 1)) < BF16_TOL
    assert cosine(dx.float(), ref_in.grad.permute(0, 2, 3, 1)) > 0.9999
