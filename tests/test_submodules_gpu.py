"""The building blocks' own forwards (public in the reference: DoubleConv / Down / Up / OutConv,
models/unet_model.py:20-21, 32-33, 50-54, 62-63) against the same torch layers in fp32: NCHW fp32 in
and out, eval mode (running statistics) and train mode under no_grad (batch statistics + buffer
update). They run the library's kernels through ub_op_* and must refuse to pretend to be
differentiable."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _randomise_bn(mod, seed):
    g = torch.Generator().manual_seed(seed)
    for m in mod.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def _torch_double_conv(seq, x):
    for i in (0, 3):
        conv, bn = seq[i], seq[i + 1]
        x = F.relu(F.batch_norm(F.conv2d(x, conv.weight, conv.bias), bn.running_mean, bn.running_var,
                                bn.weight, bn.bias, bn.training, bn.momentum, bn.eps))
    return x


@pytest.mark.parametrize("training", [False, True])
def test_double_conv_down_up_outconv_forwards(training):
    from unet_segmentation_b200.unet import UNet

    torch.manual_seed(3)
    model = UNet(1, 2).cuda()
    _randomise_bn(model, 5)
    model.train(training)
    twin = copy.deepcopy(model)                      # torch-side reference state (BN buffers move in train)
    g = torch.Generator(device="cuda").manual_seed(1)
    img = 0.4 + 0.2 * torch.rand(2, 1, 92, 92, device="cuda", generator=g)
    with torch.no_grad():
        # inc: image-side DoubleConv (fp32 first conv)
        x1 = model.inc(img)
        r1 = _torch_double_conv(twin.inc.double_conv, img)
        assert x1.shape == r1.shape == (2, 64, 88, 88) and x1.dtype == torch.float32
        assert rel_l2(x1, r1) < 8e-3
        # Down: max-pool + DoubleConv, fed the reference's tensor (teacher forcing)
        x2 = model.down1(r1)
        r2 = _torch_double_conv(twin.down1.maxpool_conv[1].double_conv, F.max_pool2d(r1, 2))
        assert x2.shape == r2.shape == (2, 128, 40, 40)
        assert rel_l2(x2, r2) < 8e-3
        # Up: transposed conv, [skip, up] concat, DoubleConv (reference Up.forward(x1, x2_cropped))
        deep = torch.randn(2, 128, 18, 18, device="cuda", generator=g)
        skip = r1[:, :, 26:62, 26:62]                # already centre-cropped to 36 x 36
        u = model.up4(deep, skip)
        ru = F.conv_transpose2d(deep, twin.up4.up.weight, twin.up4.up.bias, stride=2)
        ru = _torch_double_conv(twin.up4.conv.double_conv, torch.cat([skip, ru], dim=1))
        assert u.shape == ru.shape == (2, 64, 32, 32)
        assert rel_l2(u, ru) < 8e-3
        # OutConv
        z = model.outc(ru)
        rz = F.conv2d(ru, twin.outc.conv.weight, twin.outc.conv.bias)
        assert z.shape == rz.shape == (2, 2, 32, 32)
        assert rel_l2(z, rz) < 5e-3
    torch.cuda.synchronize()
    if training:    # batch statistics were used and the running buffers moved like torch's
        for name in ("inc.double_conv.1", "down1.maxpool_conv.1.double_conv.4", "up4.conv.double_conv.1"):
            a, b = dict(model.named_modules())[name], dict(twin.named_modules())[name]
            assert int(a.num_batches_tracked) == 1, name    # (F.batch_norm on the torch side does not count)
            assert rel_l2(a.running_mean, b.running_mean) < 1e-2, name
            assert rel_l2(a.running_var, b.running_var) < 1e-2, name
    # mismatched skip size fails like torch.cat would
    with torch.no_grad(), pytest.raises(RuntimeError, match="must match"):
        model.up4(deep, r1)


def test_submodule_forwards_refuse_autograd_and_cpu():
    from unet_segmentation_b200.unet import UNet

    model = UNet(1, 2).cuda().eval()
    img = torch.rand(1, 1, 60, 60, device="cuda")
    with pytest.raises(RuntimeError, match="no autograd graph|builds no"):
        model.inc(img)                               # grad mode on, parameters require grad
    with torch.no_grad():
        assert model.inc(img).shape == (1, 64, 56, 56)
        with pytest.raises(RuntimeError, match="CUDA"):
            model.inc(img.cpu())
