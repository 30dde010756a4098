"""Functional fp32 restatement of the reference U-Net and its loss (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/unet_model.py line by line but as pure functions over a
reference-keyed ``state_dict`` (so it travels to the GPU box where /root/reference is absent):

  DoubleConv      models/unet_model.py:11-17   conv3x3(p=0)+BN+ReLU twice
  Down            :28                          MaxPool2d(2) then DoubleConv
  Up (.up/.conv)  :45-46, 129-143              ConvTranspose2d(k=2,s=2), centre-crop skip, cat [skip, up]
  _center_crop    :88-102
  OutConv         :60
  UNet.forward    :105-146
  WeightedCrossEntropyLoss.forward  utils/losses.py:49,54,57
  center_crop_tensor / init_weights scripts/train.py:39-51, 54-61
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def center_crop(t: torch.Tensor, size) -> torch.Tensor:
    """models/unet_model.py:88-102 and scripts/train.py:39-51 (identical arithmetic)."""
    h, w = t.shape[-2:]
    th, tw = size
    hs, ws = max(0, (h - th) // 2), max(0, (w - tw) // 2)
    return t[..., hs:hs + th, ws:ws + tw]


# ------------------------------------------------------------------------------------------------
# T2 oracle: the reference arithmetic with bf16 rounding at the operands of
# every convolution except the first (input and weight, straight-through backward) and at the
# gradient w.r.t. every convolution output — the rounding points of a bf16-operand implementation.
# ------------------------------------------------------------------------------------------------
def _round_ste(t: torch.Tensor) -> torch.Tensor:
    """bf16 rounding in the forward pass, identity gradient."""
    return t + (t.bfloat16().float() - t).detach()


class _RoundGrad(torch.autograd.Function):
    """Identity in the forward pass; rounds the incoming gradient to bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _conv(x, w, b, emulate, first=False, transpose=False):
    if emulate and not first:
        x, w = _round_ste(x), _round_ste(w)
    y = F.conv_transpose2d(x, w, b, stride=2) if transpose else F.conv2d(x, w, b)
    return _RoundGrad.apply(y) if (emulate and not first) else y


def _keep(capture, key, t):
    """Teacher forcing: remember an intermediate and, when it is part of an autograd
    graph, ask autograd to keep its gradient (the upstream gradient of the producing layer)."""
    if capture is not None:
        if t.requires_grad:
            t.retain_grad()
        capture[key] = t
    return t


def _conv_bn_relu(x, sd, conv, bn, training, buffers_out, momentum=0.1, eps=1e-5, emulate=False,
                  capture=None):
    _keep(capture, f"{conv}.in", x)
    x = _conv(x, sd[f"{conv}.weight"], sd[f"{conv}.bias"], emulate,
              first=(conv == "inc.double_conv.0"))
    _keep(capture, f"{conv}.y", x)
    rm, rv = sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"]
    if training and buffers_out is not None:
        rm, rv = rm.clone(), rv.clone()
        buffers_out[f"{bn}.running_mean"], buffers_out[f"{bn}.running_var"] = rm, rv
        buffers_out[f"{bn}.num_batches_tracked"] = sd[f"{bn}.num_batches_tracked"] + 1
    elif training:
        rm = rv = None
    x = F.batch_norm(x, rm, rv, sd[f"{bn}.weight"], sd[f"{bn}.bias"], training, momentum, eps)
    return _keep(capture, f"{conv}.a", F.relu(x))


def _double_conv(x, sd, prefix, training, buffers_out, emulate=False, capture=None):
    x = _conv_bn_relu(x, sd, f"{prefix}.0", f"{prefix}.1", training, buffers_out, emulate=emulate,
                      capture=capture)
    return _conv_bn_relu(x, sd, f"{prefix}.3", f"{prefix}.4", training, buffers_out, emulate=emulate,
                         capture=capture)


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = False,
                 levels: int = 5, buffers_out: Optional[dict] = None,
                 capture: Optional[dict] = None, emulate_bf16: bool = False) -> torch.Tensor:
    """logits = UNet(x). ``buffers_out`` (dict) receives the updated BN buffers in training mode;
    ``capture`` (dict) receives the intermediate block outputs and, per conv unit ``<prefix>.{0,3}``,
    its input ``.in``, pre-BN output ``.y`` and post-ReLU activation ``.a`` (plus ``upJ.up.in/.out`` of the
    transposed convs) with ``retain_grad`` — after ``backward()`` their ``.grad`` are the upstream
    gradients each layer saw (teacher forcing);
    ``emulate_bf16`` selects the T2 oracle (bf16-rounded conv operands / output gradients)."""
    e = emulate_bf16
    feats = [_double_conv(x, sd, "inc.double_conv", training, buffers_out, e, capture)]
    for i in range(1, levels):
        p = F.max_pool2d(feats[-1], 2)
        feats.append(_double_conv(p, sd, f"down{i}.maxpool_conv.1.double_conv", training,
                                  buffers_out, e, capture))
    y = feats[-1]
    for j in range(1, levels):
        if f"up{j}.up.weight" in sd:
            _keep(capture, f"up{j}.up.in", y)
            up = _conv(y, sd[f"up{j}.up.weight"], sd[f"up{j}.up.bias"], e, transpose=True)
        else:   # bilinear=True: nn.Upsample has no parameters (models/unet_model.py:40-41)
            up = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
        _keep(capture, f"up{j}.up.out", up)
        skip = center_crop(feats[levels - 1 - j], up.shape[-2:])
        y = _double_conv(torch.cat([skip, up], dim=1), sd, f"up{j}.conv.double_conv", training,
                         buffers_out, e, capture)
        if capture is not None:
            capture[f"up{j}"] = y
    if capture is not None:
        for i, f in enumerate(feats):
            capture[f"x{i + 1}"] = f
    return _conv(y, sd["outc.conv.weight"], sd["outc.conv.bias"], e)


def weighted_cross_entropy(logits, targets, weight_maps):
    """utils/losses.py:49 (CE reduction='none'), :54 (* w), :57 (.mean())."""
    return (F.cross_entropy(logits, targets, reduction="none") * weight_maps).mean()


def out_size(h: int, levels: int = 5) -> int:
    """Spatial size of the logits for an input of size h (floor-mode pooling)."""
    for i in range(levels):
        h -= 4
        if i < levels - 1:
            h //= 2
    for _ in range(levels - 1):
        h = h * 2 - 4
    return h


# ------------------------------------------------------------------------------------------------
# Reference-identical construction of a random-init state_dict. The module tree is declared in the
# reference's order (models/unet_model.py:73-85) with stock torch layers, so that
# ``torch.manual_seed(s)`` consumes the RNG exactly like ``UNet(n_channels, n_classes)`` there,
# followed by ``model.apply(init_weights)`` (scripts/train.py:54-61,94).
# ------------------------------------------------------------------------------------------------
def _double_conv_modules(cin, cout):
    import torch.nn as nn

    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=0), nn.BatchNorm2d(cout), nn.ReLU(True),
                         nn.Conv2d(cout, cout, 3, padding=0), nn.BatchNorm2d(cout), nn.ReLU(True))


def make_state_dict(n_channels=1, n_classes=2, seed=0, base=64, levels=5, init=True,
                    device="cpu", bilinear=False):
    import torch.nn as nn

    torch.manual_seed(seed)
    c = [base << i for i in range(levels)]
    mods = {"inc.double_conv": _double_conv_modules(n_channels, c[0])}
    for i in range(1, levels):
        mods[f"down{i}.maxpool_conv.1.double_conv"] = _double_conv_modules(c[i - 1], c[i])
    for j in range(1, levels):
        cp = c[levels - j]
        if bilinear:    # Up(cp, cp/2, cp/2, True): DoubleConv(cp + cp/2, cp/2), models/unet_model.py:41-43
            mods[f"up{j}.conv.double_conv"] = _double_conv_modules(cp + cp // 2, cp // 2)
        else:
            mods[f"up{j}.up"] = nn.ConvTranspose2d(cp, cp // 2, kernel_size=2, stride=2)
            mods[f"up{j}.conv.double_conv"] = _double_conv_modules(cp, cp // 2)
    mods["outc.conv"] = nn.Conv2d(c[0], n_classes, kernel_size=1)
    if init:  # scripts/train.py:54-61 — nn.Conv2d only (ConvTranspose2d keeps torch's default)
        for m in mods.values():
            for sub in m.modules():
                if isinstance(sub, nn.Conv2d):
                    nn.init.kaiming_normal_(sub.weight, mode="fan_out", nonlinearity="relu")
                    if sub.bias is not None:
                        nn.init.constant_(sub.bias, 0)
                elif isinstance(sub, nn.BatchNorm2d):
                    nn.init.constant_(sub.weight, 1)
                    nn.init.constant_(sub.bias, 0)
    sd = {}
    for prefix, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().to(device)
    return sd


def synthetic_batch(n, size=512, out=None, seed=1234, levels=5, device="cpu", cropped=True):
    """DIC-C2DH-HeLa-shaped synthetic sample: low-contrast image in [0,1], target =
    union of random ellipses (fg ~0.45), two-valued class-balance weight map, target and
    weights centre-cropped + squeezed exactly like scripts/train.py:118-126 (non-contiguous)."""
    g = torch.Generator().manual_seed(seed)
    out = out_size(size, levels) if out is None else out
    img = 0.4 + 0.2 * torch.rand(n, 1, size, size, generator=g)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    mask = torch.zeros(n, 1, size, size, dtype=torch.bool)
    for b in range(n):
        for _ in range(10):
            cy, cx = (torch.rand(2, generator=g) * size).tolist()
            ry, rx = (size * (0.08 + 0.12 * torch.rand(2, generator=g))).tolist()
            mask[b, 0] |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
    f_fg = mask.float().mean().clamp(0.05, 0.95)
    wmap = torch.where(mask, 10 + 1 / f_fg, 10 + 1 / (1 - f_fg)).float()
    if not cropped:   # as the dataset hands them over: (N,1,size,size) (utils/dataset.py:96-115)
        return img.to(device), mask.long().to(device), wmap.to(device)
    t = center_crop(mask.long(), (out, out)).squeeze(1)
    w = center_crop(wmap, (out, out)).squeeze(1)
    return img.to(device), t.to(device), w.to(device)
