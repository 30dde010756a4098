"""Torch-facing wrappers of the operator-level C ABI (``ub_op_*``).

Used by the teacher-forced per-layer parity tests and by the loss /
post-processing modules. Tensors are only containers for device memory here: every function
passes raw ``data_ptr()`` values and the current CUDA stream to libunetb200 and launches the same
kernels the network executor uses. Activations are NHWC bf16 tensors of shape (N, H, W, C).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import UbView, check


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def view(t: torch.Tensor) -> UbView:
    """NHWC bf16 view of a (possibly sliced) tensor of shape (N, H, W, C) with unit channel stride."""
    if t.dtype != torch.bfloat16 or t.dim() != 4 or not t.is_cuda:
        raise ValueError("expected a CUDA bf16 tensor of shape (N, H, W, C)")
    if t.stride(3) != 1:
        raise ValueError("channel stride must be 1")
    n, h, w, c = t.shape
    return UbView(t.data_ptr(), n, h, w, c, t.stride(0), t.stride(1), t.stride(2))


def _vp(t: torch.Tensor | None):
    return None if t is None else C.byref(view(t))


def nhwc(x_nchw: torch.Tensor) -> torch.Tensor:
    """fp32/any NCHW -> contiguous NHWC bf16 (test helper)."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.permute(0, 3, 1, 2).float()


# ------------------------------------------------------------------------------------------------
def pack_conv3x3(w: torch.Tensor, with_dgrad: bool = True):
    lib = _lib.load()
    co, ci = w.shape[:2]
    w = w.contiguous().float()
    wf = torch.empty(9, co, ci, dtype=torch.bfloat16, device=w.device)       # [tap][Co][Ci]
    wd = torch.empty(9, ci, co, dtype=torch.bfloat16, device=w.device) if with_dgrad else None
    check(lib.ub_op_pack_conv3x3(_p(w), co, ci, _p(wf), _p(wd), _stream()), "pack_conv3x3")
    return wf, wd


def pack_convT(w: torch.Tensor, bias: torch.Tensor | None = None):
    lib = _lib.load()
    ci, co = w.shape[:2]
    w = w.contiguous().float()
    wf = torch.empty(4 * co, ci, dtype=torch.bfloat16, device=w.device)
    wb = torch.empty(4, ci, co, dtype=torch.bfloat16, device=w.device)       # [q][Ci][Co]
    b4 = torch.empty(4 * co, dtype=torch.float32, device=w.device) if bias is not None else None
    check(lib.ub_op_pack_convT(_p(w), ci, co, _p(wf), _p(wb), _p(bias), _p(b4), _stream()),
          "pack_convT")
    return wf, wb, b4


def conv3x3_forward(src0, src1, wf, bias, epilogue=1, scale=None, shift=None):
    """Returns (y, stats, info). epilogue 0 = +bias & BN statistics, 1 = +bias, 2 = affine+ReLU."""
    lib = _lib.load()
    n, h, w, _ = src0.shape
    co = wf.shape[1]
    y = torch.empty(n, h - 2, w - 2, co, dtype=torch.bfloat16, device=src0.device)
    stats = None
    info = (C.c_int * 4)()
    if epilogue == 0:
        stats = torch.zeros(int(lib.ub_op_conv_stats_floats(co)), dtype=torch.float32,
                            device=src0.device)
    check(lib.ub_op_conv3x3_forward(_vp(src0), _vp(src1), _p(wf), _p(bias), co, epilogue,
                                    _p(scale), _p(shift), _p(y), _p(stats), info, _stream()),
          "conv3x3_forward")
    return y, stats, info


def bn_finalize(stats, info, gamma, beta, running_mean=None, running_var=None, nbt=None,
                momentum=0.1, eps=1e-5):
    lib = _lib.load()
    c = gamma.numel()
    out = [torch.empty(c, dtype=torch.float32, device=gamma.device) for _ in range(4)]
    check(lib.ub_op_bn_finalize(_p(stats), info, c, _p(gamma), _p(beta), _p(running_mean),
                                _p(running_var), _p(nbt), momentum, eps, *[_p(o) for o in out],
                                _stream()), "bn_finalize")
    return out  # scale, shift, mean, rstd


def bn_apply_relu(y, scale, shift, pool=False):
    lib = _lib.load()
    n, h, w, c = y.shape
    a = torch.empty_like(y)
    p = torch.empty(n, h // 2, w // 2, c, dtype=y.dtype, device=y.device) if pool else None
    am = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=y.device) if pool else None
    check(lib.ub_op_bn_apply_relu(_p(y), _p(a), _p(p), _p(am), n, h, w, c, _p(scale), _p(shift),
                                  _stream()), "bn_apply_relu")
    return (a, p, am) if pool else (a, p)


def conv3x3_affine_relu_head(src0, src1, wf, scale, shift, head_w, head_b, want_mask=True):
    """Eval last unit: conv + folded BN + ReLU + 1x1 head (+ mask) in one kernel -> (logits, mask)."""
    lib = _lib.load()
    n, h, w, _ = src0.shape
    nc = head_w.shape[0]
    logits = torch.empty(n, nc, h - 2, w - 2, dtype=torch.float32, device=src0.device)
    mask = torch.empty(n, h - 2, w - 2, dtype=torch.uint8, device=src0.device) if want_mask else None
    check(lib.ub_op_conv3x3_affine_relu_head(_vp(src0), _vp(src1), _p(wf), _p(scale), _p(shift),
                                             _p(head_w.contiguous()), _p(head_b), nc, _p(logits),
                                             _p(mask), _stream()), "conv3x3_affine_relu_head")
    return logits, mask


def bn_apply_relu_head(y, scale, shift, head_w, head_b):
    """BN-apply + ReLU with the 1x1 head fused: returns (a bf16 NHWC, logits fp32 NCHW)."""
    lib = _lib.load()
    n, h, w, c = y.shape
    nc = head_w.shape[0]
    a = torch.empty_like(y)
    logits = torch.empty(n, nc, h, w, dtype=torch.float32, device=y.device)
    check(lib.ub_op_bn_apply_relu_head(_p(y), _p(a), n, h, w, c, _p(scale), _p(shift), nc,
                                       _p(head_w.contiguous()), _p(head_b), _p(logits), _stream()),
          "bn_apply_relu_head")
    return a, logits


def bn_relu_backward(y, scale, shift, mean, rstd, g=None, gp=None, gs=None, crop=(0, 0),
                     argmax=None):
    lib = _lib.load()
    n, h, w, c = y.shape
    ws = torch.empty(int(lib.ub_op_bn_bwd_workspace_floats(c)), dtype=torch.float32, device=y.device)
    dgamma = torch.empty(c, dtype=torch.float32, device=y.device)
    dbeta = torch.empty_like(dgamma)
    dy = torch.empty_like(y)
    check(lib.ub_op_bn_relu_backward(_p(y), n, h, w, c, _p(scale), _p(shift), _p(mean), _p(rstd),
                                     _vp(g), _vp(gp), _vp(gs), crop[0], crop[1], _p(argmax), _p(ws),
                                     _p(dgamma), _p(dbeta), _p(dy), _stream()), "bn_relu_backward")
    return dy, dgamma, dbeta


def conv3x3_dgrad(dy, wd):
    lib = _lib.load()
    n, h, w, _ = dy.shape
    ci = wd.shape[1]
    dx = torch.empty(n, h + 2, w + 2, ci, dtype=torch.bfloat16, device=dy.device)
    check(lib.ub_op_conv3x3_dgrad(_vp(dy), _p(wd), ci, _p(dx), _stream()), "conv3x3_dgrad")
    return dx


def conv3x3_dgrad_bnred(dy, wd, y, scale, shift, mean):
    """Data gradient with the BN + ReLU backward REDUCE pass of the receiving layer fused into its
    epilogue. Returns (dx, partial, info) — feed the latter two to ``bn_relu_backward_fused``."""
    lib = _lib.load()
    n, h, w, _ = dy.shape
    ci = wd.shape[1]
    dx = torch.empty(n, h + 2, w + 2, ci, dtype=torch.bfloat16, device=dy.device)
    partial = torch.zeros(int(lib.ub_op_conv_stats_floats(ci)), dtype=torch.float32, device=dy.device)
    info = (C.c_int * 4)()
    check(lib.ub_op_conv3x3_dgrad_bnred(_vp(dy), _p(wd), ci, _p(dx), _p(y), _p(scale), _p(shift),
                                        _p(mean), _p(partial), info, _stream()), "conv3x3_dgrad_bnred")
    return dx, partial, info


def bn_relu_backward_fused(y, scale, shift, mean, rstd, g, partial, info):
    lib = _lib.load()
    n, h, w, c = y.shape
    dgamma = torch.empty(c, dtype=torch.float32, device=y.device)
    dbeta = torch.empty_like(dgamma)
    dy = torch.empty_like(y)
    check(lib.ub_op_bn_relu_backward_fused(_p(y), n, h, w, c, _p(scale), _p(shift), _p(mean), _p(rstd),
                                           _vp(g), _p(partial), info, _p(dgamma), _p(dbeta), _p(dy),
                                           _stream()), "bn_relu_backward_fused")
    return dy, dgamma, dbeta


def conv3x3_wgrad(src0, src1, dy):
    lib = _lib.load()
    n, h, w, c0 = src0.shape
    ci = c0 + (src1.shape[3] if src1 is not None else 0)
    co = dy.shape[3]
    nws = int(lib.ub_op_wgrad_workspace_floats(9 * ci, co, n * (h - 2) * (w - 2)))
    ws = torch.empty(nws, dtype=torch.float32, device=dy.device)
    dw = torch.empty(co, ci, 3, 3, dtype=torch.float32, device=dy.device)
    check(lib.ub_op_conv3x3_wgrad(_vp(src0), _vp(src1), _p(dy), co, _p(ws), nws, _p(dw), _stream()),
          "conv3x3_wgrad")
    return dw


def convT_forward(x, wf, bias4, dst):
    """Writes the up-sampled tensor into ``dst`` (an (N,2H,2W,Co) view, may be a channel slice)."""
    lib = _lib.load()
    co = wf.shape[0] // 4
    check(lib.ub_op_convT_forward(_vp(x), _p(wf), _p(bias4), co, _vp(dst), _stream()),
          "convT_forward")
    return dst


def convT_dgrad(dup, wb):
    lib = _lib.load()
    n, h2, w2, _ = dup.shape
    ci = wb.shape[1]
    dx = torch.empty(n, h2 // 2, w2 // 2, ci, dtype=torch.bfloat16, device=dup.device)
    check(lib.ub_op_convT_dgrad(_vp(dup), _p(wb), ci, _p(dx), _stream()), "convT_dgrad")
    return dx


def convT_wgrad(dup, x):
    lib = _lib.load()
    n, h, w, ci = x.shape
    co = dup.shape[3]
    nws = int(lib.ub_op_wgrad_workspace_floats(4 * co, ci, n * h * w))
    ws = torch.empty(nws, dtype=torch.float32, device=x.device)
    dw = torch.empty(ci, co, 2, 2, dtype=torch.float32, device=x.device)
    check(lib.ub_op_convT_wgrad(_vp(dup), _p(x), ci, _p(ws), nws, _p(dw), _stream()), "convT_wgrad")
    return dw


def first_conv_forward(x, w, bias, gamma, beta, running_mean=None, running_var=None, nbt=None,
                       momentum=0.1, eps=1e-5):
    lib = _lib.load()
    n, ci, h, wd_ = x.shape
    co = w.shape[0]
    ws = torch.empty(int(lib.ub_op_first_conv_workspace_floats(co)), dtype=torch.float32,
                     device=x.device)
    st = [torch.empty(co, dtype=torch.float32, device=x.device) for _ in range(4)]
    a = torch.empty(n, h - 2, wd_ - 2, co, dtype=torch.bfloat16, device=x.device)
    check(lib.ub_op_first_conv_forward(_p(x), n, ci, h, wd_, _p(w), _p(bias), co, _p(gamma),
                                       _p(beta), _p(running_mean), _p(running_var), _p(nbt),
                                       momentum, eps, _p(ws), *[_p(s) for s in st], _p(a),
                                       _stream()), "first_conv_forward")
    return a, st


def first_conv_affine_relu(x, w, scale, shift):
    """Eval form of the first conv: relu(conv(x, w) * scale + shift) -> bf16 NHWC."""
    lib = _lib.load()
    n, ci, h, wd_ = x.shape
    co = w.shape[0]
    a = torch.empty(n, h - 2, wd_ - 2, co, dtype=torch.bfloat16, device=x.device)
    check(lib.ub_op_first_conv_affine_relu(_p(x), n, ci, h, wd_, _p(w), co, _p(scale), _p(shift),
                                           _p(a), _stream()), "first_conv_affine_relu")
    return a


def first_conv_backward(x, w, bias, st, g, a):
    lib = _lib.load()
    n, ci, h, wd_ = x.shape
    co = w.shape[0]
    ws = torch.empty(int(lib.ub_op_first_conv_workspace_floats(co)), dtype=torch.float32,
                     device=x.device)
    dgamma = torch.empty(co, dtype=torch.float32, device=x.device)
    dbeta = torch.empty_like(dgamma)
    dw = torch.empty(co, ci, 3, 3, dtype=torch.float32, device=x.device)
    check(lib.ub_op_first_conv_backward(_p(x), n, ci, h, wd_, _p(w), _p(bias), co,
                                        *[_p(s) for s in st], _vp(g), _p(a), _p(ws), _p(dgamma),
                                        _p(dbeta), _p(dw), _stream()), "first_conv_backward")
    return dw, dgamma, dbeta


def head_forward(a, w, b, want_mask=False):
    lib = _lib.load()
    n, h, wd_, k = a.shape
    nc = w.shape[0]
    logits = torch.empty(n, nc, h, wd_, dtype=torch.float32, device=a.device)
    mask = torch.empty(n, h, wd_, dtype=torch.uint8, device=a.device) if want_mask else None
    check(lib.ub_op_head_forward(_p(a), n, h, wd_, k, nc, _p(w), _p(b), _p(logits), _p(mask),
                                 _stream()), "head_forward")
    return logits, mask


def head_backward(dlogits, a, w):
    lib = _lib.load()
    n, h, wd_, k = a.shape
    nc = w.shape[0]
    ws = torch.empty(int(lib.ub_op_head_bwd_workspace_floats(k, nc)), dtype=torch.float32,
                     device=a.device)
    da = torch.empty_like(a)
    dw = torch.empty(nc, k, dtype=torch.float32, device=a.device)
    db = torch.empty(nc, dtype=torch.float32, device=a.device)
    check(lib.ub_op_head_backward(_p(dlogits), _p(a), n, h, wd_, k, nc, _p(w), _p(da), _p(ws),
                                  _p(dw), _p(db), _stream()), "head_backward")
    return da, dw, db


def upsample2x(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) on an NHWC bf16 view."""
    lib = _lib.load()
    n, h, w, c = x.shape
    out = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=x.device)
    check(lib.ub_op_upsample2x_forward(_vp(x), _p(out), _stream()), "upsample2x_forward")
    return out


def upsample2x_backward(g):
    """Adjoint of ``upsample2x``: g (N, 2H, 2W, C) NHWC bf16 view -> (N, H, W, C)."""
    lib = _lib.load()
    n, h2, w2, c = g.shape
    dx = torch.empty(n, h2 // 2, w2 // 2, c, dtype=torch.bfloat16, device=g.device)
    check(lib.ub_op_upsample2x_backward(_vp(g), _p(dx), _stream()), "upsample2x_backward")
    return dx


def maxpool2(a):
    lib = _lib.load()
    n, h, w, c = a.shape
    p = torch.empty(n, h // 2, w // 2, c, dtype=a.dtype, device=a.device)
    check(lib.ub_op_maxpool2(_p(a), _p(p), n, h, w, c, _stream()), "maxpool2")
    return p


def wce_forward(logits, targets, weight_maps, want_grad=True):
    """Returns (loss 0-dim fp32, dlogits or None). Inputs may be arbitrary strided views."""
    lib = _lib.load()
    if not (logits.is_cuda and targets.is_cuda and weight_maps.is_cuda):
        raise RuntimeError("WeightedCrossEntropyLoss (B200) needs CUDA tensors; there is no CPU path")
    if logits.dtype != torch.float32:
        logits = logits.float()
    if targets.dtype != torch.int64:
        targets = targets.long()
    if weight_maps.dtype != torch.float32:
        weight_maps = weight_maps.float()
    n, c, h, w = logits.shape
    if tuple(targets.shape) != (n, h, w) or tuple(weight_maps.shape) != (n, h, w):
        raise ValueError(f"shape mismatch: logits {tuple(logits.shape)}, targets "
                         f"{tuple(targets.shape)}, weight_maps {tuple(weight_maps.shape)}")
    dev = logits.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dz = torch.empty(n, c, h, w, dtype=torch.float32, device=dev) if want_grad else None
    ws = torch.empty(int(lib.ub_wce_workspace_floats()), dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ls = (C.c_int64 * 4)(*logits.stride())
    ts = (C.c_int64 * 3)(*targets.stride())
    wss = (C.c_int64 * 3)(*weight_maps.stride())
    check(lib.ub_wce_forward(_p(logits), ls, _p(targets), ts, _p(weight_maps), wss, n, c, h, w,
                             _p(loss), _p(dz), _p(ws), _p(err), _stream()), "wce_forward")
    return loss, dz, err


def scale_by_device_scalar(t, scalar):
    lib = _lib.load()
    out = torch.empty_like(t)
    check(lib.ub_scale_by_device_scalar(_p(t), _p(scalar), _p(out), t.numel(), _stream()), "scale")
    return out


def ccl_label(mask: torch.Tensor, min_size: int = 15) -> torch.Tensor:
    """uint8 (H, W) CUDA mask (>0 = foreground) -> uint16 instance labels, reference-exact."""
    lib = _lib.load()
    if mask.dim() != 2 or not mask.is_cuda:
        raise ValueError("expected a 2-D CUDA mask")
    m = (mask > 0).to(torch.uint8).contiguous()
    h, w = m.shape
    ws = torch.empty(int(lib.ub_ccl_workspace_bytes(h, w)), dtype=torch.uint8, device=m.device)
    out = torch.empty(h, w, dtype=torch.int16, device=m.device)  # reinterpret as uint16 below
    check(lib.ub_ccl_label(_p(m), h, w, min_size, _p(out), _p(ws), _stream()), "ccl_label")
    return out.view(torch.uint16)
