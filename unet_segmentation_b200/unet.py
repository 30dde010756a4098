"""Drop-in ``UNet`` for the reference's models/unet_model.py.

Same constructor, attributes, sub-module tree and ``state_dict`` keys as the reference
(models/unet_model.py:5-85: ``inc.double_conv.{0,1,3,4}``, ``downK.maxpool_conv.1.double_conv.*``,
``upK.up``, ``upK.conv.double_conv.*``, ``outc.conv``; 82 parameters, 54 buffers), so
``model.apply(init_weights)``, ``.to()``, ``load_state_dict`` and ``optim.SGD(model.parameters())``
(scripts/train.py:93-97, scripts/predict.py:120-123) work unchanged. The sub-modules are genuine
``nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d`` *parameter holders*; ``forward`` is a single
``torch.autograd.Function`` over the C-ABI network executor of libunetb200 (sm_100a). There is no
CPU path and no cuDNN/cuBLAS dispatch: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check

# ------------------------------------------------------------------------------------------------
# Stand-alone forwards of the building blocks. The reference's DoubleConv / Down / Up / OutConv are
# public modules with their own forward (models/unet_model.py:20-21, 32-33, 50-54, 62-63) although
# UNet.forward is their only caller. Here they run the same sm_100a kernels as the network executor
# through the operator-level C ABI (ub_op_*): NCHW fp32 in / out like the reference, NHWC bf16 inside,
# train mode = batch statistics + running-stat update, eval mode = folded BatchNorm. They build NO
# autograd graph (training goes through UNet.forward, which owns the backward pass).
# ------------------------------------------------------------------------------------------------
def _require_inference(mod: nn.Module, *inputs) -> None:
    if torch.is_grad_enabled() and (any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs)
                                    or any(p.requires_grad for p in mod.parameters())):
        raise RuntimeError(f"{type(mod).__name__}.forward (B200) runs the inference kernels and builds no "
                           "autograd graph; call it under torch.no_grad(), or train through "
                           "UNet.forward, which implements the backward pass")
    for t in inputs:
        if not isinstance(t, torch.Tensor) or t.dim() != 4 or not t.is_cuda:
            raise RuntimeError(f"{type(mod).__name__}.forward (B200) needs (N, C, H, W) CUDA tensors; "
                               "there is no CPU fallback")


def _to_nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.float().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _to_nchw(a: torch.Tensor) -> torch.Tensor:
    return a.permute(0, 3, 1, 2).float().contiguous()


def _folded_affine(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
    bias = conv.bias.float() if conv.bias is not None else torch.zeros_like(scale)
    return scale.contiguous(), (bn.bias.float() + (bias - bn.running_mean.float()) * scale).contiguous()


def _conv_bn_relu(conv: nn.Conv2d, bn: nn.BatchNorm2d, src0, src1=None, x_fp32=None) -> torch.Tensor:
    """One [conv3x3 -> BN -> ReLU] unit on NHWC bf16 sources (src1 = second channel range of a
    zero-copy concat), or on the fp32 NCHW image for a first conv whose C_in is not a multiple of 64."""
    from . import ops

    training = bn.training
    if training and (bn.momentum is None or not bn.track_running_stats):
        raise RuntimeError("BatchNorm2d(momentum=None / track_running_stats=False) is not supported")
    w = conv.weight.detach().float().contiguous()
    gamma, beta = bn.weight.detach().float(), bn.bias.detach().float()
    if x_fp32 is not None:
        if training:
            a, _ = ops.first_conv_forward(x_fp32, w, conv.bias.detach().float(), gamma, beta,
                                          bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                          float(bn.momentum), float(bn.eps))
            return a
        scale, shift = _folded_affine(conv, bn)
        return ops.first_conv_affine_relu(x_fp32, w, scale, shift)
    wf, _ = ops.pack_conv3x3(w, with_dgrad=False)
    if training:
        y, stats, info = ops.conv3x3_forward(src0, src1, wf, conv.bias.detach().float(), epilogue=0)
        scale, shift, _, _ = ops.bn_finalize(stats, info, gamma, beta, bn.running_mean,
                                             bn.running_var, bn.num_batches_tracked,
                                             float(bn.momentum), float(bn.eps))
        return ops.bn_apply_relu(y, scale, shift)[0]
    scale, shift = _folded_affine(conv, bn)
    return ops.conv3x3_forward(src0, src1, wf, None, epilogue=2, scale=scale, shift=shift)[0]


def _double_conv(seq: nn.Sequential, src0, src1=None, x_fp32=None) -> torch.Tensor:
    a = _conv_bn_relu(seq[0], seq[1], src0, src1, x_fp32)
    return _conv_bn_relu(seq[3], seq[4], a)


class DoubleConv(nn.Module):
    """[Conv3x3(valid) -> BatchNorm2d -> ReLU] x 2 (reference :5-21)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        layers = []
        for cin in (in_channels, out_channels):
            layers += [nn.Conv2d(cin, out_channels, kernel_size=3, padding=0),
                       nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True)]
        self.double_conv = nn.Sequential(*layers)

    def forward(self, x):
        _require_inference(self, x)
        with torch.cuda.device(x.device):
            if self.double_conv[0].in_channels % 64:      # image-side block: fp32 first conv
                return _to_nchw(_double_conv(self.double_conv, None, None, x.float().contiguous()))
            return _to_nchw(_double_conv(self.double_conv, _to_nhwc(x)))


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv (reference :23-33)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        from . import ops

        _require_inference(self, x)
        with torch.cuda.device(x.device):
            pooled = ops.maxpool2(_to_nhwc(x))              # floor mode, like nn.MaxPool2d(2)
            return _to_nchw(_double_conv(self.maxpool_conv[1].double_conv, pooled))


class Up(nn.Module):
    """Up-sampling then DoubleConv over [cropped skip, up] (reference :35-54): ConvTranspose2d(k=2,
    s=2) halving the channels, or — ``bilinear=True`` — nn.Upsample(scale_factor=2, bilinear,
    align_corners=True) keeping them (so the DoubleConv takes prev + skip channels, :41-43)."""

    def __init__(self, in_channels_from_prev_decoder: int, skip_channels: int, out_channels: int,
                 bilinear: bool = True):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConv(in_channels_from_prev_decoder + skip_channels, out_channels)
        else:
            half = in_channels_from_prev_decoder // 2
            self.up = nn.ConvTranspose2d(in_channels_from_prev_decoder, half, kernel_size=2, stride=2)
            self.conv = DoubleConv(half + skip_channels, out_channels)

    def forward(self, x1, x2_cropped):
        """x1: previous decoder output, x2_cropped: the skip tensor already centre-cropped to the
        up-sampled size (reference :50-54: ``x1 = self.up(x1); cat([x2_cropped, x1]); self.conv``).
        The concat is never materialised: the conv reads the two tensors as two channel sources."""
        from . import ops

        _require_inference(self, x1, x2_cropped)
        with torch.cuda.device(x1.device):
            x1n = _to_nhwc(x1)
            if isinstance(self.up, nn.ConvTranspose2d):
                wf, _, b4 = ops.pack_convT(self.up.weight.detach().float(), self.up.bias.detach().float())
                n, h, w, _ = x1n.shape
                up = torch.empty(n, 2 * h, 2 * w, self.up.out_channels, dtype=torch.bfloat16,
                                 device=x1.device)
                ops.convT_forward(x1n, wf, b4, up)
            else:
                up = ops.upsample2x(x1n)
            if tuple(x2_cropped.shape[2:]) != tuple(up.shape[1:3]):
                raise RuntimeError(f"Sizes of tensors must match: skip {tuple(x2_cropped.shape[2:])} vs "
                                   f"up-sampled {tuple(up.shape[1:3])} (crop the skip first)")
            return _to_nchw(_double_conv(self.conv.double_conv, _to_nhwc(x2_cropped), up))


class OutConv(nn.Module):
    """1x1 convolution to n_classes logits (reference :56-63)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        from . import ops

        _require_inference(self, x)
        with torch.cuda.device(x.device):
            w = self.conv.weight.detach().float().flatten(1).contiguous()
            b = self.conv.bias.detach().float() if self.conv.bias is not None else None
            return ops.head_forward(_to_nhwc(x), w, b)[0]


# ------------------------------------------------------------------------------------------------
class _Plan:
    """Python handle of a ``ub_plan`` (one per input shape / mode / device)."""

    def __init__(self, n, cin, h, w, base, levels, n_classes, training, device, bilinear=False):
        self.lib = _lib.load()
        self.device = device
        self.training = training
        handle = C.c_void_p()
        with torch.cuda.device(device):
            args = (C.byref(handle), n, cin, h, w, base, levels, n_classes, 1 if training else 0,
                    1 if bilinear else 0)
            status = self.lib.ub_plan_create_ex(*args)
            if status == -5:      # UB_ERR_NOMEM: the arena is cudaMalloc'ed beside torch's caching
                import gc         # allocator — hand its cached, unused blocks back and retry once

                gc.collect()
                torch.cuda.empty_cache()
                status = self.lib.ub_plan_create_ex(*args)
            check(status, "ub_plan_create")
        self.handle = handle
        oh, ow = C.c_int(), C.c_int()
        check(self.lib.ub_plan_out_hw(handle, C.byref(oh), C.byref(ow)))
        self.out_hw = (oh.value, ow.value)
        self.n, self.n_classes, self.cin, self.in_hw = n, n_classes, cin, (h, w)
        self._static = None
        self.num_params = self.lib.ub_plan_num_params(handle)
        self.num_stages = self.lib.ub_plan_num_stages(handle)
        self.numels = [int(self.lib.ub_plan_param_numel(handle, i)) for i in range(self.num_params)]
        self.offsets = [0]
        for k in self.numels:
            self.offsets.append(self.offsets[-1] + k)
        self.stage_ranges = []
        for s in range(self.num_stages):
            f, c = C.c_int(), C.c_int()
            check(self.lib.ub_plan_stage_params(handle, s, C.byref(f), C.byref(c)))
            self.stage_ranges.append((f.value, c.value))
        self.bound_ptrs = None
        self.bound_buf_ptrs = None
        self.bound_bn_cfg = None
        self.packed_versions = None
        self.generation = 0

    def __del__(self):
        try:
            if self.handle:
                self.lib.ub_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def device_bytes(self) -> int:
        return int(self.lib.ub_plan_device_bytes(self.handle))

    def bind_pointers_only(self, params, bn_modules):
        """(Re)bind parameter / buffer pointers without touching the packed operand caches."""
        ptrs = tuple(p.data_ptr() for p in params)
        if ptrs != self.bound_ptrs:
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            check(self.lib.ub_plan_bind_params(self.handle, arr, len(ptrs)), "ub_plan_bind_params")
            self.bound_ptrs = ptrs
            self.packed_versions = None

    def mark_packed(self, params, epoch):
        """The fused optimizer refreshed this plan's operand caches itself."""
        self.packed_versions = (tuple(p._version for p in params), epoch)

    def bind(self, params, bn_modules, epoch=0):
        ptrs = tuple(p.data_ptr() for p in params)
        if ptrs != self.bound_ptrs:
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            check(self.lib.ub_plan_bind_params(self.handle, arr, len(ptrs)), "ub_plan_bind_params")
            self.bound_ptrs = ptrs
            self.packed_versions = None
        bptrs = tuple((m.running_mean.data_ptr(), m.running_var.data_ptr(),
                       m.num_batches_tracked.data_ptr()) for m in bn_modules)
        if bptrs != self.bound_buf_ptrs:
            k = len(bptrs)
            rm = (C.c_void_p * k)(*[b[0] for b in bptrs])
            rv = (C.c_void_p * k)(*[b[1] for b in bptrs])
            nb = (C.c_void_p * k)(*[b[2] for b in bptrs])
            check(self.lib.ub_plan_bind_bn_buffers(self.handle, rm, rv, nb, k),
                  "ub_plan_bind_bn_buffers")
            self.bound_buf_ptrs = bptrs
        cfg = tuple((float(m.momentum), float(m.eps)) for m in bn_modules)
        if cfg != self.bound_bn_cfg:
            k = len(cfg)
            mom = (C.c_float * k)(*[c[0] for c in cfg])
            eps = (C.c_float * k)(*[c[1] for c in cfg])
            check(self.lib.ub_plan_set_bn_config(self.handle, mom, eps, k), "ub_plan_set_bn_config")
            self.bound_bn_cfg = cfg
        versions = (tuple(p._version for p in params), epoch)
        if versions != self.packed_versions:
            check(self.lib.ub_plan_pack_weights(self.handle, _stream()), "ub_plan_pack_weights")
            self.packed_versions = versions

    def static_io(self, device):
        """Eval plans own their input / logits / mask buffers: stable pointers let the library replay
        the forward pass from a CUDA graph (ub_plan_forward) and let overlap-tile inference gather
        tiles straight into the network input and stitch straight out of the mask buffer."""
        if self._static is None:
            oh, ow = self.out_hw
            self._static = (
                torch.empty(self.n, self.cin, self.in_hw[0], self.in_hw[1], dtype=torch.float32, device=device),
                torch.empty(self.n, self.n_classes, oh, ow, dtype=torch.float32, device=device),
                torch.empty(self.n, oh, ow, dtype=torch.uint8, device=device))
        return self._static

    def forward(self, x, want_mask=False, static_out=False):
        """Training plans: fresh output tensors. Eval plans: run on the plan's static buffers (x is
        copied in unless it IS the static input) and return clones — or the static buffers
        themselves with ``static_out=True`` (valid until the next forward of this plan)."""
        oh, ow = self.out_hw
        if not self.training:
            xs, logits, mask = self.static_io(x.device)
            if x.data_ptr() != xs.data_ptr():
                xs.copy_(x)
            check(self.lib.ub_plan_forward(self.handle, C.c_void_p(xs.data_ptr()),
                                           C.c_void_p(logits.data_ptr()), C.c_void_p(mask.data_ptr()),
                                           _stream()), "ub_plan_forward")
            self.generation += 1
            if static_out:
                return logits, (mask if want_mask else None)
            return logits.clone(), (mask.clone() if want_mask else None)
        logits = torch.empty(self.n, self.n_classes, oh, ow, dtype=torch.float32, device=x.device)
        check(self.lib.ub_plan_forward(self.handle, C.c_void_p(x.data_ptr()),
                                       C.c_void_p(logits.data_ptr()), C.c_void_p(0), _stream()),
              "ub_plan_forward")
        self.generation += 1
        return logits, None

    def graph_replays(self) -> int:
        return int(self.lib.ub_plan_graph_replays(self.handle))

    def join_side(self, stream: "torch.cuda.Stream") -> None:
        """Make ``stream`` wait for the weight gradients issued so far on the library's side stream."""
        check(self.lib.ub_plan_join_side(self.handle, C.c_void_p(stream.cuda_stream)), "ub_plan_join_side")

    def backward(self, dlogits, stage_hook: Optional[Callable] = None):
        flat = torch.empty(self.offsets[-1], dtype=torch.float32, device=dlogits.device)
        views = [flat[self.offsets[i]:self.offsets[i + 1]] for i in range(self.num_params)]
        arr = (C.c_void_p * self.num_params)(*[v.data_ptr() for v in views])
        dl = C.c_void_p(dlogits.data_ptr())
        # weight gradients overlap the BN-backward / data-gradient chain on an internal stream; with a
        # per-stage hook (data-parallel all-reduce) every stage is joined, otherwise only the last
        # (a hook that orders its own stream after the side stream — join_side — keeps full overlap)
        joins_itself = bool(getattr(getattr(stage_hook, "__self__", None), "joins_side_stream", False))
        check(self.lib.ub_plan_set_overlap(self.handle,
                                           1 if (stage_hook is not None and not joins_itself) else 2))
        for s in range(self.num_stages):
            check(self.lib.ub_plan_backward_stage(self.handle, s, dl, arr, _stream()),
                  f"ub_plan_backward_stage({s})")
            if stage_hook is not None:
                f, c = self.stage_ranges[s]
                stage_hook(s, flat[self.offsets[f]:self.offsets[f + c]])
        return flat, views


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        plan = module._plan_for(x, training=True)
        logits, _ = plan.forward(x)
        ctx.plan = plan
        ctx.generation = plan.generation
        ctx.module = module
        ctx.shapes = [p.shape for p in params]
        ctx.save_for_backward(x)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("UNet (B200): the saved activations of this forward pass were "
                               "overwritten by a later forward of the same shape; call backward "
                               "before the next training-mode forward")
        (x,) = ctx.saved_tensors  # keeps the input alive: the first conv recomputes from it
        dlogits = dlogits.contiguous().float()
        module = ctx.module
        _, views = plan.backward(dlogits, stage_hook=module._stage_hook)
        if module._backward_done_hook is not None:
            module._backward_done_hook()
        grads = [v.view(s) for v, s in zip(views, ctx.shapes)]
        return (None, None, *grads)


class UNet(nn.Module):
    """``UNet(n_channels, n_classes, bilinear=False)`` — reference signature
    (models/unet_model.py:65-66). ``base_channels`` / ``levels`` are trailing keyword knobs with the
    reference's hard-coded values (64-128-256-512-1024, :73-82) as defaults."""

    def __init__(self, n_channels, n_classes, bilinear=False, *, base_channels: int = 64,
                 levels: int = 5):
        super().__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        self.base_channels = base_channels
        self.levels = levels
        if levels == 5:
            c = [base_channels << i for i in range(5)]
            self.inc = DoubleConv(n_channels, c[0])
            self.down1 = Down(c[0], c[1])
            self.down2 = Down(c[1], c[2])
            self.down3 = Down(c[2], c[3])
            self.down4 = Down(c[3], c[4])
            self.up1 = Up(c[4], c[3], c[3], bilinear)
            self.up2 = Up(c[3], c[2], c[2], bilinear)
            self.up3 = Up(c[2], c[1], c[1], bilinear)
            self.up4 = Up(c[1], c[0], c[0], bilinear)
            self.outc = OutConv(c[0], n_classes)
        else:
            c = [base_channels << i for i in range(levels)]
            self.inc = DoubleConv(n_channels, c[0])
            for i in range(1, levels):
                setattr(self, f"down{i}", Down(c[i - 1], c[i]))
            for j in range(1, levels):
                cp = c[levels - j]
                setattr(self, f"up{j}", Up(cp, cp // 2, cp // 2, bilinear))
            self.outc = OutConv(c[0], n_classes)
        self._plans: dict = {}
        self._weights_epoch = 0     # bumped by FusedSGD: raw in-place updates bypass torch versions
        self._stage_hook: Optional[Callable] = None
        self._backward_done_hook: Optional[Callable] = None

    # -- reference helper kept for API parity (models/unet_model.py:88-102) -------------------
    def _center_crop(self, feature_map, target_size):
        _, _, h, w = feature_map.size()
        th, tw = target_size
        hs, ws = max(0, (h - th) // 2), max(0, (w - tw) // 2)
        return feature_map[:, :, hs:hs + th, ws:ws + tw]

    # -- parameter / buffer traversal in the library's canonical order --------------------------
    def _blocks(self):
        enc = [self.inc] + [getattr(self, f"down{i}").maxpool_conv[1] for i in range(1, self.levels)]
        ups = [getattr(self, f"up{j}") for j in range(1, self.levels)]
        return enc, ups

    def _ordered_params(self):
        enc, ups = self._blocks()
        out = []

        def dc(block):
            seq = block.double_conv
            for conv, bn in ((seq[0], seq[1]), (seq[3], seq[4])):
                out.extend([conv.weight, conv.bias, bn.weight, bn.bias])

        for b in enc:
            dc(b)
        for u in ups:
            if not self.bilinear:      # nn.Upsample has no parameters
                out.extend([u.up.weight, u.up.bias])
            dc(u.conv)
        out.extend([self.outc.conv.weight, self.outc.conv.bias])
        return out

    def _ordered_bns(self):
        enc, ups = self._blocks()
        out = []
        for b in enc + [u.conv for u in ups]:
            out.extend([b.double_conv[1], b.double_conv[4]])
        return out

    def _check_bn_modules(self, bns, training: bool) -> None:
        """The library implements nn.BatchNorm2d as the reference configures it (running statistics,
        exponential moving average, the whole net in one mode); anything else must not run silently
        with different semantics."""
        for m in bns:
            if not m.track_running_stats or m.running_mean is None:
                raise RuntimeError("UNet (B200): BatchNorm2d(track_running_stats=False) is not supported")
            if m.momentum is None:
                raise RuntimeError("UNet (B200): BatchNorm2d(momentum=None) (cumulative moving "
                                   "average) is not supported; use a float momentum")
            if not m.affine:
                raise RuntimeError("UNet (B200): BatchNorm2d(affine=False) is not supported")
            if m.training != training:
                raise RuntimeError("UNet (B200): a BatchNorm2d layer is in "
                                   f"{'train' if m.training else 'eval'} mode while the network runs in "
                                   f"{'train' if training else 'eval'} mode; per-layer frozen BatchNorm "
                                   "is not supported (call model.train() / model.eval() on the whole net)")

    def invalidate_weights(self) -> None:
        """Force a refresh of the packed bf16 operand caches at the next forward. Needed after
        writing parameters in a way torch's version counters do not see (``p.data`` assignment /
        in-place edits through ``.data``, raw-pointer writers); optimizers, ``load_state_dict`` and
        ordinary in-place tensor ops are picked up automatically."""
        self._weights_epoch += 1

    # plans hold ctypes handles of device arenas: they are per-process caches, not state
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_plans"] = {}
        state.pop("_tile_sessions", None)
        state["_stage_hook"] = None
        state["_backward_done_hook"] = None
        state.pop("_last_train_key", None)
        return state

    def __deepcopy__(self, memo):
        import copy

        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _plan_for(self, x: torch.Tensor, training: bool) -> _Plan:
        n, c, h, w = x.shape
        return self._plan_for_shape(n, c, h, w, training, x.device)

    def _plan_for_shape(self, n, c, h, w, training: bool, device) -> _Plan:
        key = (n, c, h, w, training, device.index)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:  # keep the arena bounded when shapes vary
                self._plans.pop(next(iter(self._plans)))
            plan = _Plan(n, c, h, w, self.base_channels, self.levels, self.n_classes, training,
                         device, bilinear=self.bilinear)
            self._plans[key] = plan
        bns = self._ordered_bns()
        self._check_bn_modules(bns, training)
        plan.bind(self._ordered_params(), bns, self._weights_epoch)
        if training:
            self._last_train_key = key
        return plan

    def _latest_training_plan(self):
        return self._plans.get(getattr(self, "_last_train_key", None))

    def _check_input(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("UNet expects a (N, C, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("UNet (B200) runs only on a CUDA device (sm_100a); move the model "
                               "and input to 'cuda' — there is no CPU fallback")
        if x.shape[1] != self.n_channels:
            raise ValueError(f"expected {self.n_channels} input channels, got {x.shape[1]}")
        p = self.outc.conv.weight
        if p.device != x.device or p.dtype != torch.float32:
            raise RuntimeError("UNet (B200): parameters must be fp32 on the same CUDA device as the "
                               "input (bf16 operand copies are kept internally)")
        return x.contiguous().float()

    def forward(self, x):
        x = self._check_input(x)
        if x.requires_grad and self.training and torch.is_grad_enabled():
            # the first layer's data gradient is never computed (no reference caller needs it,
            # SURVEY §2.3); returning None silently would be a quiet wrong answer
            raise RuntimeError("UNet (B200): the gradient with respect to the input image is not "
                               "computed; detach the input (x.detach()) before the forward pass")
        with torch.cuda.device(x.device):
            if self.training and torch.is_grad_enabled():
                return _UNetFunction.apply(self, x, *self._ordered_params())
            plan = self._plan_for(x, training=self.training)
            logits, _ = plan.forward(x)
            return logits

    @torch.no_grad()
    def predict_mask(self, x):
        """Eval-mode forward returning (logits, uint8 mask) with mask = 255 where
        softmax(logits)[:,1] > 0.5 (reference scripts/predict.py:81-92), fused in the head kernel."""
        if self.training:
            raise RuntimeError("predict_mask needs model.eval()")
        x = self._check_input(x)
        with torch.cuda.device(x.device):
            plan = self._plan_for(x, training=False)
            return plan.forward(x, want_mask=True)

    # -- data-parallel support -----------------------------------------------------------------
    def set_backward_hooks(self, stage_hook: Optional[Callable], done_hook: Optional[Callable]):
        """stage_hook(stage_index, flat_fp32_grad_slice) runs right after the kernels of a backward
        stage were enqueued; done_hook() after the last stage (see parallel.py)."""
        self._stage_hook = stage_hook
        self._backward_done_hook = done_hook

    def arena_bytes(self) -> int:
        return sum(p.device_bytes() for p in self._plans.values())
