"""Fused SGD for the B200 U-Net.

``FusedSGD(model, lr, momentum, ...)`` implements ``torch.optim.SGD``'s update rule (the reference
trains with ``optim.SGD(model.parameters(), lr=1e-4, momentum=0.99)``, scripts/train.py:97) in one
multi-tensor CUDA kernel that also emits the bf16 tap-major operand copies the convolution kernels
read, so the separate weight re-packing pass of the next forward disappears. State layout
(``state[p]['momentum_buffer']``) and ``state_dict`` match ``torch.optim.SGD``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0,
                 nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self._model = model
        params = model._ordered_params()
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                        nesterov=nesterov)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedSGD handles the UNet's parameters as a single group")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        params = group["params"]
        model = self._model
        plan = model._latest_training_plan()
        if plan is None:
            raise RuntimeError("FusedSGD.step: run a training-mode forward/backward first")
        grads, bufs, first = [], [], False
        for p in params:
            if p.grad is None:
                raise RuntimeError("FusedSGD.step: every parameter needs a gradient")
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            grads.append(g)
            if group["momentum"] != 0:
                st = self.state[p]
                if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                    st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    first = True
                bufs.append(st["momentum_buffer"])
        n = len(params)
        plan.bind_pointers_only(params, model._ordered_bns())
        garr = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
        barr = (C.c_void_p * n)(*[b.data_ptr() for b in bufs]) if bufs else None
        with torch.cuda.device(params[0].device):
            check(plan.lib.ub_plan_sgd_step(plan.handle, garr, barr, float(group["lr"]),
                                            float(group["momentum"]), float(group["dampening"]),
                                            float(group["weight_decay"]), int(group["nesterov"]),
                                            int(first),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "ub_plan_sgd_step")
        model._weights_epoch += 1          # other plans (e.g. the eval plan) must re-pack lazily
        plan.mark_packed(params, model._weights_epoch)
        return loss
