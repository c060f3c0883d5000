"""Fused AdamW for the drop-in.

The reference builds its optimizer reflectively — `getattr(torch.optim, config.training.optimizer.name)
(trainable_params, **params)` (utils/train_utils.py:25-29) — and steps it through a GradScaler
(`scaler.step(optimizer)`, utils/trainers/simmim_trainer.py:69-71, base_trainer.py:44). `FusedAdamW`
is a `torch.optim.Optimizer` with torch.optim.AdamW's constructor, state layout
(`step` / `exp_avg` / `exp_avg_sq` per parameter) and arithmetic, so it is selectable by config
alone: importing `vit_core` registers it as `torch.optim.VitsslAdamW`
(`training.optimizer.name=VitsslAdamW`).

One multi-tensor kernel launch per ~36 parameter tensors (csrc/optimizer.cu) does the whole update:
GradScaler unscale and skipped-step handling on the device (the class sets
`_step_supports_amp_scaling`, so `scaler.step` hands over `grad_scale` / `found_inf` instead of
unscaling in a separate pass), decoupled weight decay, moment updates, bias-corrected step, AND
the bf16 GEMM-operand shadows the next forward reads — replacing torch's optimizer kernels plus our
own multi-tensor cast pass. No host synchronisation anywhere.
"""
from __future__ import annotations

from typing import Iterable

import torch

from ._backend import functional as Fb
from ._backend import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False, *, maximize: bool = False,
                 foreach=None, capturable: bool = False, differentiable: bool = False, fused=None):
        if isinstance(lr, torch.Tensor):
            lr = float(lr)
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad or maximize or differentiable:
            raise ValueError("FusedAdamW implements plain AdamW (amsgrad / maximize / differentiable are unsupported)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=True)
        super().__init__(params, defaults)
        # torch.amp.GradScaler.step: hand grad_scale / found_inf to step() instead of unscaling first
        self._step_supports_amp_scaling = True

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        for group in self.param_groups:
            ps, gs, ms, vs, shs, steps = [], [], [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedAdamW needs fp32 CUDA parameters (no CPU fallback)")
                st = self._init_state(p)
                g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.float().contiguous()
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters")
                ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
                steps.append(st["step"])
                shs.append(Fb.shadow_destination(p))
            if not ps:
                continue
            b1, b2 = group["betas"]
            dev = ps[0].device
            sc = grad_scale.to(dev, torch.float32).reshape(()) if grad_scale is not None else None
            fi = found_inf.to(dev, torch.float32).reshape(()) if found_inf is not None else None
            ops.adamw_step([p.data for p in ps], gs, ms, vs, shs, steps, group["lr"], b1, b2, group["eps"],
                           group["weight_decay"], sc, fi)
            # the kernel wrote through raw pointers: make the update visible to version-based caches,
            # then re-sign the shadows it refreshed in the same pass
            torch.autograd.graph.increment_version(ps)
            Fb.mark_shadows_fresh([p for p, s_ in zip(ps, shs) if s_ is not None])
        return loss


def register() -> None:
    """Expose the optimizer where the reference's reflective factory looks (utils/train_utils.py:27)."""
    if not hasattr(torch.optim, "VitsslAdamW"):
        torch.optim.VitsslAdamW = FusedAdamW
