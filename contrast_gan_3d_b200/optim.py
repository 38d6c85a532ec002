"""FusedAdam: torch.optim.Adam's single-tensor math (defaults eps=1e-8, no weight decay, no amsgrad) as one
libcgan3d kernel per parameter, with the WGAN weight clip of reference trainer/Trainer.py:136-138 fused in."""
from __future__ import annotations

import torch
from torch.optim import Optimizer

from . import ops


class FusedAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip: float = 0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, clip=clip)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None, clip: float | None = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            c = group["clip"] if clip is None else clip
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                ops.adam_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"], st["step"],
                              c or 0.0)
        return loss
