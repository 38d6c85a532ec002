"""Losses with the reference's class names and call signatures (reference model/loss.py); the
reductions and their gradients are fused single-pass kernels (cgan3d_gen_loss_*, cgan3d_mean)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import ops


class ZNCCLoss(nn.Module):
    """-(cov(s,t) / (std(s) std(t) + 1e-8)) over the whole batch tensor, unbiased std with the reference's
    StableStd backward (loss.py:11-41)."""

    def forward(self, source: Tensor, target: Tensor) -> Tensor:
        return ops.GenLossFn.apply(source, target, None, 0.0, 0.0, 1.0, 0.0)[0]


class HULoss(nn.Module):
    """Masked squared hinge outside [min_HU, max_HU] divided by (mask.sum() + 1e-8) (loss.py:44-71).
    `patch_size` is accepted for signature compatibility; the bounds are scalars here."""

    def __init__(self, min_HU_contstraint: float, max_HU_constraint: float, patch_size=None):
        super().__init__()
        self.lo = float(min_HU_contstraint)
        self.hi = float(max_HU_constraint)
        self.patch_size = patch_size

    def forward(self, batch: Tensor, mask: Tensor) -> Tensor:
        return ops.GenLossFn.apply(batch, batch, mask, self.lo, self.hi, 0.0, 1.0)[1]


def fused_similarity_and_hu(opt_hat: Tensor, subopt: Tensor, mask: Tensor, hu: HULoss, w_sim: float, w_hu: float
                            ) -> Tuple[Tensor, Tensor]:
    """(w_sim * ZNCC(opt_hat, subopt), w_hu * HU(opt_hat, mask)) from ONE pass over the tensors
    (reference trainer/Trainer.py:152-153 makes ~12 passes)."""
    out = ops.GenLossFn.apply(opt_hat, subopt, mask, hu.lo, hu.hi, float(w_sim), float(w_hu))
    return out[0], out[1]


class WassersteinLoss(nn.Module):
    """mean(fake) [- mean(real)] (loss.py:74-80)."""

    @staticmethod
    def forward(fake: Tensor, real: Optional[Tensor] = None) -> Tensor:
        ret = ops.MeanFn.apply(fake, 1.0)
        if real is not None:
            ret = ret - ops.MeanFn.apply(real, 1.0)
        return ret
