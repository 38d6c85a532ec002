"""Drop-in PatchGANDiscriminator (reference model/discriminator.py:9-84): same constructor, module tree
(`model.first`, `model.middle.N`, `model.last`) and state_dict; forward runs on libcgan3d kernels."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor, nn

from .. import _lib, ops
from .blocks import ConvBlock, from_channels_last, to_channels_last
from .utils import convolution_output_shape


class PatchGANDiscriminator(nn.Module):
    def __init__(self, channels_in: int, init_channels_out: int, discriminator_depth: int, is_2D: bool = False,
                 kernel_size: int = 4, padding: int = 1, norm_layer: Optional[nn.Module] = None, **kwargs):
        super().__init__()
        if is_2D:
            raise NotImplementedError("2D variant is outside the B200 hot path (SURVEY §8f rank 4)")
        self.compute_dtype = kwargs.get("compute_dtype", torch.float32)
        # Critics without BatchNorm (norm_layer = Identity or LayerNorm: the WGAN-GP configurations) run through the
        # twice-differentiable conv Functions, so that wgan_gradient_penalty can differentiate their input gradient.
        self.twice_differentiable = norm_layer in (nn.Identity, nn.LayerNorm)
        stride = 2
        model = [("first", ConvBlock(is_2D, channels_in, init_channels_out, kernel_size, stride=stride, padding=padding,
                                     norm_layer=nn.Identity, activation_fn=nn.LeakyReLU, **kwargs))]
        middle = []
        out_ = init_channels_out
        kwargs = kwargs.copy()
        if ps := kwargs.get("patch_size"):  # per-layer LayerNorm shapes (reference discriminator.py:41-54)
            kwargs["patch_size"] = convolution_output_shape(ps, init_channels_out, kernel_size, padding, stride)
        for n in range(discriminator_depth):
            in_ = min(2 ** n, 8) * init_channels_out
            out_ = min(2 ** (n + 1), 8) * init_channels_out
            if ps := kwargs.get("patch_size"):
                kwargs["patch_size"] = convolution_output_shape(ps, out_, kernel_size, padding, stride)
            middle.append(ConvBlock(is_2D, in_, out_, kernel_size, stride=stride, padding=padding, norm_layer=norm_layer,
                                    activation_fn=nn.LeakyReLU, **kwargs))
        model.append(("middle", nn.Sequential(*middle)))
        model.append(("last", nn.Conv3d(out_, 1, kernel_size=kernel_size, stride=1, padding=padding)))
        self.model = nn.Sequential(OrderedDict(model))
        self._last_spec = ops.ConvSpec(transposed=False, cin=out_, cout=1, k=kernel_size, stride=1, pad=padding)

    def set_compute_dtype(self, dtype: torch.dtype) -> "PatchGANDiscriminator":
        self.compute_dtype = dtype
        for m in self.modules():
            if isinstance(m, ConvBlock):
                m.compute_dtype = dtype
        return self

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() != 5:
            raise ValueError(f"expected [B, C, W, H, D], got {tuple(x.shape)}")
        h = to_channels_last(x.float())
        last = self.model.last
        if self.twice_differentiable:
            h = self.model.first.forward_cl_differentiable(h)
            for blk in self.model.middle:
                h = blk.forward_cl_differentiable(h)
            h = ops.conv_differentiable(h.float(), last.weight, self._last_spec, torch.float32) + last.bias
            return from_channels_last(h)
        h = self.model.first.forward_cl(h)
        for blk in self.model.middle:
            h = blk.forward_cl(h)
        # the 1-channel logits map is tiny: keep it (and the loss means built on it) in fp32
        cfg = ops.BlockCfg(spec=self._last_spec, act=_lib.ACT_NONE, dtype=torch.float32)
        h = ops.ConvBlockFn.apply(h, last.weight, last.bias, None, None, None, None, None, None, cfg)
        return from_channels_last(h)
