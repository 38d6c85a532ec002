"""Drop-in PatchGANDiscriminator (reference model/discriminator.py:9-84): same constructor, module tree
(`model.first`, `model.middle.N`, `model.last`) and state_dict; forward runs on libcgan3d kernels."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor, nn

from .. import _lib, ops
from .blocks import ConvBlock, from_channels_last, to_channels_last


class PatchGANDiscriminator(nn.Module):
    def __init__(self, channels_in: int, init_channels_out: int, discriminator_depth: int, is_2D: bool = False,
                 kernel_size: int = 4, padding: int = 1, norm_layer: Optional[nn.Module] = None, **kwargs):
        super().__init__()
        if is_2D:
            raise NotImplementedError("2D variant is outside the B200 hot path (SURVEY §8f rank 4)")
        if kwargs.get("patch_size") is not None or norm_layer == nn.LayerNorm:
            raise NotImplementedError("LayerNorm critic (gp_layernorm.py) is a 'next' row (SURVEY §8f rank 4)")
        self.compute_dtype = kwargs.get("compute_dtype", torch.float32)
        stride = 2
        model = [("first", ConvBlock(is_2D, channels_in, init_channels_out, kernel_size, stride=stride, padding=padding,
                                     norm_layer=nn.Identity, activation_fn=nn.LeakyReLU, **kwargs))]
        middle = []
        out_ = init_channels_out
        for n in range(discriminator_depth):
            in_ = min(2 ** n, 8) * init_channels_out
            out_ = min(2 ** (n + 1), 8) * init_channels_out
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
        h = self.model.first.forward_cl(h)
        for blk in self.model.middle:
            h = blk.forward_cl(h)
        # the 1-channel logits map is tiny: keep it (and the loss means built on it) in fp32
        cfg = ops.BlockCfg(spec=self._last_spec, act=_lib.ACT_NONE, dtype=torch.float32)
        last = self.model.last
        h = ops.ConvBlockFn.apply(h, last.weight, last.bias, None, None, None, None, None, None, cfg)
        return from_channels_last(h)
