"""Drop-in ResnetGenerator (reference model/generator.py:9-90): same constructor, module tree
(`model.first`, `model.downsampling.N`, `model.resnet_backbone.N.block0/1`, `model.upsampling.N`,
`model.last_conv`, `model.tanh`) and state_dict; forward runs on libcgan3d kernels."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib, ops
from .blocks import ConvBlock, ResNetBlock, to_channels_last


class ResnetGenerator(nn.Module):
    def __init__(self, n_resnet_blocks: int, n_updownsample_blocks: int, init_channels_out: int, is_2D: bool = False,
                 resnet_dropout_prob: float = 0.0, resnet_padding_mode: str = "zeros",
                 compute_dtype: torch.dtype = torch.float32):
        assert n_resnet_blocks > 0
        super().__init__()
        if is_2D:
            raise NotImplementedError("2D variant is outside the B200 hot path (SURVEY §8f rank 4)")
        kw = {"compute_dtype": compute_dtype}
        first_and_last_common = {"kernel_size": 7, "padding_mode": "reflect", "padding": 3}
        model = [("first", ConvBlock(is_2D, 1, init_channels_out, **first_and_last_common, **kw))]
        downsampling = []
        dim_out = init_channels_out
        for i in range(n_updownsample_blocks):
            dim_in = init_channels_out * 2 ** i
            dim_out = dim_in * 2
            downsampling.append(ConvBlock(is_2D, dim_in, dim_out, kernel_size=3, stride=2, padding=1, **kw))
        model.append(("downsampling", nn.Sequential(*downsampling)))
        resnet_blocks = [ResNetBlock(is_2D, dim_out, dim_out, dropout_prob=resnet_dropout_prob,
                                     padding_mode=resnet_padding_mode, **kw) for _ in range(n_resnet_blocks)]
        model.append(("resnet_backbone", nn.Sequential(*resnet_blocks)))
        upsampling = []
        for i in range(n_updownsample_blocks, 0, -1):
            dim_in = init_channels_out * 2 ** i
            dim_out = int(dim_in / 2)
            upsampling.append(ConvBlock(is_2D, dim_in, dim_out, kernel_size=3, stride=2, padding=1, output_padding=1,
                                        upsample=True, **kw))
        model.append(("upsampling", nn.Sequential(*upsampling)))
        model.append(("last_conv", nn.Conv3d(init_channels_out, 1, **first_and_last_common, bias=True)))
        model.append(("tanh", nn.Tanh()))
        self.model = nn.Sequential(OrderedDict(model))
        self.compute_dtype = compute_dtype
        self._tail_spec = ops.ConvSpec(transposed=False, cin=init_channels_out, cout=1, k=7, stride=1, pad=3, reflect=True)

    def set_compute_dtype(self, dtype: torch.dtype) -> "ResnetGenerator":
        self.compute_dtype = dtype
        for m in self.modules():
            if isinstance(m, ConvBlock):
                m.compute_dtype = dtype
        return self

    def _trunk(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected [B, 1, W, H, D], got {tuple(x.shape)}")
        h = to_channels_last(x.float())
        h = self.model.first.forward_cl(h)
        for blk in self.model.downsampling:
            h = blk.forward_cl(h)
        for blk in self.model.resnet_backbone:
            h = blk.forward_cl(h)
        ups = list(self.model.upsampling)
        for i, blk in enumerate(ups):
            # the last up-sampling block writes its output already reflection-padded for last_conv (reference
            # generator.py:77-83: Conv3d(padding=3, padding_mode="reflect")): `_tail` then takes it as is
            h = blk.forward_cl(h, pad_out=self._tail_spec.pad if i == len(ups) - 1 else 0)
        return h

    def _tail(self, h, subopt: Optional[torch.Tensor]):
        cfg = ops.BlockCfg(spec=self._tail_spec, act=_lib.ACT_TANH, dtype=self.compute_dtype,
                           pre_padded=len(self.model.upsampling) > 0)
        return ops.GenTailFn.apply(h, self.model.last_conv.weight, self.model.last_conv.bias, subopt, cfg)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        att, _ = self._tail(self._trunk(x), None)
        return att

    def forward_corrected(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(attenuation, x - attenuation) with the subtraction fused into the tanh epilogue
        (reference trainer/Trainer.py:170-171, eval/CCTAContrastCorrector.py:79)."""
        x = x.float().contiguous()
        return self._tail(self._trunk(x), x)
