"""Shape arithmetic kept bit-compatible with reference model/utils.py:47-105."""
from __future__ import annotations

from typing import List, Optional

from torch import nn


def convolution_output_shape(dims: List[int], c_out: int, kernel_size: int, padding: int, stride: int, dilation: int = 1,
                             transpose_output_padding: Optional[int] = None) -> List[int]:
    """[c_out, *spatial_out]; the reference divides in floating point and truncates with int() (utils.py:56-70)."""
    if transpose_output_padding is None:
        def f(x):
            return int((x + 2 * padding - dilation * (kernel_size - 1) - 1) / stride + 1)
    else:
        def f(x):
            return int((x - 1) * stride - 2 * padding + dilation * (kernel_size - 1) + transpose_output_padding + 1)
    return [c_out] + [f(d) for d in dims[1:]]


def compute_convolution_filters_shape(model: nn.Module, input_shape, show: bool = True) -> List[int]:
    """Walk the conv layers of `model` and return the final [C, *spatial] (utils.py:73-95)."""
    lines = [f"Input shape: {list(input_shape)}"]
    for n, m in model.named_modules():
        if type(m) in (nn.Conv3d, nn.Conv2d, nn.ConvTranspose3d, nn.ConvTranspose2d):
            kw = {}
            if isinstance(m, (nn.ConvTranspose3d, nn.ConvTranspose2d)):
                kw = {"transpose_output_padding": m.output_padding[0]}
            input_shape = convolution_output_shape(input_shape, m.out_channels, m.kernel_size[0], m.padding[0],
                                                   m.stride[0], **kw)
            lines.append(f"{n:<40} -> {str(input_shape):<22} # params: {count_parameters(m)}")
    if show:
        print("\n".join(lines))
    return input_shape


def count_parameters(model: nn.Module, print: bool = False) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
