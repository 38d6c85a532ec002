"""Shape arithmetic kept bit-compatible with reference model/utils.py:47-105."""
from __future__ import annotations

from typing import Callable, List, Optional, Union

import numpy as np
import torch
from torch import Tensor, nn


def wgan_gradient_penalty(real_batch: Tensor, fake_batch: Tensor, critic: nn.Module, device: Union[torch.device, str] = "cpu",
                          lambda_: float = 10, rng: Optional[np.random.Generator] = None,
                          eps_fn: Optional[Callable[[int], Tensor]] = None) -> Tensor:
    """lambda * mean((||d critic(x~) / d x~||_2 - 1)^2) on random interpolates x~ = eps * real + (1 - eps) * fake, one eps per
    sample (reference model/utils.py:12-41).  The critic must be twice differentiable: here that is a critic without
    BatchNorm, whose convolutions are ops.ConvGatherFn / ConvScatterFn / ConvWgradFn in every derivative order.
    `eps_fn(n)` (ours, optional) supplies the n interpolation coefficients; default torch.rand on `device`, as the reference."""
    interp_sample_size, *t_shape = real_batch.shape
    if len(real_batch) != len(fake_batch):
        interp_sample_size = min(len(real_batch), len(fake_batch))
        rng = rng or np.random.default_rng()
        real_batch = real_batch[rng.integers(len(real_batch), size=interp_sample_size)]
        fake_batch = fake_batch[rng.integers(len(fake_batch), size=interp_sample_size)]
    shape = (interp_sample_size,) + (1,) * len(t_shape)
    eps = torch.rand(shape, device=device) if eps_fn is None else eps_fn(interp_sample_size).reshape(shape).to(device)
    eps = eps.expand_as(real_batch)
    interpolation = eps * real_batch + (1 - eps) * fake_batch
    if not interpolation.requires_grad:
        interpolation.requires_grad_(True)
    critic_logits = critic(interpolation)
    gradients, *_ = torch.autograd.grad(outputs=critic_logits, inputs=interpolation, grad_outputs=torch.ones_like(critic_logits),
                                        create_graph=True)
    gradients_norm = gradients.reshape(gradients.shape[0], -1).norm(2, dim=-1)
    return lambda_ * (gradients_norm - 1).square().mean()


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
