"""CPU oracle for the contrast-gan-3D hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference arm may
import this module.  It is the checker, never the thing measured or shipped; nothing
under `contrast_gan_3d_b200/` imports it.

What it is: a functional restatement, in plain fp32 PyTorch-on-CPU tensor arithmetic,
of the reference's generator / critic / losses / train step / patch sampler / tiler.
The reference itself contributes only graph structure and three loss formulas; the
arithmetic lives in ATen (torch 2.11.0, pinned by this image and present on the GPU
box), so the oracle calls the same ATen CPU primitives (`F.conv3d`,
`F.conv_transpose3d`, `F.batch_norm`, `F.pad(mode="reflect")`) through a flat
name->tensor parameter dictionary instead of `nn.Module`s.

Pinning: `tests/golden/make_golden.py` imports the UNMODIFIED reference
(`/root/reference`, via `oracle/ref_shim.py`) in the authoring container and dumps
seeded inputs/outputs/losses to `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks every function below against those fixtures.  The patch sampler and grid
tiler restate third-party `batchgenerators` / `patchly` functions that are NOT
installed here and are NOT vendored by the reference: those two are
"parity unpinned" beyond the reference's own call sites.

All `file:line` citations are relative to `/root/reference/contrast_gan_3D/`.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5  # torch.nn.BatchNorm3d default, used by model/blocks.py:26-27,45
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# architecture description (model/generator.py:31-87, model/discriminator.py:23-81)
# --------------------------------------------------------------------------------------
def generator_layers(n_resnet_blocks=4, n_updownsample_blocks=2, init_channels_out=16):
    """Ordered list of conv layers of ResnetGenerator (generator.py:31-85).

    Each entry: dict(name, kind 'conv'|'convT'|'res0'|'res1', cin, cout, k, stride, pad,
    pad_mode, out_pad, norm (bool), act ('relu'|'none'|'tanh'), bias (bool)).
    """
    L = []
    c = init_channels_out
    L.append(dict(name="model.first", kind="conv", cin=1, cout=c, k=7, stride=1, pad=3,
                  pad_mode="reflect", out_pad=0, norm=True, act="relu", bias=False))
    for i in range(n_updownsample_blocks):  # generator.py:40-46
        cin = init_channels_out * 2 ** i
        L.append(dict(name=f"model.downsampling.{i}", kind="conv", cin=cin, cout=2 * cin, k=3,
                      stride=2, pad=1, pad_mode="zeros", out_pad=0, norm=True, act="relu", bias=False))
        c = 2 * cin
    for i in range(n_resnet_blocks):  # generator.py:49-57, blocks.py:68-88
        L.append(dict(name=f"model.resnet_backbone.{i}.block0", kind="res0", cin=c, cout=c, k=3, stride=1,
                      pad=1, pad_mode="zeros", out_pad=0, norm=True, act="none", bias=False))
        L.append(dict(name=f"model.resnet_backbone.{i}.block1", kind="res1", cin=c, cout=c, k=3, stride=1,
                      pad=1, pad_mode="zeros", out_pad=0, norm=True, act="relu", bias=False))
    j = 0
    for i in range(n_updownsample_blocks, 0, -1):  # generator.py:60-76
        cin = init_channels_out * 2 ** i
        L.append(dict(name=f"model.upsampling.{j}", kind="convT", cin=cin, cout=cin // 2, k=3, stride=2,
                      pad=1, pad_mode="zeros", out_pad=1, norm=True, act="relu", bias=False))
        j += 1
    L.append(dict(name="model.last_conv", kind="conv", cin=init_channels_out, cout=1, k=7, stride=1, pad=3,
                  pad_mode="reflect", out_pad=0, norm=False, act="tanh", bias=True))
    return L


def critic_layers(channels_in=1, init_channels_out=8, discriminator_depth=3, kernel_size=4, padding=1, norm="batch"):
    """Ordered conv layers of PatchGANDiscriminator (discriminator.py:23-81).  norm: "batch" (default, BatchNorm3d),
    "identity" (WGAN-GP, experiments/gradient_penalty_conf.py:14: conv bias instead of a norm, blocks.py:34) or "layer"
    (experiments/gp_layernorm.py: LayerNorm over the layer's whole [C, W, H, D] output, elementwise_affine=False)."""
    L = [dict(name="model.first", kind="conv", cin=channels_in, cout=init_channels_out, k=kernel_size,
              stride=2, pad=padding, pad_mode="zeros", out_pad=0, norm=False, act="lrelu", bias=True)]
    out_ = init_channels_out
    for n in range(discriminator_depth):  # discriminator.py:48-67
        in_ = min(2 ** n, 8) * init_channels_out
        out_ = min(2 ** (n + 1), 8) * init_channels_out
        L.append(dict(name=f"model.middle.{n}", kind="conv", cin=in_, cout=out_, k=kernel_size, stride=2,
                      pad=padding, pad_mode="zeros", out_pad=0, norm={"batch": True, "identity": False, "layer": "layer"}[norm],
                      act="lrelu", bias=norm == "identity"))
    L.append(dict(name="model.last", kind="conv", cin=out_, cout=1, k=kernel_size, stride=1, pad=padding,
                  pad_mode="zeros", out_pad=0, norm=False, act="none", bias=True, bare=True))
    return L


def _conv_prefix(layer) -> str:
    # ConvBlock keeps its conv under `.conv` (blocks.py:29); the two bare convs
    # (generator.py:77-83 last_conv, discriminator.py:69-80 last) are direct nn.Conv3d.
    if layer["name"] in ("model.last_conv", "model.last"):
        return layer["name"]
    return layer["name"] + ".conv"


# --------------------------------------------------------------------------------------
# parameter init == torch.nn.Conv3d / ConvTranspose3d.reset_parameters()
# --------------------------------------------------------------------------------------
def _conv_init(shape, fan_in, with_bias, cout) -> Tuple[Tensor, Optional[Tensor]]:
    # kaiming_uniform_(a=sqrt(5)) then uniform bias, exactly torch's expressions so that a
    # seeded run draws the same numbers as the reference's module constructors.
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    std = gain / math.sqrt(fan_in)
    bound = math.sqrt(3.0) * std
    w = torch.empty(shape, dtype=torch.float32).uniform_(-bound, bound)
    b = None
    if with_bias:
        bb = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        b = torch.empty(cout, dtype=torch.float32).uniform_(-bb, bb)
    return w, b


def init_params(layers) -> Tuple["OrderedDict[str, Tensor]", "OrderedDict[str, Tensor]"]:
    """Build (params, buffers) with the reference's state_dict key names.

    Draws from torch's global CPU RNG in module-construction order, so
    `torch.manual_seed(0); init_params(generator_layers()); init_params(critic_layers())`
    reproduces `Trainer.__init__` (trainer/Trainer.py:83,89: G first, then D).
    """
    params, buffers = OrderedDict(), OrderedDict()
    for l in layers:
        k = l["k"]
        pre = _conv_prefix(l)
        if l["kind"] == "convT":
            shape = (l["cin"], l["cout"], k, k, k)  # ConvTranspose3d weight is [Cin,Cout,...]
            fan_in = l["cout"] * k ** 3  # torch uses size(1) * receptive field
        else:
            shape = (l["cout"], l["cin"], k, k, k)
            fan_in = l["cin"] * k ** 3
        w, b = _conv_init(shape, fan_in, l["bias"], l["cout"])
        params[pre + ".weight"] = w
        if b is not None:
            params[pre + ".bias"] = b
        if l["norm"] is True:
            n = l["name"] + ".normalization"
            params[n + ".weight"] = torch.ones(l["cout"])
            params[n + ".bias"] = torch.zeros(l["cout"])
            buffers[n + ".running_mean"] = torch.zeros(l["cout"])
            buffers[n + ".running_var"] = torch.ones(l["cout"])
            buffers[n + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return params, buffers


def state_dict_order(layers) -> List[str]:
    """Key order of the reference module's state_dict()."""
    keys = []
    for l in layers:
        pre = _conv_prefix(l)
        keys.append(pre + ".weight")
        if l["bias"]:
            keys.append(pre + ".bias")
        if l["norm"] is True:
            n = l["name"] + ".normalization"
            keys += [n + ".weight", n + ".bias", n + ".running_mean", n + ".running_var",
                     n + ".num_batches_tracked"]
    return keys


# --------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------
def _conv_block(x: Tensor, l, params, buffers, train: bool, negative_slope=0.2) -> Tensor:
    """ConvBlock.forward: act(norm(conv(x))) (blocks.py:52-53)."""
    pre = _conv_prefix(l)
    w = params[pre + ".weight"]
    b = params.get(pre + ".bias")
    if l["kind"] == "convT":
        y = F.conv_transpose3d(x, w, b, stride=l["stride"], padding=l["pad"], output_padding=l["out_pad"])
    else:
        if l["pad_mode"] == "reflect":
            p = l["pad"]
            x = F.pad(x, (p,) * 6, mode="reflect")
            y = F.conv3d(x, w, b, stride=l["stride"], padding=0)
        else:
            y = F.conv3d(x, w, b, stride=l["stride"], padding=l["pad"])
    if l["norm"] == "layer":  # nn.LayerNorm(patch_size = [C, W, H, D], elementwise_affine=False), blocks.py:40-45
        y = F.layer_norm(y, y.shape[1:])
    elif l["norm"]:
        n = l["name"] + ".normalization"
        if train:
            buffers[n + ".num_batches_tracked"] += 1
        y = F.batch_norm(y, buffers[n + ".running_mean"], buffers[n + ".running_var"], params[n + ".weight"],
                         params[n + ".bias"], training=train, momentum=BN_MOMENTUM, eps=BN_EPS)
    a = l["act"]
    if a == "relu":
        y = F.relu(y)
    elif a == "lrelu":
        y = F.leaky_relu(y, negative_slope)
    elif a == "tanh":
        y = torch.tanh(y)
    return y


def generator_forward(params, buffers, x: Tensor, layers=None, train: bool = True) -> Tensor:
    """ResnetGenerator.forward (generator.py:89-90); residual per blocks.py:87-88."""
    layers = layers or generator_layers()
    skip = None
    for l in layers:
        if l["kind"] == "res0":
            skip = x
        x = _conv_block(x, l, params, buffers, train)
        if l["kind"] == "res1":
            x = skip + x
    return x


def critic_forward(params, buffers, x: Tensor, layers=None, train: bool = True, negative_slope=0.2) -> Tensor:
    """PatchGANDiscriminator.forward (discriminator.py:83-84)."""
    layers = layers or critic_layers()
    for l in layers:
        x = _conv_block(x, l, params, buffers, train, negative_slope)
    return x


# --------------------------------------------------------------------------------------
# losses (model/loss.py)
# --------------------------------------------------------------------------------------
class _StableStd(torch.autograd.Function):
    """loss.py:11-29: unbiased std with the reference's hand-written backward."""

    @staticmethod
    def forward(ctx, t):
        res = torch.std(t.detach())
        ctx.save_for_backward(t.detach(), res)
        return res

    @staticmethod
    def backward(ctx, g):
        t, res = ctx.saved_tensors
        return (2.0 / (t.numel() - 1.0)) * (g / (res * 2 + 1e-6)) * (t - t.mean())


def zncc_loss(source: Tensor, target: Tensor) -> Tensor:
    """ZNCCLoss.forward, loss.py:37-41 (means/std over the WHOLE batch tensor)."""
    cc = ((source - source.mean()) * (target - target.mean())).mean()
    std = _StableStd.apply(source) * _StableStd.apply(target)
    return -(cc / (std + 1e-8))


def hu_loss(batch: Tensor, mask: Tensor, lo: float, hi: float) -> Tensor:
    """HULoss.forward, loss.py:64-71: masked squared hinge outside [lo, hi] / (sum(mask)+1e-8)."""
    lo_t = torch.full_like(batch, lo)
    hi_t = torch.full_like(batch, hi)
    below = (torch.minimum(batch, lo_t) - lo_t) ** 2
    above = (torch.maximum(batch, hi_t) - hi_t) ** 2
    loss = (below + above) * mask
    return loss.sum() / (mask.sum() + torch.tensor(1e-8))


def wasserstein_loss(fake: Tensor, real: Optional[Tensor] = None) -> Tensor:
    """WassersteinLoss.forward, loss.py:75-80."""
    r = torch.mean(fake)
    if real is not None:
        r = r - torch.mean(real)
    return r


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam single-tensor algorithm, defaults eps=1e-8, no weight decay)
# --------------------------------------------------------------------------------------
class AdamState:
    def __init__(self, params: "OrderedDict[str, Tensor]", lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def apply(self, params, grads, lr: Optional[float] = None):
        lr = self.lr if lr is None else lr
        b1, b2 = self.betas
        self.step += 1
        bc1 = 1 - b1 ** self.step
        bc2 = 1 - b2 ** self.step
        step_size = lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        for k, p in params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.m[k].lerp_(g, 1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / bc2_sqrt).add_(self.eps)
            p.data.addcdiv_(self.m[k], denom, value=-step_size)


class RMSpropState:
    """torch.optim.RMSprop single-tensor algorithm at its defaults (alpha 0.99, eps 1e-8, no momentum, not centered), the
    optimizer of experiments/rmsprop_conf.py:8-9; same interface as AdamState."""

    def __init__(self, params, lr=2e-4, alpha=0.99, eps=1e-8):
        self.lr, self.alpha, self.eps = lr, alpha, eps
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def apply(self, params, grads, lr: Optional[float] = None):
        lr = self.lr if lr is None else lr
        for k, p in params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.v[k].mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)
            p.data.addcdiv_(g, self.v[k].sqrt().add_(self.eps), value=-lr)


def multistep_lr(base_lr: float, milestones: Sequence[int], gamma: float, n_steps_done: int) -> float:
    """MultiStepLR value after `n_steps_done` scheduler steps (basic_conf.py:35-36,56-58)."""
    return base_lr * gamma ** sum(1 for m in milestones if n_steps_done >= m)


# --------------------------------------------------------------------------------------
# one training iteration (trainer/Trainer.py:108-185), weight-clip WGAN variant
# --------------------------------------------------------------------------------------
class StepState:
    """All mutable state of a Trainer: G/D params+buffers and two Adam states."""

    def __init__(self, seed: Optional[int] = 0, lr=2e-4, betas=(0.5, 0.999), g_layers=None, d_layers=None):
        if seed is not None:
            torch.manual_seed(seed)
        self.g_layers = g_layers or generator_layers()
        self.d_layers = d_layers or critic_layers()
        self.gp, self.gb = init_params(self.g_layers)  # Trainer.py:83 (G first)
        self.dp, self.db = init_params(self.d_layers)  # Trainer.py:89
        self.opt_g = AdamState(self.gp, lr, betas)
        self.opt_d = AdamState(self.dp, lr, betas)
        self.milestones, self.gamma = (6000, 8000), 0.1
        self.sched_steps_g = 0
        self.sched_steps_d = 0


def gradient_penalty(dp, db, d_layers, real: Tensor, fake: Tensor, eps: Tensor, lambda_=10.0) -> Tensor:
    """wgan_gradient_penalty (model/utils.py:12-41) for equal batch sizes; eps = the torch.rand((B,1,1,1,1)) draw."""
    interpolation = eps.expand_as(real) * real + (1 - eps.expand_as(real)) * fake
    if not interpolation.requires_grad:
        interpolation.requires_grad_(True)
    logits = critic_forward(dp, db, interpolation, d_layers)
    grads, = torch.autograd.grad(logits, interpolation, torch.ones_like(logits), create_graph=True)
    norm = grads.view(grads.shape[0], -1).norm(2, dim=-1)
    return lambda_ * (norm - 1).square().mean()


def train_step(st: StepState, opt: Tensor, low: Tensor, high: Tensor, mask_low: Tensor, mask_high: Tensor,
               iteration: int, hu_bounds=(0.18666666666666668, 0.35333333333333333), weight_clip=0.01,
               train_generator_every=1, train_critic_every=1, w_gan=1.0, w_sim=1.0, w_hu=1.0, gp_weight=10.0,
               gp_eps: Optional[Tensor] = None) -> Dict[str, float]:
    """Trainer.train_step (Trainer.py:163-185) + train_critic (:108-142) + train_generator (:144-161)."""
    for p in list(st.gp.values()) + list(st.dp.values()):
        p.requires_grad_(True)
        p.grad = None
    subopt = torch.cat([low, high])  # Trainer.py:166
    attenuation = generator_forward(st.gp, st.gb, subopt, st.g_layers, train=True)  # :170
    opt_hat = subopt - attenuation  # :171
    out: Dict[str, float] = {}
    do_g = iteration % train_generator_every == 0
    if iteration % train_critic_every == 0:
        real_logits = critic_forward(st.dp, st.db, opt, st.d_layers)  # :114
        fake_logits = critic_forward(st.dp, st.db, opt_hat.detach(), st.d_layers)  # :116
        loss_d = w_gan * wasserstein_loss(fake_logits, real_logits)  # :119-121
        if weight_clip is None:  # :122-130 (the penalty's gradient into G is dead: zeroed at :147 before G's backward)
            eps = gp_eps if gp_eps is not None else torch.rand((opt.shape[0], 1, 1, 1, 1))
            loss_d = loss_d + gradient_penalty(st.dp, st.db, st.d_layers, opt, opt_hat.detach(), eps, gp_weight)
        names = list(st.dp.keys())
        grads = torch.autograd.grad(loss_d, [st.dp[k] for k in names])
        lr = multistep_lr(st.opt_d.lr, st.milestones, st.gamma, st.sched_steps_d)
        with torch.no_grad():
            st.opt_d.apply(st.dp, dict(zip(names, grads)), lr)  # :135
            if weight_clip is not None:
                for p in st.dp.values():  # :136-138 (clamps BN gamma/beta too)
                    p.clamp_(-weight_clip, weight_clip)
        st.sched_steps_d += 1  # :139-140
        out["D"] = float(loss_d.detach())
    if do_g:
        mask = torch.cat([mask_low, mask_high])  # :182
        loss_g = w_gan * -wasserstein_loss(critic_forward(st.dp, st.db, opt_hat, st.d_layers))  # :151
        loss_sim = w_sim * zncc_loss(opt_hat, subopt)  # :152
        loss_hu = w_hu * hu_loss(opt_hat, mask, *hu_bounds)  # :153
        full = loss_g + loss_sim + loss_hu
        names = list(st.gp.keys())
        grads = torch.autograd.grad(full, [st.gp[k] for k in names])
        lr = multistep_lr(st.opt_g.lr, st.milestones, st.gamma, st.sched_steps_g)
        with torch.no_grad():
            st.opt_g.apply(st.gp, dict(zip(names, grads)), lr)  # :157
        st.sched_steps_g += 1
        out.update({"G": float(loss_g.detach()), "G-full": float(full.detach()),
                    "sim": float(loss_sim.detach()), "HU": float(loss_hu.detach())})
    for p in list(st.gp.values()) + list(st.dp.values()):
        p.requires_grad_(False)
    return out


def validate(st: StepState, batches: Sequence[Tuple[Tensor, Tensor, Tensor]]) -> Dict[str, float]:
    """Trainer.validate (Trainer.py:247-308): eval-mode networks (BatchNorm uses its running statistics), no_grad;
    `batches` = one (opt, low, high) triple per validation iteration, visited in ScanType order OPT, LOW, HIGH."""
    loss_sim = loss_g = loss_real_c = loss_fake_c = 0.0
    with torch.no_grad():
        for opt, low, high in batches:
            loss_real_c -= float(wasserstein_loss(critic_forward(st.dp, st.db, opt, st.d_layers, train=False)))  # :263-267
            for sample in (low, high):
                sample_hat = sample - generator_forward(st.gp, st.gb, sample, st.g_layers, train=False)  # :269-270
                loss_fake = float(wasserstein_loss(critic_forward(st.dp, st.db, sample_hat, st.d_layers, train=False)))  # :271-272
                loss_fake_c += loss_fake
                loss_g -= loss_fake
                loss_sim += float(zncc_loss(sample_hat, sample))  # :276
    n = len(batches)
    return {"D": (loss_real_c + loss_fake_c) / n, "G": loss_g / (n * 2), "sim": loss_sim / (n * 2)}  # :300-304


# --------------------------------------------------------------------------------------
# data: scaler, synthetic HU patches, crop/pad index law, conv shape arithmetic, tiler
# --------------------------------------------------------------------------------------
def scaler_shift(low: int = -1024, high: int = 1500) -> int:
    """ZeroCenterScaler.__post_init__, data/Scaler.py:26-27."""
    return (high - abs(low)) // 2


def scale_hu(x, low=-1024, high=1500, factor=600):
    """FactorZeroCenterScaler.__call__, data/Scaler.py:41-42."""
    return (x - scaler_shift(low, high)) / factor


def unscale_hu(x, low=-1024, high=1500, factor=600):
    """FactorZeroCenterScaler.unscale, data/Scaler.py:44-45."""
    return x * factor + scaler_shift(low, high)


def synthetic_patches(gen: torch.Generator, shape) -> Tensor:
    """SURVEY §8d synthetic law: x = (clamp(300 N(0,1) + 100, -1024, 1500) - 238) / 600."""
    hu = (torch.randn(shape, generator=gen) * 300 + 100).clamp(-1024, 1500)
    return (hu - 238) / 600


def synthetic_masks(gen: torch.Generator, shape, p=1e-3) -> Tensor:
    return torch.rand(shape, generator=gen) < p


def pad_nd_image_shape(old_shape: Sequence[int], new_shape: Sequence[int]):
    """batchgenerators `pad_nd_image` restated [upstream, unpinned]; call site data/CCTADataLoader.py:83.

    Pads the LAST len(new_shape) axes up to max(new, old): below = diff//2, above = diff//2 + diff%2.
    Returns (padded_shape, [(below, above), ...] for every axis)."""
    old_shape = list(old_shape)
    n = len(new_shape)
    lead = len(old_shape) - n
    tgt = [max(a, b) for a, b in zip(new_shape, old_shape[lead:])]
    pads = [(0, 0)] * lead
    for o, t in zip(old_shape[lead:], tgt):
        d = t - o
        pads.append((d // 2, d // 2 + d % 2))
    return old_shape[:lead] + tgt, pads


def pad_nd_image(img: np.ndarray, new_shape: Sequence[int]) -> np.ndarray:
    _, pads = pad_nd_image_shape(img.shape, new_shape)
    return np.pad(img, pads, mode="constant", constant_values=0)


def random_crop_lbs(spatial_shape: Sequence[int], crop_size: Sequence[int], rs=np.random) -> List[int]:
    """batchgenerators `get_lbs_for_random_crop` with margins 0 restated [upstream, unpinned].

    Per axis (W,H,D order): lb = randint(0, dim - size) (high-exclusive) if dim - size > 0 else (dim-size)//2.
    Draws from the legacy global `np.random` state (NOT the loader's rng)."""
    lbs = []
    for d, c in zip(spatial_shape, crop_size):
        if d - c > 0:
            lbs.append(int(rs.randint(0, d - c)))
        else:
            lbs.append((d - c) // 2)
    return lbs


def crop_with_lbs(vol: np.ndarray, lbs: Sequence[int], crop_size: Sequence[int]) -> np.ndarray:
    """Crop [C, W, H, D] to crop_size at lower bounds lbs; out-of-range parts zero-padded."""
    C = vol.shape[0]
    out = np.zeros((C, *crop_size), dtype=vol.dtype)
    src, dst = [slice(None)], [slice(None)]
    for d, lb, c in zip(vol.shape[1:], lbs, crop_size):
        lo, hi = max(lb, 0), min(lb + c, d)
        src.append(slice(lo, hi))
        dst.append(slice(lo - lb, hi - lb))
    out[tuple(dst)] = vol[tuple(src)]
    return out


def generate_one(ccta_and_seg: np.ndarray, patch_shape: Sequence[int], rs=np.random):
    """CCTADataLoader.generate_one for the 3D sampler (data/CCTADataLoader.py:76-95).

    `ccta_and_seg`: int16 [W,H,D,2] (HU, centerline mask). Returns (scaled patch f32 [1,1,*patch],
    mask f32 [1,1,*patch], lbs)."""
    x = ccta_and_seg[None, None]  # :79
    x = pad_nd_image(x, (*patch_shape, 2))  # :83
    x = x.astype(np.float32)  # :85
    lbs = random_crop_lbs(x.shape[2:5], patch_shape, rs)  # :86-91 (crop_type="random")
    data = crop_with_lbs(x[0, ..., 0], lbs, patch_shape)[None]
    seg = crop_with_lbs(x[0, ..., 1], lbs, patch_shape)[None]
    return scale_hu(data).astype(np.float32), seg, lbs


def convolution_output_shape(dims, c_out, kernel_size, padding, stride, dilation=1, transpose_output_padding=None):
    """model/utils.py:47-70 restated (float division then int(), as the reference does)."""
    if transpose_output_padding is None:
        f = lambda x: int((x + 2 * padding - dilation * (kernel_size - 1) - 1) / stride + 1)
    else:
        f = lambda x: int((x - 1) * stride - 2 * padding + dilation * (kernel_size - 1) + transpose_output_padding + 1)
    return [c_out] + [f(d) for d in dims[1:]]


def grid_tiles(volume_shape: Sequence[int], patch: Sequence[int]) -> List[Tuple[int, int, int]]:
    """patchly GridSampler restated for the divisible case [upstream, unpinned]; call site
    eval/CCTAContrastCorrector.py:63. step = patch; row-major over (x, y, z), z fastest.
    Non-divisible sizes: the last tile is squeezed back inside the volume (SAMPLE_SQUEEZE)."""
    starts = []
    for s, p in zip(volume_shape, patch):
        if s < p:
            raise ValueError("volume smaller than patch")
        a = list(range(0, s - p + 1, p))
        if a[-1] + p < s:
            a.append(s - p)
        starts.append(a)
    return [(x, y, z) for x in starts[0] for y in starts[1] for z in starts[2]]


def correct_scan_3d(params, buffers, ccta: np.ndarray, patch=(128, 128, 128), batch_size=16, layers=None) -> Tensor:
    """CCTAContrastCorrector.__call__/correct_scan_3D (eval/CCTAContrastCorrector.py:60-81,101-106).

    BatchNorm stays in TRAIN mode (the reference never calls .eval()); tiles are averaged where they overlap
    (patchly Aggregator default weights='avg')."""
    layers = layers or generator_layers()
    tiles = grid_tiles(ccta.shape, patch)
    acc = torch.zeros(ccta.shape, dtype=torch.float32)
    cnt = torch.zeros(ccta.shape, dtype=torch.float32)
    with torch.no_grad():
        for i in range(0, len(tiles), batch_size):
            chunk = tiles[i:i + batch_size]
            xs = []
            for (x, y, z) in chunk:
                p = ccta[x:x + patch[0], y:y + patch[1], z:z + patch[2]].astype(np.float32)
                xs.append(torch.from_numpy(scale_hu(p).astype(np.float32))[None])
            xb = torch.stack(xs)
            att = generator_forward(params, buffers, xb, layers, train=True)
            if att.shape[2:] != xb.shape[2:]:  # nn.Upsample(size=inference_patch_size), default mode "nearest" (:42-52)
                att = F.interpolate(att, size=tuple(patch))
            corrected = xb - att
            for (x, y, z), c in zip(chunk, corrected):
                acc[x:x + patch[0], y:y + patch[1], z:z + patch[2]] += c[0]
                cnt[x:x + patch[0], y:y + patch[1], z:z + patch[2]] += 1
    return unscale_hu(acc / cnt)


# --------------------------------------------------------------------------------------
# tiny independent restatements used to pin the ATen primitives themselves (small cases)
# --------------------------------------------------------------------------------------
def naive_conv3d(x: np.ndarray, w: np.ndarray, stride=1, pad=0, pad_mode="zeros") -> np.ndarray:
    """Direct cross-correlation, float64, [B,Cin,X,Y,Z] * [Cout,Cin,k,k,k]."""
    x = x.astype(np.float64)
    w = w.astype(np.float64)
    if pad:
        mode = "reflect" if pad_mode == "reflect" else "constant"
        x = np.pad(x, [(0, 0), (0, 0)] + [(pad, pad)] * 3, mode=mode)
    B, Ci, X, Y, Z = x.shape
    Co, _, k, _, _ = w.shape
    ox, oy, oz = (X - k) // stride + 1, (Y - k) // stride + 1, (Z - k) // stride + 1
    out = np.zeros((B, Co, ox, oy, oz))
    for a in range(k):
        for b in range(k):
            for c in range(k):
                xs = x[:, :, a:a + stride * ox:stride, b:b + stride * oy:stride, c:c + stride * oz:stride]
                out += np.einsum("bixyz,oi->boxyz", xs, w[:, :, a, b, c])
    return out


def naive_conv_transpose3d(x: np.ndarray, w: np.ndarray, stride=2, pad=1, out_pad=1) -> np.ndarray:
    """Direct transposed conv (scatter form), float64, w [Cin,Cout,k,k,k]."""
    x = x.astype(np.float64)
    w = w.astype(np.float64)
    B, Ci, X, Y, Z = x.shape
    _, Co, k, _, _ = w.shape
    full = [(d - 1) * stride + k for d in (X, Y, Z)]
    buf = np.zeros((B, Co, full[0] + out_pad, full[1] + out_pad, full[2] + out_pad))
    for a in range(k):
        for b in range(k):
            for c in range(k):
                buf[:, :, a:a + stride * X:stride, b:b + stride * Y:stride, c:c + stride * Z:stride] += np.einsum(
                    "bixyz,io->boxyz", x, w[:, :, a, b, c])
    ox = [(d - 1) * stride - 2 * pad + k + out_pad for d in (X, Y, Z)]
    return buf[:, :, pad:pad + ox[0], pad:pad + ox[1], pad:pad + ox[2]]
