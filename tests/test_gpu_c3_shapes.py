"""GPU parity at the BENCHMARK shapes (BASELINE config C3: 16 pairs of 1x128^3 per step).

Every generator and critic layer, all three operators (fprop / dgrad / wgrad), at the batch sizes the bench runs
(B = 16) and at an odd batch (B = 3: unpaired columns of the CTA-pair kernels, ragged split-K), tcgen05 against the
CUDA-core kernels on identical bf16 operands; against ATen (fp32, CPU) at B = 1; a 3-step bf16 train step at 128^3 whose
POST-STEP weights and BatchNorm running statistics are compared with the fp32 CPU oracle.

Reference: model/generator.py:31-87, model/discriminator.py:23-81 (layer shapes), trainer/Trainer.py:108-161 (step).
"""
import ctypes
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cgan_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"

# (name, transposed module?, cin, cout, k, stride, pad, out_pad, spatial_in of the module at 128^3 patches)
C3_LAYERS = [
    ("G.first", False, 1, 16, 7, 1, 0, 0, (134, 134, 134)),  # input already reflection-padded (generator.py:31-38)
    ("G.down0", False, 16, 32, 3, 2, 1, 0, (128, 128, 128)),
    ("G.down1", False, 32, 64, 3, 2, 1, 0, (64, 64, 64)),
    ("G.res", False, 64, 64, 3, 1, 1, 0, (32, 32, 32)),
    ("G.up0", True, 64, 32, 3, 2, 1, 1, (32, 32, 32)),
    ("G.up1", True, 32, 16, 3, 2, 1, 1, (64, 64, 64)),
    ("G.last_conv", False, 16, 1, 7, 1, 0, 0, (134, 134, 134)),
    ("D.first", False, 1, 8, 4, 2, 1, 0, (128, 128, 128)),
    ("D.mid0", False, 8, 16, 4, 2, 1, 0, (64, 64, 64)),
    ("D.mid1", False, 16, 32, 4, 2, 1, 0, (32, 32, 32)),
    ("D.mid2", False, 32, 64, 4, 2, 1, 0, (16, 16, 16)),
]


def _ops():
    from contrast_gan_3d_b200 import _lib, ops

    return _lib, ops


def _one_ulp(got, ref, what):
    """two fp32-accumulating kernels on identical bf16 operands: at most one bf16 ulp apart after the storage rounding"""
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all(), what
    d = (got - ref).abs()
    bad = d > 2 ** -7 * ref.abs() + 1e-3
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} elements differ, max {float(d.max()):.3e} (ref max {float(ref.abs().max()):.3e})"


def _layer_tensors(case, B, seed):
    _lib, ops = _ops()
    name, tr, cin, cout, k, s, p, op, sp = case
    spec = ops.ConvSpec(transposed=tr, cin=cin, cout=cout, k=k, stride=s, pad=p, out_pad=op)
    g, out_sp = spec.geometry(B, sp)
    gen = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((B, *sp, cin), generator=gen, device=DEV).bfloat16()
    gy = torch.randn((B, *out_sp, cout), generator=gen, device=DEV).bfloat16()
    wshape = (cin, cout, k, k, k) if tr else (cout, cin, k, k, k)
    w = (torch.randn(wshape, generator=gen, device=DEV) / (cin * k ** 3) ** 0.5).bfloat16().float()
    return spec, g, x, gy, w


@pytest.mark.parametrize("B", [16, 3], ids=["B16", "B3"])
@pytest.mark.parametrize("case", C3_LAYERS, ids=[c[0] for c in C3_LAYERS])
def test_c3_layer_all_ops_tcgen05_vs_cuda_core(case, B):
    """fprop, dgrad and wgrad of one C3 layer at the bench batch size: tcgen05 vs the CUDA-core kernel."""
    _lib, ops = _ops()
    name, tr = case[0], case[1]
    spec, g, x, gy, w = _layer_tensors(case, B, seed=len(name) + B)
    wp = ops.pack_weights(w, torch.bfloat16)
    fwd_op, bwd_op = (1, 0) if tr else (0, 1)
    sel = [_lib.lib().cgan3d_conv_select(ctypes.byref(g), _lib.BF16, o) for o in range(3)]
    assert sel == [2, 2, 2], f"{name}: every operator of a C3 layer must be on the tcgen05 path, got {sel}"
    fwd = (lambda impl: ops.conv_scatter(g, x, wp, impl=impl)) if tr else (lambda impl: ops.conv_gather(g, x, wp, impl=impl))
    bwd = (lambda impl: ops.conv_gather(g, gy, wp, impl=impl)) if tr else (lambda impl: ops.conv_scatter(g, gy, wp, impl=impl))
    _one_ulp(fwd(_lib.IMPL_TC), fwd(_lib.IMPL_GENERIC), f"{name} B={B} fprop")
    _one_ulp(bwd(_lib.IMPL_TC), bwd(_lib.IMPL_GENERIC), f"{name} B={B} dgrad")
    big, small = (gy, x) if tr else (x, gy)
    dw_tc = ops.conv_wgrad(g, big, small, impl=_lib.IMPL_TC)
    dw_gen = ops.conv_wgrad(g, big, small, impl=_lib.IMPL_GENERIC)
    torch.cuda.synchronize()
    assert torch.isfinite(dw_tc).all()
    # K = B * voxels up to 33.5 M fp32 additions in different orders (split-K + atomics): relative to the largest entry
    scale = float(dw_gen.abs().max())
    d = float((dw_tc - dw_gen).abs().max())
    assert d <= 2e-3 * scale, f"{name} B={B} wgrad: max diff {d:.3e} vs scale {scale:.3e}"
    # fused BatchNorm statistics variant (where the layer offers it) must not change the output
    if ops.conv_fuses_bnstats(g, torch.bfloat16, tr):
        fused, sums = ops.conv_bnstats(g, x, wp, tr)
        plain = fwd(_lib.IMPL_TC)
        assert torch.equal(fused, plain), f"{name} B={B}: conv_bnstats output differs from the plain conv"
        yf = plain.float().reshape(-1, plain.shape[-1]).double()
        ref = torch.cat([yf.sum(0), (yf * yf).sum(0)])
        # sums come from the fp32 accumulators, `ref` from the bf16-rounded output: 2^-9 relative per element, random sign
        assert torch.allclose(sums, ref, rtol=2e-3, atol=2e-3 * float(ref.abs().max())), f"{name} B={B}: fused statistics"


@pytest.mark.parametrize("case", C3_LAYERS, ids=[c[0] for c in C3_LAYERS])
def test_c3_layer_all_ops_vs_aten_fp32(case):
    """The same layers at B = 1 against ATen's fp32 CPU convolution and its autograd (the reference's arithmetic)."""
    _lib, ops = _ops()
    name, tr, cin, cout, k, s, p, op, sp = case
    spec, g, x, gy, w = _layer_tensors(case, 1, seed=len(name))
    wp = ops.pack_weights(w, torch.bfloat16)
    if tr:
        y = ops.conv_scatter(g, x, wp, impl=_lib.IMPL_TC)
        dx = ops.conv_gather(g, gy, wp, impl=_lib.IMPL_TC)
        dw = ops.conv_wgrad(g, gy, x, impl=_lib.IMPL_TC)
    else:
        y = ops.conv_gather(g, x, wp, impl=_lib.IMPL_TC)
        dx = ops.conv_scatter(g, gy, wp, impl=_lib.IMPL_TC)
        dw = ops.conv_wgrad(g, x, gy, impl=_lib.IMPL_TC)
    xr = x.float().cpu().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    wr = w.cpu().clone().requires_grad_(True)
    gyr = gy.float().cpu().permute(0, 4, 1, 2, 3).contiguous()
    yr = F.conv_transpose3d(xr, wr, stride=s, padding=p, output_padding=op) if tr else F.conv3d(xr, wr, stride=s, padding=p)
    dxr, dwr = torch.autograd.grad(yr, (xr, wr), gyr)
    for got, ref, nm in ((y, yr.detach(), "fprop"), (dx, dxr, "dgrad")):
        got = got.float().cpu().permute(0, 4, 1, 2, 3)
        err = (got - ref).abs()
        assert bool((err <= 8e-3 * ref.abs() + 2e-3 * float(ref.abs().max())).all()), f"{name} {nm} vs ATen: {float(err.max()):.3e}"
    err = (dw.cpu() - dwr).abs()
    assert float(err.max()) <= 2e-3 * float(dwr.abs().max()), f"{name} wgrad vs ATen: {float(err.max()):.3e} / {float(dwr.abs().max()):.3e}"


def _make_trainer(dtype):
    from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator
    from contrast_gan_3d_b200.optim import FusedAdam
    from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer

    torch.manual_seed(0)
    return Trainer(10, 2, None, 1, 1, 1, 0,
                   partial(ResnetGenerator, 4, 2, 16, compute_dtype=dtype),
                   partial(PatchGANDiscriminator, 1, 8, 3, negative_slope=0.2, compute_dtype=dtype),
                   partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
                   HULoss(0.18666666666666668, 0.35333333333333333), NullLogger(), torch.device(DEV), weight_clip=0.01,
                   checkpoint_every=None)


KEYS = ("D", "G", "G-full", "sim", "HU")


def test_train_steps_bf16_full_patch_size_post_step_weights():
    """Three full G+D steps at the BASELINE patch size (128^3, 2 pairs), bf16 path, against the fp32 CPU oracle:
    the logged losses of every step (rtol 2e-2), and AFTER the steps the parameter UPDATES of every conv layer (they exist
    only through the dgrad / wgrad chain at the bench shapes), the BatchNorm running statistics and the clipped critic.

    Adam's first updates are lr * m/sqrt(v) ~ lr * sign(g): a weight whose gradient is near zero may move the other way
    under bf16 noise, so the update is compared as a vector: cosine similarity with the oracle's update >= 0.9 and
    >= 90 % of the entries moving in the same direction, per tensor."""
    patch = (128, 128, 128)
    steps = 3
    st = O.StepState(seed=0)
    tr = _make_trainer(torch.bfloat16)
    g0 = {k: v.detach().clone().cpu() for k, v in tr.generator.state_dict().items()}
    d0 = {k: v.detach().clone().cpu() for k, v in tr.critic.state_dict().items()}
    for k in g0:  # identical seeded init (bit-exact) is the precondition of the comparison
        ref0 = {**st.gp, **st.gb}[k]
        assert torch.equal(g0[k], ref0.detach()), k
    gen = torch.Generator().manual_seed(5)
    atol = dict(zip(KEYS, (2e-3, 2e-3, 2e-3, 1e-3, 1e-3)))
    tr.generator.train(); tr.critic.train()
    for it in range(steps):
        opt = O.synthetic_patches(gen, (2, 1, *patch)); low = O.synthetic_patches(gen, (1, 1, *patch)); high = O.synthetic_patches(gen, (1, 1, *patch))
        ml = O.synthetic_masks(gen, (1, 1, *patch)); mh = O.synthetic_masks(gen, (1, 1, *patch))
        ref = O.train_step(st, opt, low, high, ml, mh, it)
        logs = tr.train_step([dict(data=opt, seg=None, name=[]), dict(data=low, seg=ml, name=[]), dict(data=high, seg=mh, name=[])], it)
        for k in KEYS:
            got = float(logs[k].detach())
            assert abs(got - ref[k]) <= 2e-2 * abs(ref[k]) + atol[k], (it, k, got, ref[k])
    torch.cuda.synchronize()
    gs = {k: v.detach().cpu() for k, v in tr.generator.state_dict().items()}
    ds = {k: v.detach().cpu() for k, v in tr.critic.state_dict().items()}
    refs_g, refs_d = {**st.gp, **st.gb}, {**st.dp, **st.db}
    checked = 0
    for k, v in gs.items():
        r = refs_g[k].detach()
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(r) == steps, k
        elif "running_" in k:
            # the stated bf16 criterion: |d| <= 2e-2 |ref| + 2e-2 max|ref|
            assert torch.allclose(v, r, rtol=2e-2, atol=2e-2 * float(r.abs().max()) + 1e-6), (k, float((v - r).abs().max()))
        elif k.endswith("conv.weight") or k.endswith("last_conv.weight"):
            du, dr = (v - g0[k]).flatten().double(), (r - g0[k]).flatten().double()
            cos = float(du @ dr / (du.norm() * dr.norm() + 1e-30))
            same = float(((du * dr) > 0).double().mean())
            assert cos >= 0.9 and same >= 0.9, f"{k}: update cosine {cos:.4f}, same-direction fraction {same:.4f}"
            assert float((v - r).abs().max()) <= 4 * 2e-4 * steps, k  # a sign flip moves a weight by up to 2 lr per step; early Adam steps can exceed lr
            checked += 1
        else:  # BN gamma / beta, last_conv.bias: few entries, compare values
            assert torch.allclose(v, r, rtol=0, atol=2 * 2e-4 * steps + 1e-6), k
    assert checked == 14
    for k, v in ds.items():
        r = refs_d[k].detach()
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(r) == 3 * steps, k  # three critic forward passes per step (Trainer.py:114,116,151)
        elif "running_" in k:
            assert torch.allclose(v, r, rtol=2e-2, atol=2e-2 * float(r.abs().max()) + 1e-6), (k, float((v - r).abs().max()))
        else:
            assert float(v.abs().max()) <= 0.01 + 1e-7, k  # weight clip (Trainer.py:136-138)
            if k.endswith("conv.weight") or k == "model.last.weight":
                # the clip saturates most of BN gamma but only part of the conv weights: compare the entries the oracle
                # left strictly inside the clip range
                inside = r.abs() < 0.0099
                du, dr = (v - d0[k])[inside].double(), (r - d0[k])[inside].double()
                if du.numel() > 100:
                    cos = float(du @ dr / (du.norm() * dr.norm() + 1e-30))
                    assert cos >= 0.9, f"critic {k}: update cosine {cos:.4f}"


def test_validate_at_reference_patch_size_tcgen05_vs_cuda_core():
    """Trainer.validate at the reference's validation patch size (256, 256, 128), batch 2 per scan type (reference
    constants.py:12, trainer/Trainer.py:247-308): eval-mode BatchNorm, non-cubic extents through every generator / critic
    kernel.  The tcgen05 path against the CUDA-core kernels on the same bf16 weights (the CPU oracle needs minutes here)."""
    from contrast_gan_3d_b200 import _lib, ops
    from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator
    from contrast_gan_3d_b200.optim import FusedAdam
    from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer

    torch.manual_seed(0)
    dt = torch.bfloat16
    tr = Trainer(10, 1, None, 1, 1, 1, 0, partial(ResnetGenerator, 4, 2, 16, compute_dtype=dt),
                 partial(PatchGANDiscriminator, 1, 8, 3, negative_slope=0.2, compute_dtype=dt),
                 partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
                 HULoss(0.18666666666666668, 0.35333333333333333), NullLogger(), torch.device(DEV), weight_clip=0.01,
                 checkpoint_every=None)
    # running statistics away from their initial (0, 1) so that the eval-mode normalisation is not the identity
    gen = torch.Generator().manual_seed(5)
    for m in list(tr.generator.modules()) + list(tr.critic.modules()):
        if hasattr(m, "running_mean") and m.running_mean is not None:
            m.running_mean.copy_((torch.rand(m.running_mean.shape, generator=gen) - 0.5) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) * 0.5 + 0.75)
    vpatch = (256, 256, 128)
    batches = [O.synthetic_patches(gen, (2, 1, *vpatch)) for _ in range(3)]

    def run(impl):
        old = ops.set_conv_impl(impl)
        try:
            loaders = {0: iter([dict(data=batches[0])]), -1: iter([dict(data=batches[1])]), 1: iter([dict(data=batches[2])])}
            out = tr.validate(loaders, 400)
            torch.cuda.synchronize()
            return {k: float(v) for k, v in out.items()}
        finally:
            ops.set_conv_impl(old)

    got, ref = run(_lib.IMPL_AUTO), run(_lib.IMPL_GENERIC)
    for k in ("D", "G", "sim"):
        assert np.isfinite(got[k]) and abs(got[k] - ref[k]) <= 2e-2 * abs(ref[k]) + 2e-3, (k, got[k], ref[k])
