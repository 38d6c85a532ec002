"""CPU-only tests: host logic of the package, the state_dict / init contract against the golden fixtures, and the
C-ABI library (loads, exports every symbol include/cgan3d.h declares; no compute calls without a GPU)."""
import copy
import ctypes
import re
from functools import partial
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _fp(v):
    v = v.detach().double().flatten()
    return np.array([v.sum().item(), v.abs().sum().item(), (v * v).sum().item(), v[0].item(), v[-1].item()])


def test_cabi_library_exports_every_declared_symbol():
    from contrast_gan_3d_b200 import _lib

    header = (ROOT / "include" / "cgan3d.h").read_text()
    declared = set(re.findall(r"\b(cgan3d_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cgan3d.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.cgan3d_version() == 100
    assert lib.cgan3d_capabilities() & 1


def test_cabi_argument_errors_are_reported_without_a_gpu():
    from contrast_gan_3d_b200 import _lib

    lib = _lib.lib()
    g = _lib.ConvGeom(1, 8, 8, 8, 4, 8, 8, 7, 4, 3, 1, 1)  # inconsistent small extent
    rc = lib.cgan3d_conv_gather(ctypes.byref(g), 0, 1, 1, None, 1, None, 0, 1, None)
    assert rc == -2 and b"small extent" in lib.cgan3d_last_error()
    rc = lib.cgan3d_conv_gather(ctypes.byref(g), 7, 1, 1, None, 1, None, 0, 1, None)
    assert rc == -3
    with pytest.raises(_lib.Cgan3dError):
        _lib.call("cgan3d_bn_stats", None, 0, 10, 4, None, None)


def test_modules_match_reference_state_dict_contract_and_seeded_init(golden_dir):
    from contrast_gan_3d_b200.model import PatchGANDiscriminator, ResnetGenerator

    g = np.load(golden_dir / "modules_forward.npz")
    torch.manual_seed(0)
    G = ResnetGenerator(4, 2, 16)
    D = PatchGANDiscriminator(1, 8, 3, negative_slope=0.2)
    assert list(G.state_dict().keys()) == list(g["G_keys"])
    assert list(D.state_dict().keys()) == list(g["D_keys"])
    assert [str(tuple(v.shape)) for v in G.state_dict().values()] == list(g["G_shapes"])
    assert [str(tuple(v.shape)) for v in D.state_dict().values()] == list(g["D_shapes"])
    for k, v in G.state_dict().items():
        np.testing.assert_array_equal(_fp(v), g["G/" + k], err_msg=k)
    for k, v in D.state_dict().items():
        np.testing.assert_array_equal(_fp(v), g["D/" + k], err_msg=k)
    assert D.model.first.activation_fn.negative_slope == 0.2


def test_cpu_tensors_fail_loudly():
    from contrast_gan_3d_b200.model import ResnetGenerator, ZNCCLoss

    G = ResnetGenerator(1, 1, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        G(torch.zeros(1, 1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ZNCCLoss()(torch.zeros(8), torch.zeros(8))


def test_unsupported_variants_say_so():
    from contrast_gan_3d_b200.model import PatchGANDiscriminator, ResnetGenerator
    from torch import nn

    with pytest.raises(NotImplementedError):
        ResnetGenerator(4, 2, 16, is_2D=True)
    with pytest.raises(NotImplementedError):
        PatchGANDiscriminator(1, 8, 3, norm_layer=nn.InstanceNorm3d)
    # the LayerNorm critic of experiments/gp_layernorm.py: per-layer normalised shapes as in reference discriminator.py:41-54
    D = PatchGANDiscriminator(1, 8, 3, norm_layer=nn.LayerNorm, patch_size=(1, 128, 128, 32), elementwise_affine=False)
    assert [tuple(m.normalization.normalized_shape) for m in D.model.middle] == [(16, 32, 32, 8), (32, 16, 16, 4), (64, 8, 8, 2)]
    assert D.twice_differentiable and list(D.state_dict()) == ["model.first.conv.weight", "model.first.conv.bias",
                                                               "model.middle.0.conv.weight", "model.middle.1.conv.weight",
                                                               "model.middle.2.conv.weight", "model.last.weight", "model.last.bias"]


def test_conv_geometry_matches_reference_shape_arithmetic(golden_dir):
    from contrast_gan_3d_b200.model.utils import convolution_output_shape
    from contrast_gan_3d_b200.ops import ConvSpec

    g = np.load(golden_dir / "integer_helpers.npz")
    for row in g["conv_shapes"]:
        dims, (k, p, s, op), want = [int(v) for v in row[:4]], [int(v) for v in row[4:8]], [int(v) for v in row[8:]]
        got = convolution_output_shape(dims, 5, k, p, s, transpose_output_padding=None if op < 0 else op)
        assert got == want
        spec = ConvSpec(transposed=op >= 0, cin=3, cout=5, k=k, stride=s, pad=p, out_pad=max(op, 0))
        _, out = spec.geometry(2, tuple(dims[1:]))
        assert list(out) == want[1:]


def test_scaler_and_sampler_index_law(golden_dir):
    from contrast_gan_3d_b200.data import FactorZeroCenterScaler, pad_amounts, random_crop_lower_bounds
    from oracle import cgan_oracle as O

    g = np.load(golden_dir / "integer_helpers.npz")
    sc = FactorZeroCenterScaler(-1024, 1500, 600)
    assert sc.shift == int(g["scaler_shift"])
    np.testing.assert_array_equal(sc(g["scaler_in"]), g["scaler_out"])
    np.testing.assert_array_equal(sc.unscale(sc(g["scaler_in"])), g["unscale"])
    for shape, patch in (((40, 37, 20), (16, 16, 16)), ((10, 37, 13), (16, 16, 16)), ((16, 16, 16), (16, 16, 16))):
        _, pads = O.pad_nd_image_shape(shape, patch)
        assert pad_amounts(shape, patch) == pads
        padded = [max(a, b) for a, b in zip(shape, patch)]
        np.random.seed(7)
        a = random_crop_lower_bounds(padded, patch)
        np.random.seed(7)
        b = O.random_crop_lbs(padded, patch)
        assert a == b


def test_grid_tiles_match_oracle():
    from contrast_gan_3d_b200.eval import grid_tiles
    from oracle import cgan_oracle as O

    for vol, patch in (((512, 512, 256), (128, 128, 128)), ((20, 16, 33), (16, 16, 16)), ((16, 16, 16), (16, 16, 16))):
        assert grid_tiles(vol, patch) == O.grid_tiles(vol, patch)
    with pytest.raises(ValueError):
        grid_tiles((8, 16, 16), (16, 16, 16))


def test_trainer_checkpoint_roundtrip_and_reference_layout(tmp_path):
    from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator
    from contrast_gan_3d_b200.optim import FusedAdam
    from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer, find_latest_checkpoint

    def make(ckpt_dir):
        torch.manual_seed(0)
        return Trainer(10, 2, None, 1, 1, 1, 0, partial(ResnetGenerator, 1, 1, 4), partial(PatchGANDiscriminator, 1, 4, 1),
                       partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
                       HULoss(0.1, 0.3), NullLogger(), torch.device("cpu"), weight_clip=0.01, checkpoint_dir=ckpt_dir)

    tr = make(tmp_path)
    with torch.no_grad():
        for p in tr.critic.parameters():
            p.add_(1.0)
    tr.save_checkpoint(7)
    ck = torch.load(tmp_path / "7.pt")
    assert list(ck.keys())[:7] == ["iteration", "generator", "optimizer_G", "lr_scheduler_G", "discriminator",
                                   "optimizer_D", "lr_scheduler_D"]
    assert ck["discriminator"] is None and "critic_state_dict" in ck
    assert find_latest_checkpoint(tmp_path).name == "7.pt"
    tr2 = make(tmp_path)  # auto-resume
    assert tr2.iteration == 7
    for a, b in zip(tr.critic.parameters(), tr2.critic.parameters()):
        assert torch.equal(a, b)
    # WGAN-GP mode constructs (weight_clip=None, reference Trainer.py:122-130)
    Trainer(10, 2, None, 1, 1, 1, 0, partial(ResnetGenerator, 1, 1, 4), partial(PatchGANDiscriminator, 1, 4, 1, norm_layer=torch.nn.Identity),
            partial(FusedAdam), partial(FusedAdam), HULoss(0.1, 0.3), NullLogger(), torch.device("cpu"), weight_clip=None)


def test_optimizer_state_cross_loads_with_torch_adam():
    """Checkpoints are exchanged with the reference, whose optimizers are torch.optim.Adam (reference
    experiments/basic_conf.py:55,67, trainer/Trainer.py:311-339): a state_dict of either optimizer must load into the
    other and leave it able to step."""
    from contrast_gan_3d_b200.optim import FusedAdam

    torch.manual_seed(0)
    p_ref = [torch.randn(5, requires_grad=True), torch.randn(3, 2, requires_grad=True)]
    adam = torch.optim.Adam(p_ref, lr=2e-4, betas=(0.5, 0.999))
    for p in p_ref:
        p.grad = torch.randn_like(p)
    adam.step(); adam.step()
    # reference -> ours
    p_my = [p.detach().clone().requires_grad_(True) for p in p_ref]
    fused = FusedAdam(p_my, lr=1e-3)
    fused.load_state_dict(copy.deepcopy(adam.state_dict()))  # load_state_dict aliases same-dtype tensors
    g = fused.param_groups[0]
    assert g["clip"] == 0.0 and g["lr"] == 2e-4 and tuple(g["betas"]) == (0.5, 0.999)
    for p, q in zip(p_my, p_ref):
        st = fused.state[p]
        assert isinstance(st["step"], int) and st["step"] == 2
        assert torch.equal(st["exp_avg"], adam.state[q]["exp_avg"]) and torch.equal(st["exp_avg_sq"], adam.state[q]["exp_avg_sq"])
    # ours -> reference: torch Adam must be able to STEP after loading (it reads weight_decay, amsgrad, maximize, ...)
    adam2 = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in p_ref], lr=1e-3)
    adam2.load_state_dict(copy.deepcopy(fused.state_dict()))
    for p, q in zip(adam2.param_groups[0]["params"], p_ref):
        p.grad = q.grad.clone()
    adam2.step()
    adam.step()
    for p, q in zip(adam2.param_groups[0]["params"], p_ref):
        assert torch.allclose(p, q, rtol=0, atol=0), "torch Adam resumed from a FusedAdam state_dict took a different step"
    with pytest.raises(NotImplementedError):
        sd = adam.state_dict()
        sd["param_groups"][0]["amsgrad"] = True
        fused.load_state_dict(sd)
