"""Generate golden fixtures by running the UNMODIFIED reference (`/root/reference`) on CPU.

Run in the authoring container only (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py

Writes small `.npz` files next to this script.  The recipe:
model seed `torch.manual_seed(0)` with G constructed before D; data from
`torch.Generator().manual_seed(1)`: x = (clamp(300 N(0,1)+100, -1024, 1500) - 238)/600,
masks Bernoulli(1e-3); Adam(2e-4, (0.5, 0.999)), weight_clip=0.01, train_generator_every=1.
"""
from __future__ import annotations

import sys
from functools import partial
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_shim  # noqa: E402
from oracle import cgan_oracle as O  # noqa: E402


def fingerprint(sd):
    out = {}
    for k, v in sd.items():
        v = v.detach().double().flatten()
        out[k] = np.array([v.sum().item(), v.abs().sum().item(), (v * v).sum().item(), v[0].item(), v[-1].item()])
    return out


class _NullLogger:
    class _L:
        @staticmethod
        def log_loss(*a, **k):
            pass

    logger = _L()

    def __call__(self, *a, **k):
        pass

    def end_hook(self):
        pass


def make_trainer(ref, T, patch, n_sub, weight_clip=0.01, gen_every=1):
    torch.manual_seed(0)
    G = partial(ref.generator.ResnetGenerator, n_resnet_blocks=4, n_updownsample_blocks=2, init_channels_out=16)
    D = partial(ref.discriminator.PatchGANDiscriminator, channels_in=1, init_channels_out=8,
                discriminator_depth=3, negative_slope=0.2)
    scaler = ref.scaler.FactorZeroCenterScaler(-1024, 1500, 600)
    lo, hi = scaler(np.array([350, 450]))
    hu = ref.loss.HULoss(float(lo), float(hi), (n_sub, 1, *patch))
    adam = partial(torch.optim.Adam, lr=2e-4, betas=(0.5, 0.999))
    sched = partial(torch.optim.lr_scheduler.MultiStepLR, milestones=[6000, 8000], gamma=0.1)
    tr = T.Trainer(10, 2, None, gen_every, 1, 1, 10 ** 9, G, D, adam, adam, hu, _NullLogger(),
                   torch.device("cpu"), weight_clip=weight_clip, generator_lr_scheduler_class=sched,
                   critic_lr_scheduler_class=sched, checkpoint_every=None)
    tr.generator.train()
    tr.critic.train()
    return tr, (float(lo), float(hi))


def run_steps(ref, T, patch, n_opt, n_low, n_high, n_steps):
    tr, bounds = make_trainer(ref, T, patch, n_low + n_high)
    gen = torch.Generator().manual_seed(1)
    logs = []
    # capture the loss dicts that train_step would log
    orig_c, orig_g = tr.train_critic, tr.train_generator
    cur = {}

    def tc(*a, **k):
        r = orig_c(*a, **k)
        cur.update({k_: float(v.detach()) for k_, v in r.items()})
        return r

    def tg(*a, **k):
        r = orig_g(*a, **k)
        cur.update({k_: float(v.detach()) for k_, v in r.items()})
        return r

    tr.train_critic, tr.train_generator = tc, tg
    for it in range(n_steps):
        opt = O.synthetic_patches(gen, (n_opt, 1, *patch))
        low = O.synthetic_patches(gen, (n_low, 1, *patch))
        high = O.synthetic_patches(gen, (n_high, 1, *patch))
        ml = O.synthetic_masks(gen, (n_low, 1, *patch))
        mh = O.synthetic_masks(gen, (n_high, 1, *patch))
        patches = [dict(data=opt, seg=torch.zeros_like(opt, dtype=torch.bool), name=["o"] * n_opt),
                   dict(data=low, seg=ml, name=["l"] * n_low), dict(data=high, seg=mh, name=["h"] * n_high)]
        cur.clear()
        tr.train_step(patches, it)
        logs.append(dict(cur))
    keys = ["D", "G", "G-full", "sim", "HU"]
    out = {"losses": np.array([[l[k] for k in keys] for l in logs], dtype=np.float64),
           "bounds": np.array(bounds)}
    for k, v in fingerprint(tr.generator.state_dict()).items():
        out["G/" + k] = v
    for k, v in fingerprint(tr.critic.state_dict()).items():
        out["D/" + k] = v
    return out


def main():
    ref = ref_shim.load()
    T = ref_shim.load_trainer()
    torch.set_num_threads(8)

    # ---- 1. module forward goldens (train-mode BN), tiny sizes -----------------------
    torch.manual_seed(0)
    G = ref.generator.ResnetGenerator(4, 2, 16)
    D = ref.discriminator.PatchGANDiscriminator(1, 8, 3, negative_slope=0.2)
    out = {}
    for k, v in fingerprint(G.state_dict()).items():
        out["G/" + k] = v
    for k, v in fingerprint(D.state_dict()).items():
        out["D/" + k] = v
    out["G_keys"] = np.array(list(G.state_dict().keys()))
    out["D_keys"] = np.array(list(D.state_dict().keys()))
    out["G_shapes"] = np.array([str(tuple(v.shape)) for v in G.state_dict().values()])
    out["D_shapes"] = np.array([str(tuple(v.shape)) for v in D.state_dict().values()])
    out["n_params"] = np.array([ref.model_utils.count_parameters(G), ref.model_utils.count_parameters(D)])
    gen = torch.Generator().manual_seed(1)
    xg = O.synthetic_patches(gen, (2, 1, 16, 16, 16))
    xd = O.synthetic_patches(gen, (2, 1, 32, 32, 32))
    G.train(); D.train()
    with torch.no_grad():
        yg = G(xg)
        yd = D(xd)
    out.update(xg=xg.numpy(), yg=yg.numpy(), xd=xd.numpy(), yd=yd.numpy())
    # non-cubic input through G (validation-like aspect ratio) and eval-mode BN
    xr = O.synthetic_patches(gen, (1, 1, 16, 24, 8))
    with torch.no_grad():
        yr = G(xr)
    G.eval()
    with torch.no_grad():
        yg_eval = G(xg)
    out.update(xr=xr.numpy(), yr=yr.numpy(), yg_eval=yg_eval.numpy())
    for k, v in fingerprint(G.state_dict()).items():
        out["G_after/" + k] = v
    np.savez_compressed(HERE / "modules_forward.npz", **out)

    # ---- 2. losses and their gradients -------------------------------------------------
    gen = torch.Generator().manual_seed(2)
    a = O.synthetic_patches(gen, (2, 1, 8, 8, 8)).requires_grad_(True)
    b = O.synthetic_patches(gen, (2, 1, 8, 8, 8))
    m = torch.rand((2, 1, 8, 8, 8), generator=gen) < 0.2
    z = ref.loss.ZNCCLoss()(a, b)
    gz, = torch.autograd.grad(z, a)
    hu_mod = ref.loss.HULoss(0.18666666666666668, 0.35333333333333333, (2, 1, 8, 8, 8))
    h = hu_mod(a, m)
    gh, = torch.autograd.grad(h, a)
    h0 = hu_mod(a, torch.zeros_like(m))
    w = ref.loss.WassersteinLoss()(a.detach().clone(), b)
    w1 = ref.loss.WassersteinLoss()(a.detach().clone())
    np.savez_compressed(HERE / "losses.npz", a=a.detach().numpy(), b=b.numpy(), m=m.numpy(), zncc=z.item(),
                        zncc_grad=gz.numpy(), hu=h.item(), hu_grad=gh.numpy(), hu_empty=h0.item(),
                        wass=w.item(), wass_fake_only=w1.item())

    # ---- 3. train steps through the reference Trainer -----------------------------------
    st = run_steps(ref, T, (32, 32, 32), 2, 1, 1, 3)
    np.savez_compressed(HERE / "train_steps_32.npz", **st)
    st = run_steps(ref, T, (64, 64, 64), 2, 1, 1, 2)  # BASELINE config C1
    np.savez_compressed(HERE / "train_steps_c1_64.npz", **st)
    print("C1 losses [D,G,G-full,sim,HU]:\n", st["losses"])

    # ---- 4. integer helpers ----------------------------------------------------------------
    rows = []
    for dims in ([1, 128, 128, 128], [1, 64, 64, 64], [1, 256, 256, 128], [1, 33, 47, 19]):
        for (k, p, s, op) in ((7, 3, 1, None), (3, 1, 2, None), (4, 1, 2, None), (4, 1, 1, None), (3, 1, 2, 1)):
            r = ref.model_utils.convolution_output_shape(dims, 5, k, p, s, transpose_output_padding=op)
            rows.append(dims + [k, p, s, -1 if op is None else op] + r)
    sc = ref.scaler.FactorZeroCenterScaler(-1024, 1500, 600)
    v = np.array([-1024, -3, 0, 238, 350, 450, 1500], dtype=np.float32)
    g_shape = ref.model_utils.compute_convolution_filters_shape(G, (1, 128, 128, 128), show=False)
    d_shape = ref.model_utils.compute_convolution_filters_shape(D, (1, 128, 128, 128), show=False)
    np.savez_compressed(HERE / "integer_helpers.npz", conv_shapes=np.array(rows), scaler_in=v, scaler_out=sc(v),
                        scaler_shift=sc.shift, unscale=sc.unscale(sc(v)), g_out_shape=np.array(g_shape),
                        d_out_shape=np.array(d_shape),
                        scan_type_order=np.array([s.value for s in ref.alias.ScanType]))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
