"""Golden fixtures of the WGAN-GP training mode, produced by the UNMODIFIED reference Trainer on CPU
(reference trainer/Trainer.py:108-161 with weight_clip=None, model/utils.py:12-41, experiments/gradient_penalty_conf.py,
experiments/gp_layernorm.py).  Run in the authoring container only:

    python tests/golden/make_golden_gp.py

Recipe: model seed torch.manual_seed(0) (G before D); data from torch.Generator().manual_seed(1) as in make_golden.py;
Adam(lr 1e-4, betas (0.0, 0.9); torch 2.11 rejects the int 0 of gradient_penalty_conf.py:10), gp_weight 10 (Trainer default), generator trained every iteration; the interpolation
coefficients of step `it` are the first torch.rand draw after torch.manual_seed(9000 + it), set right before the step
(nothing else in the step consumes the global RNG).
"""
from __future__ import annotations

import sys
from functools import partial
from pathlib import Path

import numpy as np
import torch
from torch import nn

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_shim  # noqa: E402
from oracle import cgan_oracle as O  # noqa: E402
from tests.golden.make_golden import _NullLogger, fingerprint  # noqa: E402


def run(ref, T, patch, norm, steps=3):
    torch.manual_seed(0)
    G = partial(ref.generator.ResnetGenerator, n_resnet_blocks=4, n_updownsample_blocks=2, init_channels_out=16)
    cargs = dict(channels_in=1, init_channels_out=8, discriminator_depth=3, negative_slope=0.2)
    if norm == "identity":
        cargs.update(norm_layer=nn.Identity)
    else:
        cargs.update(norm_layer=nn.LayerNorm, patch_size=(1, *patch), elementwise_affine=False)
    D = partial(ref.discriminator.PatchGANDiscriminator, **cargs)
    scaler = ref.scaler.FactorZeroCenterScaler(-1024, 1500, 600)
    lo, hi = scaler(np.array([350, 450]))
    adam = partial(torch.optim.Adam, lr=1e-4, betas=(0.0, 0.9))
    tr = T.Trainer(10, 2, None, 1, 1, 1, 10 ** 9, G, D, adam, adam, ref.loss.HULoss(float(lo), float(hi), (2, 1, *patch)),
                   _NullLogger(), torch.device("cpu"), weight_clip=None, checkpoint_every=None)
    tr.generator.train(); tr.critic.train()
    cur = {}
    oc, og = tr.train_critic, tr.train_generator
    tr.train_critic = lambda *a, **k: (lambda r: (cur.update({n: float(v.detach()) for n, v in r.items()}), r)[1])(oc(*a, **k))
    tr.train_generator = lambda *a, **k: (lambda r: (cur.update({n: float(v.detach()) for n, v in r.items()}), r)[1])(og(*a, **k))
    gen = torch.Generator().manual_seed(1)
    rows = []
    for it in range(steps):
        opt = O.synthetic_patches(gen, (2, 1, *patch)); low = O.synthetic_patches(gen, (1, 1, *patch)); high = O.synthetic_patches(gen, (1, 1, *patch))
        ml = O.synthetic_masks(gen, (1, 1, *patch)); mh = O.synthetic_masks(gen, (1, 1, *patch))
        patches = [dict(data=opt, seg=torch.zeros_like(opt, dtype=torch.bool), name=["o"] * 2), dict(data=low, seg=ml, name=["l"]),
                   dict(data=high, seg=mh, name=["h"])]
        cur.clear()
        torch.manual_seed(9000 + it)
        tr.train_step(patches, it)
        rows.append([cur[k] for k in ("D", "G", "G-full", "sim", "HU")])
    out = {"losses": np.array(rows, dtype=np.float64), "D_keys": np.array(list(tr.critic.state_dict().keys()))}
    for k, v in fingerprint(tr.generator.state_dict()).items():
        out["G/" + k] = v
    for k, v in fingerprint(tr.critic.state_dict()).items():
        out["D/" + k] = v
    return out


def main():
    ref, T = ref_shim.load(), ref_shim.load_trainer()
    torch.set_num_threads(8)
    np.savez_compressed(HERE / "train_steps_gp_32.npz", **run(ref, T, (32, 32, 32), "identity"))
    np.savez_compressed(HERE / "train_steps_gp_layernorm_32.npz", **run(ref, T, (32, 32, 32), "layer"))
    print("WGAN-GP golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
