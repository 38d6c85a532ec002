"""Robustness sweep: one bf16 G+D train step (and a validate pass) at several patch shapes; every conv must stay on tcgen05."""
import sys, time
from functools import partial
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import HU_BOUNDS
from contrast_gan_3d_b200 import ops
from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator
from contrast_gan_3d_b200.optim import FusedAdam
from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer
from oracle import cgan_oracle as O

dev = torch.device("cuda:0")
shapes = [(64, 64, 64), (128, 128, 32), (256, 256, 128), (96, 80, 48), (128, 128, 128)]
for patch in shapes:
    torch.manual_seed(0)
    tr = Trainer(10 ** 9, 1, None, 1, 1, 0, 0, partial(ResnetGenerator, 4, 2, 16, compute_dtype=torch.bfloat16),
                 partial(PatchGANDiscriminator, 1, 8, 3, negative_slope=0.2, compute_dtype=torch.bfloat16),
                 partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
                 HULoss(*HU_BOUNDS), NullLogger(), dev, weight_clip=0.01, checkpoint_every=None)
    gen = torch.Generator().manual_seed(1)
    n = 2
    b = [dict(data=O.synthetic_patches(gen, (n, 1, *patch)).to(dev), seg=None, name=[]),
         dict(data=O.synthetic_patches(gen, (n // 2, 1, *patch)).to(dev), seg=O.synthetic_masks(gen, (n // 2, 1, *patch)).to(dev), name=[]),
         dict(data=O.synthetic_patches(gen, (n // 2, 1, *patch)).to(dev), seg=O.synthetic_masks(gen, (n // 2, 1, *patch)).to(dev), name=[])]
    ops.enable_conv_timing(True)
    logs = tr.train_step(b, 0)
    torch.cuda.synchronize()
    impls = {k[0] + str(list(k[2:])): v[3] for k, v in ops.conv_timing_summary().items()}
    ops.enable_conv_timing(False)
    generic = [k for k, v in impls.items() if v != 2 and "8, 1, 4, 1, 1]" not in k and ", 1, 4, 1, 1]" not in k]
    t0 = time.perf_counter()
    for _ in range(3):
        logs = tr.train_step(b, 0)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    loaders = {0: iter([b[0]]), -1: iter([b[1]]), 1: iter([b[2]])}
    val = tr.validate(loaders, 400)
    ok = all(torch.isfinite(v).all() for v in logs.values()) and all(torch.isfinite(v).all() for v in val.values())
    print(f"patch {patch}: {ms:.2f} ms/step for {n} pairs, finite={ok}, convs not on tcgen05: {generic}", flush=True)
