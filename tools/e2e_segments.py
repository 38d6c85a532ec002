"""Where does the end-to-end (host batches) step spend more device time than the resident step?
Times the three phases of Trainer.train_step with CUDA events for resident and for pinned-host batches.

    python tools/e2e_segments.py [--pairs 16] [--patch 128]
"""
import argparse
import sys
from functools import partial
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator  # noqa: E402
from contrast_gan_3d_b200.optim import FusedAdam  # noqa: E402
from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    tr = Trainer(10 ** 9, 2, None, 1, 1, 0, 0, partial(ResnetGenerator, 4, 2, 16, compute_dtype=dt),
                 partial(PatchGANDiscriminator, 1, 8, 3, negative_slope=0.2, compute_dtype=dt),
                 partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
                 HULoss(*bench.HU_BOUNDS), NullLogger(), dev, weight_clip=0.01, checkpoint_every=None)
    tr.generator.train(); tr.critic.train()
    gen = torch.Generator().manual_seed(1)
    host = bench.synth_batch(gen, args.pairs, args.pairs // 2, args.pairs // 2, (args.patch,) * 3, pin=True)
    resident = [dict(data=b["data"].to(dev), seg=None if b["seg"] is None else b["seg"].to(dev), name=[]) for b in host]

    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    def wrap(obj, attr, name):
        fn = getattr(obj, attr)

        def w(*a, **k):
            mark(name + ":begin")
            r = fn(*a, **k)
            mark(name + ":end")
            return r
        setattr(obj, attr, w)

    wrap(tr, "_generate", "G.forward")
    wrap(tr, "train_critic", "critic step")
    wrap(tr, "train_generator", "generator step")
    for mode, batches in (("resident", resident), ("host", host), ("host+prefetch", host)):
        for _ in range(4):
            tr.train_step(batches, 0)
        torch.cuda.synchronize()
        marks.clear()
        for it in range(args.steps):
            mark("step:begin")
            tr.train_step(batches, 0)
            if mode == "host+prefetch" and it + 1 < args.steps:
                tr.prefetch(batches)
            mark("step:end")
        torch.cuda.synchronize()
        acc = {}
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            acc.setdefault(f"{n0} -> {n1}", []).append(e0.elapsed_time(e1))
        print(f"== {mode}")
        for k, v in acc.items():
            print(f"   {k:45s} {sum(v) / len(v):8.3f} ms  (n={len(v)})")


if __name__ == "__main__":
    main()
