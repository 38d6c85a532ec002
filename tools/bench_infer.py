"""BASELINE config C2: generator-only whole-volume inference, synthetic 512x512x256 int16 CCTA volume tiled into 128^3
patches (32 tiles, batches of 16) on one B200, through CCTAContrastCorrector.__call__ (host int16 in, host fp32 out).

    python tools/bench_infer.py [--dtype bf16|f32] [--iters 3] [--cpu]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from functools import partial
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

G_FPROP_GFLOP_128 = 125.762


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--shape", default="512,512,256")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on 2 tiles and scale")
    args = ap.parse_args()
    from contrast_gan_3d_b200.data import FactorZeroCenterScaler
    from contrast_gan_3d_b200.eval import CCTAContrastCorrector
    from contrast_gan_3d_b200.model import ResnetGenerator

    shape = tuple(int(v) for v in args.shape.split(","))
    rng = np.random.default_rng(0)
    vol = np.clip(rng.normal(100, 300, size=shape), -1024, 1500).astype(np.int16)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    corr = CCTAContrastCorrector(partial(ResnetGenerator, 4, 2, 16, compute_dtype=dt), FactorZeroCenterScaler(-1024, 1500, 600),
                                 torch.device("cuda:0"), inference_patch_size=(128, 128, 128))
    for _ in range(2):
        out = corr(vol, batch_size=16)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        t0 = time.perf_counter()
        out = corr(vol, batch_size=16)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    ntiles = (shape[0] // 128) * (shape[1] // 128) * (shape[2] // 128)
    res = {"config": "C2 generator-only inference", "volume": list(shape), "tiles": ntiles, "dtype": args.dtype,
           "seconds_per_volume_e2e": t, "tiles_per_sec": ntiles / t, "g_fprop_tflops": ntiles * G_FPROP_GFLOP_128 / t / 1e3,
           "out_shape": list(out.shape), "finite": bool(torch.isfinite(out).all())}
    if args.cpu:
        from oracle import cgan_oracle as O
        torch.manual_seed(0)
        gp, gb = O.init_params(O.generator_layers())
        x = torch.randn(2, 1, 128, 128, 128)
        with torch.no_grad():
            O.generator_forward(gp, gb, x)
            t0 = time.perf_counter()
            O.generator_forward(gp, gb, x)
            tc = time.perf_counter() - t0
        res["cpu_oracle_seconds_per_volume_scaled"] = tc / 2 * ntiles
        res["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
