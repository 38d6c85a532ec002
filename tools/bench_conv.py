"""Micro-benchmark of the convolution kernels (BASELINE config C5 style): per-op device time with CUDA events.

    python tools/bench_conv.py [--cases res,down,...] [--impls tc,generic,cudnn,cpu] [--iters 20]
    python tools/bench_conv.py --sweep --impls tc,cudnn,cpu --ops gather,scatter,wgrad     # BASELINE config C5
Prints one JSON line per (case, op, impl): TFLOP/s of algorithmic 2*MAC work and the fraction of the measured bf16 peak.
`cudnn` is the comparison line of SURVEY §8d: aten::convolution / convolution_backward in bf16 channels_last_3d with
cudnn.benchmark = True on the same shapes (library code, never on the product path; non-transposed cases only).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from contrast_gan_3d_b200 import _lib, ops  # noqa: E402

CASES = {
    # name: (transposed, cin, cout, k, stride, pad, out_pad, B, spatial_in)
    "res": (False, 64, 64, 3, 1, 1, 0, 16, (32, 32, 32)),
    "res_b2": (False, 64, 64, 3, 1, 1, 0, 2, (32, 32, 32)),
    "c32_64": (False, 32, 32, 3, 1, 1, 0, 4, (64, 64, 64)),
    "c16_128": (False, 16, 16, 3, 1, 1, 0, 2, (128, 128, 128)),
    "c128_32": (False, 128, 128, 3, 1, 1, 0, 8, (32, 32, 32)),
    "down0": (False, 16, 32, 3, 2, 1, 0, 4, (128, 128, 128)),
    "down0_c3": (False, 16, 32, 3, 2, 1, 0, 16, (128, 128, 128)),
    "down1": (False, 32, 64, 3, 2, 1, 0, 8, (64, 64, 64)),
    "up0": (True, 64, 32, 3, 2, 1, 1, 8, (32, 32, 32)),
    "up1": (True, 32, 16, 3, 2, 1, 1, 4, (64, 64, 64)),
    "up1_c3": (True, 32, 16, 3, 2, 1, 1, 16, (64, 64, 64)),
    "first_c3": (False, 1, 16, 7, 1, 0, 0, 16, (134, 134, 134)),
    "last_c3": (False, 16, 1, 7, 1, 0, 0, 16, (134, 134, 134)),
    "first": (False, 1, 16, 7, 1, 0, 0, 2, (134, 134, 134)),
    "last": (False, 16, 1, 7, 1, 0, 0, 2, (134, 134, 134)),
    "d_first": (False, 1, 8, 4, 2, 1, 0, 16, (128, 128, 128)),
    "d_mid0": (False, 8, 16, 4, 2, 1, 0, 16, (64, 64, 64)),
    "d_last": (False, 64, 1, 4, 1, 1, 0, 16, (8, 8, 8)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="res")
    ap.add_argument("--impls", default="tc,generic")
    ap.add_argument("--ops", default="gather,scatter,wgrad")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--sweep", action="store_true", help="BASELINE config C5: channels 16-256 x spatial 32^3-128^3 x stride 1/2, k = 3")
    ap.add_argument("--cpu-gflop-cap", type=float, default=120.0, help="`cpu` line only for cases up to this many GFLOP per op (B = 1)")
    args = ap.parse_args()
    if args.sweep:
        # x [B, C, S, S, S], C in {16..256}, S in {32, 64, 128}, k = 3, stride in {1, 2}, C -> C channels, B = the largest
        # power of two <= 16 that keeps the bf16 tensor <= 2 GiB
        names = []
        for C in (16, 32, 64, 128, 256):
            for S in (32, 64, 128):
                for st in (1, 2):
                    B = 16
                    while B > 1 and B * C * S ** 3 * 2 > 2 << 30:
                        B //= 2
                    name = f"c5_C{C}_S{S}_s{st}"
                    CASES[name] = (False, C, C, 3, st, 1, 0, B, (S, S, S))
                    names.append(name)
        args.cases = ",".join(names)
    peak = 1396.9
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = json.loads(pk.read_text())["bf16_tflops"]
    dev = torch.device("cuda:0")
    for name in args.cases.split(","):
        tr, cin, cout, k, s, p, op_, B, sp = CASES[name]
        spec = ops.ConvSpec(transposed=tr, cin=cin, cout=cout, k=k, stride=s, pad=p, out_pad=op_)
        g, out_sp = spec.geometry(B, sp)
        big = torch.randn((g.B, g.Xb, g.Yb, g.Zb, g.Cb), device=dev).bfloat16()
        small = torch.randn((g.B, g.Xs, g.Ys, g.Zs, g.Cs), device=dev).bfloat16()
        w = torch.randn((g.Cs, g.Cb, k, k, k), device=dev) / (g.Cb * k ** 3) ** 0.5
        wp = ops.pack_weights(w, torch.bfloat16)
        flops = 2.0 * g.B * g.Xs * g.Ys * g.Zs * g.Cs * g.Cb * k ** 3
        for opn in args.ops.split(","):
            opi = {"gather": 0, "scatter": 1, "wgrad": 2}[opn]
            for impl in args.impls.split(","):
                if impl == "cudnn":
                    if tr:
                        continue
                    torch.backends.cudnn.benchmark = True
                    xc = big.permute(0, 4, 1, 2, 3)   # NCDHW view of the channels-last tensor (channels_last_3d strides)
                    yc = small.permute(0, 4, 1, 2, 3)
                    wc = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
                    st3, pd3 = [s, s, s], [p, p, p]
                    bwd = torch.ops.aten.convolution_backward
                    fn = {"gather": lambda: torch.ops.aten.convolution(xc, wc, None, st3, pd3, [1, 1, 1], False, [0, 0, 0], 1),
                          "scatter": lambda: bwd(yc, xc, wc, None, st3, pd3, [1, 1, 1], False, [0, 0, 0], 1, [True, False, False]),
                          "wgrad": lambda: bwd(yc, xc, wc, None, st3, pd3, [1, 1, 1], False, [0, 0, 0], 1, [False, True, False])}[opn]
                    try:
                        for _ in range(3):
                            fn()
                    except Exception as e:  # cuDNN has no engine for some thin shapes
                        print(json.dumps({"case": name, "op": opn, "impl": impl, "error": str(e)[:120]}), flush=True)
                        continue
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.iters):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.iters
                    tf = flops / (ms * 1e-3) / 1e12
                    print(json.dumps({"case": name, "op": opn, "impl": impl, "ms": round(ms, 4), "tflops": round(tf, 2),
                                      "frac_of_bf16_burst_peak": round(tf / peak, 4), "gflop": round(flops / 1e9, 2)}), flush=True)
                    continue
                if impl == "cpu":
                    # the reference's path: aten::convolution / convolution_backward on the host cores, fp32, B = 1
                    f1 = flops / g.B
                    if f1 / 1e9 > args.cpu_gflop_cap:
                        print(json.dumps({"case": name, "op": opn, "impl": impl, "skipped": f"{f1 / 1e9:.0f} GFLOP at B = 1 exceeds --cpu-gflop-cap"}), flush=True)
                        continue
                    import time
                    xc = big[:1].float().cpu().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
                    wc = w.float().cpu().requires_grad_(True)
                    if tr:
                        continue
                    yc = torch.nn.functional.conv3d(xc, wc, stride=s, padding=p)
                    gyc = torch.randn_like(yc)
                    fn = {"gather": lambda: torch.nn.functional.conv3d(xc, wc, stride=s, padding=p),
                          "scatter": lambda: torch.autograd.grad(yc, xc, gyc, retain_graph=True),
                          "wgrad": lambda: torch.autograd.grad(yc, wc, gyc, retain_graph=True)}[opn]
                    fn()
                    t0 = time.perf_counter()
                    for _ in range(2):
                        fn()
                    ms = (time.perf_counter() - t0) / 2 * 1e3
                    print(json.dumps({"case": name, "op": opn, "impl": impl, "ms": round(ms, 2), "tflops": round(f1 / (ms * 1e-3) / 1e12, 4),
                                      "gflop": round(f1 / 1e9, 2), "batch": 1, "threads": torch.get_num_threads()}), flush=True)
                    continue
                ii = {"tc": _lib.IMPL_TC, "generic": _lib.IMPL_GENERIC}[impl]
                if impl == "tc" and _lib.lib().cgan3d_conv_select(ctypes.byref(g), _lib.BF16, opi) != 2:
                    print(json.dumps({"case": name, "op": opn, "impl": impl, "unsupported": "no tcgen05 plan for this shape (CUDA-core kernel would run)"}), flush=True)
                    continue
                fn = {"gather": lambda: ops.conv_gather(g, big, wp, impl=ii), "scatter": lambda: ops.conv_scatter(g, small, wp, impl=ii),
                      "wgrad": lambda: ops.conv_wgrad(g, big, small, impl=ii)}[opn]
                iters = args.iters if impl == "tc" else max(2, args.iters // 5)
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                tf = flops / (ms * 1e-3) / 1e12
                print(json.dumps({"case": name, "op": opn, "impl": impl, "ms": round(ms, 4), "tflops": round(tf, 2),
                                  "frac_of_bf16_burst_peak": round(tf / peak, 4), "gflop": round(flops / 1e9, 2), "batch": g.B}), flush=True)


if __name__ == "__main__":
    main()
