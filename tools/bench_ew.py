"""Micro-benchmark of the HBM-bound BatchNorm kernels: GB/s of algorithmic traffic vs the measured HBM peak."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from contrast_gan_3d_b200 import _lib, ops
from contrast_gan_3d_b200.ops import _p, _st, call

def timeit(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    dev = "cuda"
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    for (rows, C) in ((16 * 128 ** 3, 16), (16 * 64 ** 3, 32), (16 * 32 ** 3, 64)):
        y = torch.randn(rows, C, device=dev).bfloat16()
        dz = torch.randn(rows, C, device=dev).bfloat16()
        z = torch.empty_like(y)
        sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
        mi = torch.rand(2 * C, device=dev) + 0.5
        gamma = torch.rand(C, device=dev) + 0.5
        beta = torch.rand(C, device=dev)
        dg = torch.empty(C, device=dev); db = torch.empty(C, device=dev)
        dt = _lib.BF16
        nbytes = y.numel() * 2
        cases = {
            "bn_stats": (lambda: call("cgan3d_bn_stats", _p(y), dt, rows, C, _p(sums), _st()), 1),
            "bn_apply": (lambda: call("cgan3d_bn_apply", _p(y), _p(z), dt, rows, C, _p(mi), _p(gamma), _p(beta), _lib.ACT_RELU, 0.0, None, _st()), 2),
            "bn_apply+res": (lambda: call("cgan3d_bn_apply", _p(y), _p(z), dt, rows, C, _p(mi), _p(gamma), _p(beta), _lib.ACT_RELU, 0.0, _p(dz), _st()), 3),
            "bn_bwd_reduce": (lambda: call("cgan3d_bn_backward_reduce", _p(dz), _p(y), dt, rows, C, _p(mi), _p(gamma), _p(beta), _lib.ACT_RELU, 0.0, _p(sums), _st()), 2),
            "bn_bwd_apply": (lambda: call("cgan3d_bn_backward_apply", _p(dz), _p(y), _p(z), dt, rows, C, _p(mi), _p(gamma), _p(beta), _lib.ACT_RELU, 0.0, _p(sums), _p(dg), _p(db), 0.0, _st()), 3),
            "torch_copy": (lambda: z.copy_(y), 2),
        }
        for name, (fn, passes) in cases.items():
            ms = timeit(fn)
            gbs = passes * nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": name, "rows": rows, "C": C, "ms": round(ms, 4), "GBps": round(gbs, 1), "frac_hbm_peak": round(gbs / peak, 3)}), flush=True)
main()
