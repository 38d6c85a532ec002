"""Does a concurrent pinned host->device copy slow kernels down?  (explains end-to-end vs resident step time)

    python tools/h2d_interference.py
Times an HBM-bound copy kernel and the tcgen05 ResNet convolution alone and under a back-to-back H2D stream.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from contrast_gan_3d_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    host = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty_like(host, device=dev)
    side = torch.cuda.Stream()
    a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    spec = ops.ConvSpec(transposed=False, cin=64, cout=64, k=3, stride=1, pad=1, out_pad=0)
    g, _ = spec.geometry(16, (32, 32, 32))
    x = torch.randn((16, 32, 32, 32, 64), device=dev).bfloat16()
    wp = ops.pack_weights(torch.randn((64, 64, 3, 3, 3), device=dev) * 0.02, torch.bfloat16)

    def run(fn, iters, with_h2d):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if with_h2d:
            with torch.cuda.stream(side):
                for _ in range(40):
                    dst.copy_(host, non_blocking=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for name, fn, iters in (("hbm copy 1 GiB", lambda: b.copy_(a), 100), ("resnet conv tcgen05", lambda: ops.conv_gather(g, x, wp), 300)):
        alone = run(fn, iters, False)
        busy = run(fn, iters, True)
        print(f"{name}: alone {alone:.4f} ms, under H2D {busy:.4f} ms ({100 * (busy / alone - 1):+.1f} %)")


if __name__ == "__main__":
    main()
