"""Debug driver: one thin-conv op through the C ABI (used under compute-sanitizer)."""
import sys, ctypes
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from contrast_gan_3d_b200 import _lib, ops

def main():
    op = sys.argv[1] if len(sys.argv) > 1 else "gather"
    cin, cout = (1, 16) if (len(sys.argv) < 3 or sys.argv[2] == "first") else (16, 1)
    B, sp, pad = 2, (10, 9, 12), 0
    if len(sys.argv) > 3:
        sp = tuple(int(v) for v in sys.argv[3].split("x"))
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((B, cin, *sp), generator=gen).bfloat16().float().requires_grad_(True)
    w = (torch.randn((cout, cin, 7, 7, 7), generator=gen) / (cin * 343) ** 0.5).bfloat16().float().requires_grad_(True)
    y = F.conv3d(x, w, padding=pad)
    gy = torch.randn(y.shape, generator=gen).bfloat16().float()
    gx_ref, gw_ref = torch.autograd.grad(y, (x, w), gy)
    spec = ops.ConvSpec(transposed=False, cin=cin, cout=cout, k=7, stride=1, pad=pad)
    g, _ = spec.geometry(B, sp)
    cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()
    ncl = lambda t: t.permute(0, 4, 1, 2, 3).contiguous()
    xd = cl(x.detach()).to("cuda", torch.bfloat16)
    gyd = cl(gy).to("cuda", torch.bfloat16)
    wp = ops.pack_weights(w.detach().to("cuda"), torch.bfloat16)
    torch.cuda.synchronize()
    print("select", [_lib.lib().cgan3d_conv_select(ctypes.byref(g), _lib.BF16, o) for o in range(3)], flush=True)
    if op == "gather":
        got, ref = ncl(ops.conv_gather(g, xd, wp, impl=_lib.IMPL_TC)).float().cpu(), y.detach()
    elif op == "scatter":
        got, ref = ncl(ops.conv_scatter(g, gyd, wp, impl=_lib.IMPL_TC)).float().cpu(), gx_ref
    else:
        got, ref = ops.conv_wgrad(g, xd, gyd, impl=_lib.IMPL_TC).float().cpu(), gw_ref
    torch.cuda.synchronize()
    err = (got - ref).abs()
    print(op, cin, cout, sp, "max err", err.max().item(), "ref max", ref.abs().max().item(), flush=True)

main()
