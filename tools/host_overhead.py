"""Host-side cost of one train_step (Python + ctypes + allocator), measured without waiting for the GPU."""
import sys, time
from functools import partial
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import synth_batch, HU_BOUNDS
from contrast_gan_3d_b200 import _lib
from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator
from contrast_gan_3d_b200.optim import FusedAdam
from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer

dev = torch.device("cuda:0")
patch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0)
tr = Trainer(10 ** 9, 2, None, 1, 1, 0, 0, partial(ResnetGenerator, 4, 2, 16, compute_dtype=torch.bfloat16),
             partial(PatchGANDiscriminator, 1, 8, 3, negative_slope=0.2, compute_dtype=torch.bfloat16),
             partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
             HULoss(*HU_BOUNDS), NullLogger(), dev, weight_clip=0.01, checkpoint_every=None)
gen = torch.Generator().manual_seed(1)
host = synth_batch(gen, n, n // 2, n // 2, (patch,) * 3, pin=True)
res = [dict(data=b["data"].to(dev), seg=None if b["seg"] is None else b["seg"].to(dev), name=[]) for b in host]
for _ in range(3):
    tr.train_step(res, 0)
torch.cuda.synchronize()
n0 = _lib.launch_count
t0 = time.perf_counter()
for _ in range(5):
    tr.train_step(res, 0)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"patch {patch} pairs {n}: host {1e3 * (t1 - t0) / 5:.2f} ms/step, wall incl. GPU {1e3 * (t2 - t0) / 5:.2f} ms/step, "
      f"{(_lib.launch_count - n0) / 5:.0f} lib calls/step")


def timed(batches, read_loss, steps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        logs = tr.train_step(batches, 0)
        if read_loss:
            float(logs["G-full"].detach())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name, b, rl in (("resident, no sync", res, False), ("resident, loss read each step", res, True),
                    ("pinned host, no sync", host, False), ("pinned host, loss read each step (= e2e)", host, True)):
    print(f"{name}: {timed(b, rl):.2f} ms/step")

# diagnostics for the un-synchronised host-fed loop
import types
orig = Trainer._side_upload
def main_stream_upload(self, key, parts):
    if all(t.device == self.device for t in parts):
        return (parts[0] if len(parts) == 1 else torch.cat(parts)), None
    out = torch.empty((sum(t.shape[0] for t in parts), *parts[0].shape[1:]), dtype=parts[0].dtype, device=self.device)
    o = 0
    for t in parts:
        out[o:o + t.shape[0]].copy_(t, non_blocking=True); o += t.shape[0]
    return out, None
tr._side_upload = types.MethodType(main_stream_upload, tr)
print(f"[diag] all uploads on the compute stream, no sync: {timed(host, False):.2f} ms/step; with sync: {timed(host, True):.2f}")
tr._side_upload = types.MethodType(orig, tr)
t0 = time.perf_counter(); ms = timed(host, False, steps=10); t1 = time.perf_counter()
print(f"[diag] side stream, no sync, 10 steps: {ms:.2f} ms/step (wall {1e3 * (t1 - t0) / 10:.2f})")
