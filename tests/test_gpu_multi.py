"""Multi-GPU numerical test (needs >= 2 GPUs; skipped otherwise): a 2-rank NCCL data-parallel run of Trainer.train_step
equals two "virtual ranks" run on ONE GPU whose gradients are averaged on the host side.

DDP semantics: BatchNorm statistics and the batch-global ZNCC / HU losses are per rank, only the parameter
gradients are averaged — so the virtual ranks are two Trainers stepping in lockstep (two host threads) whose
`grad_reducer.reduce` exchanges and averages the gradients between the threads.

    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q
"""
import socket
import subprocess
import sys
import threading
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


class _VirtualReducer:
    """Averages .grad across N Trainers running in N threads of this process."""

    def __init__(self, n):
        self.n, self.barrier, self.slots = n, threading.Barrier(n), [None] * n

    def handle(self, r):
        outer = self

        class H:
            def reduce(self, params):
                params = [p for p in params if p.grad is not None]
                torch.cuda.synchronize()
                outer.slots[r] = [p.grad for p in params]
                outer.barrier.wait()
                avg = [sum(outer.slots[i][j] for i in range(outer.n)) / outer.n for j in range(len(params))]
                torch.cuda.synchronize()
                outer.barrier.wait()
                for p, a in zip(params, avg):
                    p.grad = a
                outer.barrier.wait()

        return H()


def _virtual_ranks(dtype, steps, size, world=2):
    import dp_worker as W

    dev = torch.device("cuda", 0)
    red = _VirtualReducer(world)
    trainers = [W.make_trainer(dtype, dev, red.handle(r)) for r in range(world)]
    for t in trainers[1:]:
        t.generator.load_state_dict(trainers[0].generator.state_dict())
        t.critic.load_state_dict(trainers[0].critic.state_dict())
    errs = []

    def run(r):
        try:
            torch.cuda.set_device(0)
            tr = trainers[r]
            tr.generator.train(); tr.critic.train()
            for it in range(steps):
                tr.train_step(W.shard_batches(r, it, (size,) * 3, 2), it)
            torch.cuda.synchronize()
        except Exception as e:  # surface in the main thread
            errs.append(e)
            red.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return trainers


@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_two_rank_nccl_step_equals_averaged_virtual_ranks(tmp_path, dtype_name):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    steps, size = 2, 32
    out = tmp_path / "dp.pt"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "dp_worker.py"), str(out), dtype_name,
                        str(steps), str(size)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = torch.load(out)
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    trainers = _virtual_ranks(dtype, steps, size)
    lr = 2e-4
    for net, mod in (("G", trainers[0].generator), ("D", trainers[0].critic)):
        other = dict((trainers[1].generator if net == "G" else trainers[1].critic).named_parameters())
        for k, p in mod.named_parameters():
            assert torch.equal(p, other[k]), f"virtual ranks diverged on {net}.{k}"
            a, b = got[net][k].float(), p.detach().float().cpu()
            # gradients are summed with floating-point atomics: a weight whose gradient is ~0 may take its first Adam steps
            # (+-lr each) in the other direction; everything else agrees to rounding noise
            # (the critic's BatchNorm shifts are the typical case: their gradient is the difference of two nearly equal sums
            # over the real and the fake batch). So: nothing moves further than Adam can, the typical element agrees to a
            # few % of one Adam step, and at most 2 % of a tensor's elements (at least one) sit more than half a step apart.
            d = (a - b).abs().flatten()
            assert float(d.max()) <= 2 * lr * steps + 1e-7, f"{net}.{k}: {float(d.max())}"
            assert float(d.median()) <= 0.05 * lr, f"{net}.{k}: median |diff| {float(d.median()):.3e}"
            flipped = int((d > 0.5 * lr).sum())
            assert flipped <= max(1, d.numel() // 50), f"{net}.{k}: {flipped} of {d.numel()} elements differ by > lr/2"
    # rank 0's BatchNorm running statistics are those of virtual rank 0 (per-rank statistics)
    for k, v in trainers[0].generator.state_dict().items():
        if "running_" in k:
            assert torch.allclose(got["G"][k], v.cpu(), rtol=1e-3, atol=1e-5), k
