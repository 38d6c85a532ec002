"""Worker of tests/test_gpu_multi.py: one data-parallel rank (launched by torch.distributed.run, one process per GPU).

Runs `steps` full G+D train steps on this rank's shard through Trainer + GradBucketReducer over NCCL and lets rank 0
save the post-step weights.  Every rank also checks that its replica equals rank 0's afterwards (bit for bit: the
averaged gradients are the same tensors on every rank)."""
import os
import sys
from functools import partial
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def shard_batches(rank: int, step: int, patch, n_pairs: int):
    from oracle import cgan_oracle as O  # data law only

    gen = torch.Generator().manual_seed(1000 + 17 * rank + step)
    opt = O.synthetic_patches(gen, (n_pairs, 1, *patch))
    low = O.synthetic_patches(gen, (n_pairs // 2, 1, *patch))
    high = O.synthetic_patches(gen, (n_pairs - n_pairs // 2, 1, *patch))
    ml = O.synthetic_masks(gen, (n_pairs // 2, 1, *patch), p=0.01)
    mh = O.synthetic_masks(gen, (n_pairs - n_pairs // 2, 1, *patch), p=0.01)
    return [dict(data=opt, seg=None, name=[]), dict(data=low, seg=ml, name=[]), dict(data=high, seg=mh, name=[])]


def make_trainer(dtype, dev, reducer):
    from contrast_gan_3d_b200.model import HULoss, PatchGANDiscriminator, ResnetGenerator
    from contrast_gan_3d_b200.optim import FusedAdam
    from contrast_gan_3d_b200.trainer.Trainer import NullLogger, Trainer

    torch.manual_seed(0)
    return Trainer(10, 2, None, 1, 1, 0, 0, partial(ResnetGenerator, 4, 2, 16, compute_dtype=dtype),
                   partial(PatchGANDiscriminator, 1, 8, 3, negative_slope=0.2, compute_dtype=dtype),
                   partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)), partial(FusedAdam, lr=2e-4, betas=(0.5, 0.999)),
                   HULoss(0.18666666666666668, 0.35333333333333333), NullLogger(), dev, weight_clip=0.01,
                   checkpoint_every=None, grad_reducer=reducer)


def main():
    out, dtype_name, steps, size = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from contrast_gan_3d_b200.parallel import GradBucketReducer, broadcast_module

    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    tr = make_trainer(dtype, dev, GradBucketReducer())
    broadcast_module(tr.generator); broadcast_module(tr.critic)
    tr.generator.train(); tr.critic.train()
    losses = []
    for it in range(steps):
        logs = tr.train_step(shard_batches(rank, it, (size,) * 3, 2), it)
        losses.append({k: float(v.detach()) for k, v in logs.items()})
    torch.cuda.synchronize()
    for mod in (tr.generator, tr.critic):  # replicas stay identical (BatchNorm running stats are per rank by design)
        for k, p in mod.named_parameters():
            ref = p.detach().clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(ref, p.detach()), f"rank {rank}: parameter {k} diverged from rank 0"
    if rank == 0:
        torch.save({"G": {k: v.cpu() for k, v in tr.generator.state_dict().items()},
                    "D": {k: v.cpu() for k, v in tr.critic.state_dict().items()}, "losses": losses, "world": world}, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
