"""world_size-2 gloo tests (CPU) for the data-parallel host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from contrast_gan_3d_b200.model import PatchGANDiscriminator
    from contrast_gan_3d_b200.parallel import GradBucketReducer, broadcast_module

    torch.manual_seed(100 + rank)  # different init per rank on purpose
    D = PatchGANDiscriminator(1, 4, 2)
    broadcast_module(D)
    torch.manual_seed(7 + rank)
    for p in D.parameters():
        p.grad = torch.randn_like(p)
    local = [p.grad.clone() for p in D.parameters()]
    GradBucketReducer(bucket_bytes=4096).reduce(D.parameters())  # small buckets: several all-reduces
    reduced = [p.grad.numpy().copy() for p in D.parameters()]
    # overlapped path: gradients are views into the buckets, all-reduce launched from autograd hooks during backward
    red = GradBucketReducer(bucket_bytes=4096)
    for p in D.parameters():
        p.grad = None
    red.prepare(D.parameters())
    loss = sum((p * l).sum() for p, l in zip(D.parameters(), local))  # d loss / d p == the same local gradients
    loss.backward()
    red.finish(D.parameters())
    for p, r in zip(D.parameters(), reduced):
        assert torch.allclose(p.grad, torch.from_numpy(r), rtol=1e-6, atol=1e-7), "overlapped reducer != pack/unpack reducer"
    q.put((rank, [p.detach().numpy().copy() for p in D.parameters()], [t.numpy().copy() for t in local], reduced))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_average_and_broadcast_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0, l0, g0), (_, w1, l1, g1) = res
    import numpy as np

    for a, b in zip(w0, w1):
        np.testing.assert_array_equal(a, b)  # broadcast made the replicas identical
    for a, b, x, y in zip(g0, g1, l0, l1):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_allclose(a, (x + y) / 2, rtol=1e-6, atol=1e-7)
