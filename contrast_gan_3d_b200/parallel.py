"""Data-parallel plumbing: one process per GPU, replicas of G and D, gradients averaged over ranks.

The reference has no parallelism at all; this is the batch-sharded path of BASELINE config C4.
Semantics ("DDP semantics"): BatchNorm statistics and the batch-global ZNCC / HU losses are per
rank; only parameter gradients are exchanged, as flat fp32 buckets (G 4.14 MB, D 0.71 MB: latency-bound on
NVLink, so few large buckets rather than many small ones)."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, process_group=None, bucket_bytes: int = 8 << 20):
        self.pg = process_group
        self.bucket_bytes = bucket_bytes
        self._flat = {}

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.pg) if dist.is_initialized() else 1

    def _buckets(self, params: List[torch.nn.Parameter]):
        cur, size, out = [], 0, []
        for p in reversed(params):  # reverse registration order == order in which backward produces them
            n = p.numel() * 4
            if cur and size + n > self.bucket_bytes:
                out.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            out.append(cur)
        return out

    def reduce(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Average `.grad` of every parameter over all ranks (in place)."""
        params = [p for p in params if p.grad is not None]
        if not params or self.world_size == 1:
            return
        works = []
        for bi, bucket in enumerate(self._buckets(params)):
            total = sum(p.numel() for p in bucket)
            key = (bi, total, bucket[0].device)
            flat = self._flat.get(key)
            if flat is None:
                flat = torch.empty(total, dtype=torch.float32, device=bucket[0].device)
                self._flat[key] = flat
            off = 0
            for p in bucket:
                n = p.numel()
                flat[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), flat, bucket))
        inv = 1.0 / self.world_size
        for work, flat, bucket in works:
            work.wait()
            off = 0
            for p in bucket:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                p.grad.mul_(inv)
                off += n


def broadcast_module(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
