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
    """Two ways to use it:

    * `reduce(params)` after backward: pack -> all-reduce -> unpack (simple, used by the CPU tests);
    * `prepare(params)` before backward + `finish(params)` after it (what `Trainer` does): the gradients ARE views
      into flat fp32 buckets (no pack / unpack kernels), every bucket is all-reduced asynchronously from an autograd
      post-accumulate hook as soon as its last gradient has been produced, i.e. overlapped with the rest of backward
      (buckets follow reverse registration order = the order in which backward produces them), and `finish` only waits.
    """

    def __init__(self, process_group=None, bucket_bytes: int = 8 << 20):
        self.pg = process_group
        self.bucket_bytes = bucket_bytes
        self._flat = {}
        self._plans = {}   # id(first param) -> bucket plan of a parameter set
        self._hooked = set()

    # ------------------------------------------------------------------ overlapped path
    def _plan(self, params: List[torch.nn.Parameter]):
        key = tuple(id(p) for p in params)
        plan = self._plans.get(key)
        if plan is None:
            buckets = []
            for bucket in self._buckets(params):
                total = sum(p.numel() for p in bucket)
                flat = torch.zeros(total, dtype=torch.float32, device=bucket[0].device)
                buckets.append(dict(params=bucket, flat=flat, ready=0, work=None))
            index = {}
            for bi, b in enumerate(buckets):
                for p in b["params"]:
                    index[id(p)] = bi
            plan = dict(buckets=buckets, index=index, active=False)
            self._plans[key] = plan
            for p in params:
                if id(p) not in self._hooked:
                    self._hooked.add(id(p))
                    p.register_post_accumulate_grad_hook(self._make_hook(plan))
        return plan

    def _make_hook(self, plan):
        def hook(param):
            if not plan["active"]:
                return
            b = plan["buckets"][plan["index"][id(param)]]
            b["ready"] += 1
            if b["ready"] == len(b["params"]) and b["work"] is None:
                b["work"] = self._all_reduce_avg(b["flat"])
        return hook

    def _all_reduce_avg(self, flat: torch.Tensor):
        if flat.is_cuda:  # NCCL averages inside the collective
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True), None
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), 1.0 / self.world_size

    def prepare(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Call after zero_grad and before backward: gradients become zeroed views into the flat buckets."""
        params = [p for p in params if p.requires_grad]
        if not params or self.world_size == 1:
            return
        plan = self._plan(params)
        for b in plan["buckets"]:
            b["flat"].zero_()
            b["ready"], b["work"] = 0, None
            off = 0
            for p in b["params"]:
                n = p.numel()
                p.grad = b["flat"][off:off + n].view_as(p)
                off += n
        plan["active"] = True

    def finish(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Call after backward: launches the buckets whose hooks did not all fire (unused parameters) and waits."""
        params = [p for p in params if p.requires_grad]
        if not params or self.world_size == 1:
            return
        plan = self._plan(params)
        if not plan["active"]:
            raise RuntimeError("GradBucketReducer.finish() without prepare()")
        plan["active"] = False
        for b in plan["buckets"]:
            if b["work"] is None:
                b["work"] = self._all_reduce_avg(b["flat"])
        for b in plan["buckets"]:
            work, scale = b["work"]
            work.wait()
            if scale is not None:
                b["flat"].mul_(scale)
            b["work"] = None

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
