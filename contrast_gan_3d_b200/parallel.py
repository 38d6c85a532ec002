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
      into flat fp32 buckets (no pack / unpack kernels), every bucket is all-reduced asynchronously as soon as its last
      gradient has been produced, i.e. overlapped with the rest of backward (buckets follow reverse registration order
      = the order in which backward produces them), and `finish` only waits.

    Between `prepare` and `finish` the reducer is also the **weight-gradient sink** of the convolution Functions
    (`ops.set_grad_sink`): a conv weight gradient is not returned to autograd but accumulated by the wgrad kernel itself
    (`beta = 1`) straight into the parameter's bucket slice, on a SIDE STREAM, so that the tensor-core-bound wgrad of
    layer L runs concurrently with the HBM-bound BatchNorm backward passes of layer L-1 on the compute stream.
    `finish` joins the side stream.  This part also runs at world size 1 (no collective), which is what the single-GPU
    bench uses.  Bucket size: G's 4.14 MB of gradients split into ~3 buckets (1.5 MB) so that the first all-reduces
    hide under the wgrads of the early layers (NVLink moves a bucket in ~5 us; the cost is launch latency).
    """

    def __init__(self, process_group=None, bucket_bytes: int = 3 << 19, direct_wgrad: bool = True):
        self.pg = process_group
        self.bucket_bytes = bucket_bytes
        self.direct_wgrad = direct_wgrad
        self._flat = {}
        self._plans = {}   # id(first param) -> bucket plan of a parameter set
        self._hooked = set()
        self._side = {}    # device index -> side stream for the weight-gradient kernels
        self._active_plan = None

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
            plan = dict(buckets=buckets, index=index, active=False, direct_done=set())
            self._plans[key] = plan
            for p in params:
                if id(p) not in self._hooked:
                    self._hooked.add(id(p))
                    p.register_post_accumulate_grad_hook(self._make_hook(plan))
        return plan

    def _make_hook(self, plan):
        def hook(param):
            if not plan["active"] or id(param) in plan["direct_done"]:
                return
            self._param_ready(plan, param)
        return hook

    def _param_ready(self, plan, param):
        b = plan["buckets"][plan["index"][id(param)]]
        b["ready"] += 1
        if b["ready"] == len(b["params"]) and b["work"] is None and self.world_size > 1:
            b["work"] = self._launch(b["flat"])

    def _launch(self, flat: torch.Tensor):
        """All-reduce one bucket once everything enqueued so far on the compute stream AND on the wgrad side stream has
        produced its gradients: the collective is issued from the side stream after that stream has been ordered behind the
        compute stream's current tail."""
        if flat.is_cuda and flat.device.index in self._side:
            side = self._side[flat.device.index]
            side.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(side):
                return self._all_reduce_avg(flat)
        return self._all_reduce_avg(flat)

    # ------------------------------------------------------------------ weight-gradient sink (ops.ConvBlockFn / GenTailFn)
    def side_stream(self, device) -> "torch.cuda.Stream":
        idx = torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
        st = self._side.get(idx)
        if st is None:
            st = self._side[idx] = torch.cuda.Stream(device=idx)
        return st

    def accepts(self, param) -> bool:
        """True when `param`'s gradient should be accumulated in place by the wgrad kernel (between prepare and finish)."""
        plan = self._active_plan
        return (self.direct_wgrad and plan is not None and plan["active"] and id(param) in plan["index"] and param.is_cuda
                and param.grad is not None and param.grad.dtype == torch.float32 and param.grad.is_contiguous())

    def direct_done(self, param) -> None:
        """The wgrad kernel of `param` has been enqueued on the side stream (its gradient never passes through autograd)."""
        plan = self._active_plan
        if id(param) in plan["direct_done"]:
            return  # a second accumulation into the same parameter (a module applied twice): already counted
        plan["direct_done"].add(id(param))
        self._param_ready(plan, param)

    def _all_reduce_avg(self, flat: torch.Tensor):
        if flat.is_cuda:  # NCCL averages inside the collective
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True), None
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), 1.0 / self.world_size

    def prepare(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Call after zero_grad and before backward: gradients become zeroed views into the flat buckets."""
        params = [p for p in params if p.requires_grad]
        if not params or (self.world_size == 1 and not (self.direct_wgrad and params[0].is_cuda)):
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
        plan["direct_done"] = set()
        plan["active"] = True
        self._active_plan = plan
        if self.direct_wgrad and params[0].is_cuda:
            from . import ops

            self.side_stream(params[0].device)
            ops.set_grad_sink(self)

    def finish(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Call after backward: launches the buckets whose hooks did not all fire (unused parameters) and waits."""
        params = [p for p in params if p.requires_grad]
        if not params or (self.world_size == 1 and not (self.direct_wgrad and params[0].is_cuda)):
            return
        plan = self._plan(params)
        if not plan["active"]:
            raise RuntimeError("GradBucketReducer.finish() without prepare()")
        plan["active"] = False
        self._active_plan = None
        if self.direct_wgrad and params[0].is_cuda:
            from . import ops

            ops.set_grad_sink(None)
            ops.forget_forward_uses(params)  # applications that were never back-propagated are stale from here on
        if self.world_size > 1:
            for b in plan["buckets"]:
                if b["work"] is None:
                    b["work"] = self._launch(b["flat"])
            for b in plan["buckets"]:
                work, scale = b["work"]
                work.wait()
                if scale is not None:
                    b["flat"].mul_(scale)
                b["work"] = None
        if params[0].is_cuda and params[0].device.index in self._side:  # join the weight-gradient stream
            torch.cuda.current_stream(params[0].device).wait_stream(self._side[params[0].device.index])

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
