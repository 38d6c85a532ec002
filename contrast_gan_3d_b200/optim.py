"""FusedAdam: torch.optim.Adam's single-tensor math (defaults eps=1e-8, no weight decay, no amsgrad) as one
libcgan3d launch per 48 parameter tensors (pointer table in the kernel parameters), with the WGAN weight clip of
reference trainer/Trainer.py:136-138 fused in.

The learning rate and the step count are kept in a small DEVICE tensor per parameter group (torch's
`capturable=True` idea), so that an optimizer step captured in a CUDA graph stays valid when a scheduler changes the
rate; the host-side `state[p]["step"]` / `group["lr"]` remain the source of truth for checkpoints and are mirrored to
the device whenever they change."""
from __future__ import annotations

import ctypes as C

import torch
from torch.optim import Optimizer

from . import ops
from ._lib import call


class FusedAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip: float = 0.0):
        # the torch.optim.Adam keys are carried (at their no-op values) so that a checkpoint written here can be stepped by
        # the reference's torch.optim.Adam after load_state_dict, and vice versa (reference trainer/Trainer.py:311-339)
        defaults = dict(lr=lr, betas=betas, eps=eps, clip=clip, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._dev = {}  # id(group) -> dict(hyper=device float[2] {lr, step}, lr=mirrored lr, step=mirrored step)

    def __setstate__(self, state):
        super().__setstate__(state)
        self._dev = {}
        for group in self.param_groups:  # param groups loaded from a torch.optim.Adam checkpoint have no "clip"
            group.setdefault("clip", 0.0)
            if group.get("weight_decay", 0) or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("FusedAdam implements Adam without weight decay / amsgrad / maximize (the reference's settings)")
        for st in self.state.values():  # torch.optim.Adam stores `step` as a tensor
            if torch.is_tensor(st.get("step")):
                st["step"] = int(st["step"].item())

    def _hyper(self, group, device, lr: float, step_before: int) -> torch.Tensor:
        """The group's device-side {lr, step}; host values that changed since the last call (scheduler, load_state_dict)
        are written with fill kernels.  Nothing is written while a CUDA graph is being captured: the caller syncs outside."""
        d = self._dev.get(id(group))
        if d is None or d["hyper"].device != device:
            d = dict(hyper=torch.tensor([lr, float(step_before)], dtype=torch.float32, device=device), lr=lr, step=step_before)
            self._dev[id(group)] = d
        if d["lr"] != lr or d["step"] != step_before:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: lr / step changed on the host while a CUDA graph is being captured; call "
                                   "sync_device_hyper() before the capture")
            if d["lr"] != lr:
                d["hyper"][0:1].fill_(lr)
                d["lr"] = lr
            if d["step"] != step_before:
                d["hyper"][1:2].fill_(float(step_before))
                d["step"] = step_before
        return d["hyper"]

    def sync_device_hyper(self) -> None:
        """Mirror group["lr"] to the device copies (call after an LR scheduler step when the optimizer step itself is
        replayed from a CUDA graph and therefore does not run this Python code)."""
        for group in self.param_groups:
            d = self._dev.get(id(group))
            if d is not None and d["lr"] != float(group["lr"]):
                d["hyper"][0:1].fill_(float(group["lr"]))
                d["lr"] = float(group["lr"])

    def note_graph_replay(self) -> None:
        """A captured step() has been replayed: advance the host-side step counts that the device already advanced."""
        for group in self.param_groups:
            d = self._dev.get(id(group))
            if d is None:
                continue
            d["step"] += 1
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] += 1

    @torch.no_grad()
    def step(self, closure=None, clip: float | None = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            c = group.get("clip", 0.0) if clip is None else clip
            by_step = {}  # tensors that share a step count go into one multi-tensor launch
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if torch.is_tensor(st["step"]):
                    st["step"] = int(st["step"].item())
                st["step"] += 1
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                ops._need_cuda(p, g)
                if not p.data.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous parameters")
                by_step.setdefault(st["step"], []).append((p.data, g, st["exp_avg"], st["exp_avg_sq"]))
            for step, items in by_step.items():
                n = len(items)
                tabs = [(C.c_void_p * n)(*[t[k].data_ptr() for t in items]) for k in range(4)]
                numels = (C.c_int64 * n)(*[t[0].numel() for t in items])
                if len(by_step) == 1:
                    # the usual case (every parameter has been stepped equally often): device-resident lr / step
                    hyper = self._hyper(group, items[0][0].device, float(group["lr"]), step - 1)
                    call("cgan3d_adam_tick", hyper.data_ptr(), ops._st())
                    self._dev[id(group)]["step"] = step
                    call("cgan3d_adam_step_multi_dev", n, tabs[0], tabs[1], tabs[2], tabs[3], numels, hyper.data_ptr(), float(b1),
                         float(b2), float(group["eps"]), float(c or 0.0), ops._st())
                else:
                    call("cgan3d_adam_step_multi", n, tabs[0], tabs[1], tabs[2], tabs[3], numels, float(group["lr"]), float(b1),
                         float(b2), float(group["eps"]), int(step), float(c or 0.0), ops._st())
        return loss


class FusedRMSprop(Optimizer):
    """torch.optim.RMSprop with its default settings (alpha 0.99, eps 1e-8, no momentum, not centered, no weight decay:
    what reference experiments/rmsprop_conf.py:8-9 constructs) as one libcgan3d launch per 48 tensors, with the critic
    weight clip fused in like FusedAdam."""

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, clip: float = 0.0):
        defaults = dict(lr=lr, alpha=alpha, eps=eps, clip=clip, weight_decay=0, momentum=0, centered=False, capturable=False,
                        foreach=None, maximize=False, differentiable=False)
        super().__init__(params, defaults)

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("clip", 0.0)
            if group.get("momentum", 0) or group.get("centered", False) or group.get("weight_decay", 0):
                raise NotImplementedError("FusedRMSprop implements plain RMSprop (the reference's settings)")
        for st in self.state.values():
            if torch.is_tensor(st.get("step")):
                st["step"] = int(st["step"].item())

    @torch.no_grad()
    def step(self, closure=None, clip: float | None = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            c = group.get("clip", 0.0) if clip is None else clip
            items = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                ops._need_cuda(p, g)
                items.append((p.data, g, st["square_avg"]))
            if items:
                n = len(items)
                tabs = [(C.c_void_p * n)(*[t[k].data_ptr() for t in items]) for k in range(3)]
                numels = (C.c_int64 * n)(*[t[0].numel() for t in items])
                call("cgan3d_rmsprop_step_multi", n, tabs[0], tabs[1], tabs[2], numels, float(group["lr"]), float(group["alpha"]),
                     float(group["eps"]), float(c or 0.0), ops._st())
        return loss
