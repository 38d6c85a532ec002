"""Trainer with the reference's constructor, method names, loss-dict keys and checkpoint layout
(reference trainer/Trainer.py:34-363); the step internals run on libcgan3d kernels.

Differences, all deliberate and listed in DESIGN.md:
  * `opt_hat = subopt - G(subopt)` is fused into the generator's tanh epilogue when the generator offers
    `forward_corrected` (Trainer.py:170-171);
  * ZNCC + HU are one fused pass when the HU loss is ours (Trainer.py:152-153);
  * the critic weight clip is fused into FusedAdam when the critic optimizer is ours (Trainer.py:136-138);
  * checkpoints additionally carry `critic_state_dict` (the reference's "discriminator" entry is always None
    because of an attribute-name mismatch, Trainer.py:311-319 vs :89); the reference can still load ours;
  * optional `grad_reducer` (data-parallel gradient averaging, no reference counterpart).
"""
from __future__ import annotations

import logging
from functools import partial
from pathlib import Path
from typing import Dict, List, Optional, Union

import numpy as np
import torch
from torch import Tensor, nn

from .. import ops
from ..model.loss import HULoss, WassersteinLoss, ZNCCLoss, fused_similarity_and_hu
from ..model.utils import wgan_gradient_penalty
from ..optim import FusedAdam, FusedRMSprop

logger = logging.getLogger(__name__)

SCAN_TYPE_ORDER = (0, -1, 1)  # ScanType iteration order OPT, LOW, HIGH (reference alias.py:24-27)


def find_latest_checkpoint(ckpt_dir: Union[Path, str]) -> Optional[Path]:
    """Highest-numbered `<int>.pt` in the directory (reference trainer/utils.py:26-34)."""
    nums = []
    for f in Path(ckpt_dir).glob("*.pt"):
        try:
            nums.append(int(f.stem))
        except ValueError:
            pass
    return None if not nums else Path(ckpt_dir) / f"{max(nums)}.pt"


class NullLogger:
    """Logger interface stand-in (reference trainer/logger/LoggerInterface.py:14-32)."""

    class _Inner:
        def __init__(self):
            self.records = []

        def log_loss(self, losses, iteration, stage):
            self.records.append((stage, iteration, {k: float(v) for k, v in losses.items()}))

    def __init__(self):
        self.logger = NullLogger._Inner()

    def __call__(self, *a, **k):
        pass

    def end_hook(self):
        pass


class Trainer:
    def __init__(
        self,
        train_iterations: int,
        val_iterations: int,
        validate_every: int,
        train_generator_every: int,
        train_critic_every: int,
        log_every: int,
        log_images_every: int,
        generator_class: partial,
        critic_class: partial,
        generator_optim_class: partial,
        critic_optim_class: partial,
        hu_loss_instance: nn.Module,
        logger_interface,
        device: torch.device,
        debug: bool = False,
        checkpoint_dir: Optional[Union[str, Path]] = None,
        weight_clip: Optional[float] = None,
        generator_lr_scheduler_class: Optional[partial] = None,
        critic_lr_scheduler_class: Optional[partial] = None,
        hu_loss_weight: float = 1.0,
        sim_loss_weight: float = 1.0,
        gan_loss_weight: float = 1.0,
        gp_weight: float = 10,
        checkpoint_every: Optional[int] = 1000,
        rng: Optional[np.random.Generator] = None,
        grad_reducer=None,
        hu_scaler=None,
    ):
        self.rng = rng
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            # an index-less "cuda" never compares equal to a tensor's "cuda:N": pin it down once so that device-resident
            # batches are recognised as such
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.debug = debug
        self.train_log_sample_size, self.val_log_sample_size = None, None
        self.train_iterations = train_iterations
        self.val_iterations = val_iterations
        self.val_every = validate_every
        self.train_generator_every = train_generator_every
        self.train_critic_every = train_critic_every
        self.log_every = log_every
        self.log_images_every = log_images_every
        self.hu_loss_w = hu_loss_weight
        self.sim_loss_w = sim_loss_weight
        self.gan_loss_w = gan_loss_weight
        self.gp_w = gp_weight
        # Optional (ours): batches may carry RAW int16 HU patches in "data" (the reference's on-disk format,
        # data/CCTADataLoader.py:76-92) instead of scaled fp32; they cross PCIe at 2 bytes per voxel and are scaled on the
        # device with `hu_scaler` ((hu - shift) / factor, bit-identical to data/Scaler.py:41-42 in fp32).
        self.hu_scaler = hu_scaler
        self._defer_schedulers = False
        self._graphs = None  # enable_cuda_graph()
        self.weight_clip = weight_clip
        self.gp_eps_fn = None  # optional source of the WGAN-GP interpolation coefficients (tests); default torch.rand

        # construction order G then D matters for seeded-init parity (reference Trainer.py:83,89)
        self.generator: nn.Module = generator_class().to(self.device)
        self.optimizer_G = generator_optim_class(self.generator.parameters())
        self.lr_scheduler_G = generator_lr_scheduler_class
        if self.lr_scheduler_G is not None:
            self.lr_scheduler_G = self.lr_scheduler_G(self.optimizer_G)
        self.critic: nn.Module = critic_class().to(self.device)
        self.optimizer_D = critic_optim_class(self.critic.parameters())
        self.lr_scheduler_D = critic_lr_scheduler_class
        if self.lr_scheduler_D is not None:
            self.lr_scheduler_D = self.lr_scheduler_D(self.optimizer_D)

        self.loss_GAN = WassersteinLoss()
        self.loss_similarity = ZNCCLoss()
        self.loss_HU = hu_loss_instance
        self.logger_interface = logger_interface
        self.grad_reducer = grad_reducer
        if grad_reducer is None and self.device.type == "cuda":
            # world size 1: no collective; the reducer is only the side-stream weight-gradient sink (parallel.py)
            from ..parallel import GradBucketReducer

            self.grad_reducer = GradBucketReducer()

        self.iteration = 0
        self.checkpoint_every = checkpoint_every
        self.checkpoint_dir = checkpoint_dir
        if self.checkpoint_dir is not None:
            self.checkpoint_dir = Path(self.checkpoint_dir)
            self.checkpoint_dir.mkdir(exist_ok=True, parents=True)
            self.load_checkpoint(find_latest_checkpoint(self.checkpoint_dir))

    # ---------------------------------------------------------------- critic / generator updates
    def train_critic(self, real: Tensor, reconstructions: Tensor, retain_graph: bool) -> Dict[str, Tensor]:
        self.optimizer_D.zero_grad(set_to_none=True)
        overlap = self.grad_reducer is not None and hasattr(self.grad_reducer, "prepare")
        if overlap:
            self.grad_reducer.prepare(self.critic.parameters())  # grads become views into the all-reduce buckets
        real_logits = self.critic(real)
        fake_logits = self.critic(reconstructions.detach())
        loss_critic = self.gan_loss_w * self.loss_GAN(fake_logits, real_logits)
        if self.weight_clip is None:
            # WGAN-GP (reference Trainer.py:122-130).  The reference passes `reconstructions` un-detached, which also sends the
            # penalty's gradient into the generator; those generator gradients are dead (train_generator zeroes them before
            # its own backward, Trainer.py:147), so the interpolates are built from the detached reconstructions here.
            loss_critic = loss_critic + wgan_gradient_penalty(real, reconstructions.detach(), self.critic, device=self.device,
                                                              lambda_=self.gp_w, rng=self.rng, eps_fn=self.gp_eps_fn)
        loss_critic.backward()
        if overlap:
            self.grad_reducer.finish(self.critic.parameters())
        elif self.grad_reducer is not None:
            self.grad_reducer.reduce(self.critic.parameters())
        if isinstance(self.optimizer_D, (FusedAdam, FusedRMSprop)):
            self.optimizer_D.step(clip=self.weight_clip or 0.0)  # update + clamp(+-clip) in one kernel
        else:
            self.optimizer_D.step()
            if self.weight_clip is not None:
                for p in self.critic.parameters():
                    p.data.clamp_(-self.weight_clip, self.weight_clip)
        if self.lr_scheduler_D is not None and not self._defer_schedulers:
            self.lr_scheduler_D.step()
        return {"D": loss_critic}

    def train_generator(self, inputs: Tensor, reconstructions: Tensor, centerlines_masks: Tensor) -> Dict[str, Tensor]:
        self.optimizer_G.zero_grad(set_to_none=True)
        overlap = self.grad_reducer is not None and hasattr(self.grad_reducer, "prepare")
        if overlap:
            self.grad_reducer.prepare(self.generator.parameters())
        # The critic's parameter gradients of this pass are dead (the reference lets autograd compute them and zeroes them
        # before their next use, Trainer.py:111): freeze the critic so that only its dgrad chain runs.
        critic_params = [p for p in self.critic.parameters() if p.requires_grad]
        for p in critic_params:
            p.requires_grad_(False)
        try:
            loss_G = self.gan_loss_w * -self.loss_GAN(self.critic(reconstructions))
        finally:
            for p in critic_params:
                p.requires_grad_(True)
        if isinstance(self.loss_HU, HULoss):
            loss_sim, loss_hu = fused_similarity_and_hu(reconstructions, inputs, centerlines_masks, self.loss_HU,
                                                        self.sim_loss_w, self.hu_loss_w)
        else:
            loss_sim = self.sim_loss_w * self.loss_similarity(reconstructions, inputs)
            loss_hu = self.hu_loss_w * self.loss_HU(reconstructions, centerlines_masks)
        full_loss_G = loss_G + loss_sim + loss_hu
        full_loss_G.backward()
        if overlap:
            self.grad_reducer.finish(self.generator.parameters())
        elif self.grad_reducer is not None:
            self.grad_reducer.reduce(self.generator.parameters())
        self.optimizer_G.step()
        if self.lr_scheduler_G is not None and not self._defer_schedulers:
            self.lr_scheduler_G.step()
        return {"G": loss_G, "G-full": full_loss_G, "sim": loss_sim, "HU": loss_hu}

    def _upload_cat(self, a: Tensor, b: Tensor) -> Tensor:
        """torch.cat([a, b]).to(device) (reference Trainer.py:166,182) without the pageable host-side concatenation:
        each part goes straight from its (pinned) host buffer into its slice of one device tensor."""
        out = torch.empty((a.shape[0] + b.shape[0], *a.shape[1:]), dtype=a.dtype, device=self.device)
        out[: a.shape[0]].copy_(a, non_blocking=True)
        out[a.shape[0]:].copy_(b, non_blocking=True)
        return out

    def _side_upload(self, key: str, parts: List[Tensor], consumed_now: bool = True):
        """Host->device copy of `torch.cat(parts)` on a side stream, so that tensors that are only needed later in the step
        (the real batch for the critic, the masks for the generator loss) cross PCIe while G's forward pass runs.
        The destination is a persistent per-key staging tensor (no cross-stream traffic through the caching allocator):
        the copy waits for the previous step's consumers, the consumer stream must wait on the returned event.
        Returns (device tensor, event or None)."""
        if all(self._on_device(t) for t in parts):
            return (parts[0] if len(parts) == 1 else torch.cat(parts)), None
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging, self._staging_done, self._staging_used = {}, {}, set()
        if consumed_now:
            self._staging_used.add(key)
        main = torch.cuda.current_stream(self.device)
        shape = (sum(t.shape[0] for t in parts), *parts[0].shape[1:])
        buf = self._staging.get(key)
        if buf is None or tuple(buf.shape) != shape or buf.dtype != parts[0].dtype:
            buf = self._staging[key] = torch.empty(shape, dtype=parts[0].dtype, device=self.device)
            self._staging_done.pop(key, None)
            self._copy_stream.wait_stream(main)  # order the copy after the allocation
        done = self._staging_done.get(key)
        if done is not None:
            self._copy_stream.wait_event(done)
        if any(t.is_cuda for t in parts):
            # a device-resident part was produced by kernels on the compute stream: the side-stream copy must see them
            self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            o = 0
            for t in parts:
                buf[o:o + t.shape[0]].copy_(t, non_blocking=True)
                o += t.shape[0]
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return buf, ev

    def _on_device(self, t: Tensor) -> bool:
        return t.is_cuda == (self.device.type == "cuda") and (not t.is_cuda or t.device.index == self.device.index)

    def _release_staging(self):
        """Mark the staging tensors as consumed by everything enqueued so far on the compute stream."""
        if getattr(self, "_copy_stream", None) is None:
            return
        main = torch.cuda.current_stream(self.device)
        # only the buffers this step has read: the one `prefetch` is about to fill for the next step was released a step ago
        for key in getattr(self, "_staging_used", ()):
            ev = torch.cuda.Event()
            ev.record(main)
            self._staging_done[key] = ev
        self._staging_used = set()

    def prefetch(self, patches: List[dict]) -> None:
        """Start the host->device copies of the NEXT step's batch (generator input [low; high], the real batch and the
        masks) on the side stream while the current step is still running.  `fit` calls this with the batch it has just
        drawn for the following iteration; `train_step` recognises the batch by the identity of its host tensors and only
        waits on the copies' events.  Two alternating sets of staging tensors: the step in flight still reads the other
        set (the similarity loss uses the generator input until the end of the step)."""
        opt, low, high = patches
        groups = {"subopt": [low["data"], high["data"]], "opt": [opt["data"]]}
        if low.get("seg") is not None and high.get("seg") is not None:
            groups["mask"] = [low["seg"], high["seg"]]
        if all(self._on_device(t) for parts in groups.values() for t in parts):
            return
        flip = 1 - getattr(self, "_prefetch_flip", 1)
        self._prefetch_flip = flip
        got = {}
        for name, parts in groups.items():
            key = f"{name}{flip}"
            buf, ev = self._side_upload(key, parts, consumed_now=False)
            got[name] = (tuple(id(t) for t in parts), buf, ev, key)
        self._prefetched = got

    def _take_prefetched(self, name: str, parts: List[Tensor]):
        """(device tensor, event) of a part uploaded by `prefetch` for exactly these host tensors, else None."""
        pf = getattr(self, "_prefetched", None)
        if not pf or name not in pf:
            return None
        ids, buf, ev, key = pf.pop(name)
        if ids != tuple(id(t) for t in parts):
            return None
        self._staging_used.add(key)
        return buf, ev

    def _scaled(self, t: Optional[Tensor], out: Optional[Tensor] = None) -> Optional[Tensor]:
        """fp32 network input from a device batch: int16 raw HU -> (hu - shift) / factor on the device, fp32 passes through
        (copied when a destination is given)."""
        if t is None:
            return t
        if t.dtype != torch.int16:
            if out is not None:
                out.copy_(t, non_blocking=True)
                return out
            return t
        if self.hu_scaler is None:
            raise ValueError("int16 (raw HU) batches need Trainer(hu_scaler=FactorZeroCenterScaler(...))")
        t = t.contiguous()
        if out is None:
            out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
        ops.call("cgan3d_scale_i16", t.data_ptr(), out.data_ptr(), t.numel(), float(self.hu_scaler.shift),
                 float(getattr(self.hu_scaler, "factor", 1)), ops._st())
        return out

    def _generate(self, subopt: Tensor):
        if hasattr(self.generator, "forward_corrected"):
            return self.generator.forward_corrected(subopt)
        attenuation = self.generator(subopt)
        return attenuation, subopt - attenuation

    def _step_core(self, subopt: Tensor, opt_t: Optional[Tensor], mask_t: Optional[Tensor], do_train_critic: bool,
                   do_train_generator: bool):
        """The device part of reference Trainer.train_step (Trainer.py:168-185) on device-resident fp32 inputs."""
        attenuation, opt_hat = self._generate(subopt)
        log_dict: Dict[str, Tensor] = {}
        if do_train_critic:
            log_dict = self.train_critic(opt_t, opt_hat, do_train_generator)
        if do_train_generator:
            log_dict |= self.train_generator(subopt, opt_hat, mask_t)
        return log_dict, attenuation, opt_hat

    def train_step(self, patches: List[dict], iteration: int) -> Dict[str, Tensor]:
        opt, low, high = patches
        ops.forget_forward_uses(self.generator.parameters())  # a generator graph of a critic-only iteration is never back-propagated
        do_train_generator = iteration % self.train_generator_every == 0
        do_train_critic = iteration % self.train_critic_every == 0
        main = torch.cuda.current_stream(self.device)
        # parts copied during the previous step (when `prefetch` was called with this batch) only need their events waited on
        pf = self._take_prefetched("subopt", [low["data"], high["data"]])
        if pf is not None:
            subopt = pf[0]
            if pf[1] is not None:
                main.wait_event(pf[1])
        else:
            subopt = self._upload_cat(low["data"], high["data"])  # needed first: on the compute stream
        opt_t = opt_ev = mask_t = mask_ev = None
        if do_train_critic:
            opt_t, opt_ev = self._take_prefetched("opt", [opt["data"]]) or self._side_upload("opt", [opt["data"]])
        if do_train_generator:
            mask_t, mask_ev = (self._take_prefetched("mask", [low["seg"], high["seg"]])
                               or self._side_upload("mask", [low["seg"], high["seg"]]))
        self._prefetched = None

        if self._graphs is not None:
            log_dict, attenuation, opt_hat = self._graphed_step(subopt, opt_t, opt_ev, mask_t, mask_ev, do_train_critic, do_train_generator)
        else:
            subopt = self._scaled(subopt)
            if opt_ev is not None:
                main.wait_event(opt_ev)
            if mask_ev is not None:
                main.wait_event(mask_ev)
            log_dict, attenuation, opt_hat = self._step_core(subopt, self._scaled(opt_t), mask_t, do_train_critic, do_train_generator)
        self._release_staging()

        if self.log_every and iteration % self.log_every == 0:
            self.logger_interface.logger.log_loss({k: v.detach().mean() for k, v in log_dict.items()}, iteration, "train")
        if self.log_images_every and iteration % self.log_images_every == 0:
            self.maybe_set_log_images_sample_size("train", patches[0]["data"].shape)
            cut = len(low["data"])
            self.logger_interface(patches, [None, opt_hat[:cut], opt_hat[cut:]], [None, attenuation[:cut], attenuation[cut:]],
                                  list(SCAN_TYPE_ORDER), iteration, "train", self.train_log_sample_size)
        return log_dict

    # ---------------------------------------------------------------- CUDA graph of the step
    def enable_cuda_graph(self, warmup: int = 3) -> None:
        """Replay the device part of train_step from a CUDA graph (ours; the reference has no counterpart).  The ~300
        kernels of a step are enqueued by ONE launch, which takes the Python / ctypes / autograd host cost (~8 ms per step)
        off the critical path: a small-batch step (BASELINE config C4, 8 pairs per GPU) is otherwise host-bound.
        One graph per (critic?, generator?, input shapes) variant, captured after `warmup` eager steps of that variant;
        inputs are copied (or scaled from int16) into static tensors before each replay; learning rate and Adam step
        count live on the device (FusedAdam), LR schedulers keep running on the host after each replay.
        The loss tensors returned by train_step are the graph's static outputs: read them before the next step."""
        if self.device.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device")
        for o in (self.optimizer_G, self.optimizer_D):
            if not isinstance(o, FusedAdam):
                raise NotImplementedError("the captured step needs FusedAdam (device-resident lr / step count)")
        self._graphs = {"warmup": int(warmup), "entries": {}}
        self._defer_schedulers = True

    def _graphed_step(self, subopt_in, opt_in, opt_ev, mask_in, mask_ev, do_c: bool, do_g: bool):
        main = torch.cuda.current_stream(self.device)
        key = (do_c, do_g, tuple(subopt_in.shape), subopt_in.dtype, None if opt_in is None else (tuple(opt_in.shape), opt_in.dtype),
               None if mask_in is None else (tuple(mask_in.shape), mask_in.dtype))
        ent = self._graphs["entries"].get(key)
        if ent is None:
            f32 = dict(dtype=torch.float32, device=self.device)
            ent = dict(sub=torch.empty(tuple(subopt_in.shape), **f32), opt=None if opt_in is None else torch.empty(tuple(opt_in.shape), **f32),
                       mask=None if mask_in is None else torch.empty_like(mask_in, device=self.device), graph=None, out=None, seen=0,
                       launches=0)
            self._graphs["entries"][key] = ent
        # inputs -> the graph's static tensors (on the compute stream, after everything the previous replay enqueued)
        self._scaled(subopt_in, out=ent["sub"])
        if opt_in is not None:
            if opt_ev is not None:
                main.wait_event(opt_ev)
            self._scaled(opt_in, out=ent["opt"])
        if mask_in is not None:
            if mask_ev is not None:
                main.wait_event(mask_ev)
            ent["mask"].copy_(mask_in, non_blocking=True)
        # Everything the graphed path hands back is DETACHED: a loss tensor that still referenced the step's autograd graph
        # would keep its AccumulateGrad nodes (bound to the stream of that step) alive into the capture, where autograd
        # would synchronise the capturing stream with the default stream and invalidate the capture.
        core = lambda: tuple(({k: v.detach() for k, v in o.items()} if isinstance(o, dict) else o.detach())
                             for o in self._step_core(ent["sub"], ent["opt"], ent["mask"], do_c, do_g))
        if ent["graph"] is not None:
            ent["graph"].replay()
            self.optimizer_G.note_graph_replay() if do_g else None
            self.optimizer_D.note_graph_replay() if do_c else None
            out = ent["out"]
        elif ent["seen"] < self._graphs["warmup"]:
            ent["seen"] += 1
            out = core()
        else:
            from .. import _lib

            self.optimizer_G.sync_device_hyper(); self.optimizer_D.sync_device_hyper()
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = core()
            ent["launches"] = _lib.launch_count - n0
            ent["graph"], ent["out"] = g, out
            g.replay()  # the capture only recorded this step; run it (its optimizer host counters were advanced while recording)
        # host-side schedule bookkeeping (reference Trainer.py:139-140,158-159), then mirror a changed rate to the device
        if do_c and self.lr_scheduler_D is not None:
            self.lr_scheduler_D.step()
            self.optimizer_D.sync_device_hyper()
        if do_g and self.lr_scheduler_G is not None:
            self.lr_scheduler_G.step()
            self.optimizer_G.sync_device_hyper()
        return out

    def release_cuda_graphs(self) -> None:
        """Drop the captured graphs (and their static tensors / private memory pool) and return to eager steps.  Call this
        before tearing down a process group whose collectives were recorded in the graphs."""
        if self._graphs is not None:
            torch.cuda.synchronize(self.device)
            self._graphs["entries"].clear()
            self._graphs = None
            self._defer_schedulers = False

    def graph_launches_per_step(self) -> int:
        """libcgan3d entry-point calls recorded in the captured graphs (bench.py: gpu_launches)."""
        if self._graphs is None:
            return 0
        return max([e["launches"] for e in self._graphs["entries"].values()] + [0])

    # ---------------------------------------------------------------- loop
    def fit(self, train_loaders: Dict[int, object], val_loaders: Dict[int, object], profiler=None):
        self.generator.train()
        self.critic.train()
        augmenters = {"train": train_loaders, "val": val_loaders}
        self._manage_augmenters(augmenters, "start")
        draw = lambda: [next(train_loaders[st]) for st in SCAN_TYPE_ORDER]
        upcoming = draw() if self.iteration < self.train_iterations else None
        for iteration in range(self.iteration, self.train_iterations):
            patches = upcoming
            self.train_step(patches, iteration)
            # same draw order as the reference loop (Trainer.py:206-209), one iteration early: the next batch crosses PCIe
            # while this step's kernels run
            upcoming = draw() if iteration + 1 < self.train_iterations else None
            if upcoming is not None:
                self.prefetch(upcoming)
            if self.val_every is not None and iteration != 0 and iteration % self.val_every == 0:
                self.validate(val_loaders, iteration)
            if self.checkpoint_every is not None and iteration != 0 and iteration % self.checkpoint_every == 0:
                self.save_checkpoint(iteration)
            if profiler:
                profiler.step()
        if profiler:
            profiler.stop()
        if self.checkpoint_every is not None and self.checkpoint_dir is not None:
            self.save_checkpoint(self.train_iterations)
        self._manage_augmenters(augmenters, "end")
        self.logger_interface.end_hook()

    def validate(self, val_loaders: Dict[int, object], train_iteration: int):
        self.critic.eval()
        self.generator.eval()
        z = torch.zeros(4, dtype=torch.float32, device=self.device)
        loss_sim, loss_G, loss_real_C, loss_fake_C = z.chunk(4)
        loggable = []
        with torch.no_grad():
            for i in range(self.val_iterations):
                for st in SCAN_TYPE_ORDER:
                    batch = next(val_loaders[st])
                    sample = batch["data"].to(self.device, non_blocking=True)
                    if st == 0:
                        loss_real_C -= self.loss_GAN(self.critic(sample))
                    else:
                        attenuation, sample_hat = self._generate(sample)
                        loss_fake = self.loss_GAN(self.critic(sample_hat))
                        loss_fake_C += loss_fake
                        loss_G -= loss_fake
                        loss_sim += self.loss_similarity(sample_hat, sample)
                    if i == 0 and st != 0 and loggable is not None:
                        # first validation iteration: the LOW and HIGH batches go to the image logger (reference
                        # Trainer.py:278-296)
                        loggable.append([batch, sample_hat, attenuation])
                        if len(loggable) == len(SCAN_TYPE_ORDER) - 1:
                            patches, reconstructions, attenuations = list(zip(*loggable))
                            self.maybe_set_log_images_sample_size("val", patches[0]["data"].shape)
                            self.logger_interface(patches, list(reconstructions), list(attenuations), list(SCAN_TYPE_ORDER)[1:],
                                                  train_iteration, "validation", self.val_log_sample_size)
                            loggable = None
        self.critic.train()
        self.generator.train()
        val_loss = {"D": (loss_real_C + loss_fake_C) / self.val_iterations,
                    "G": loss_G / (self.val_iterations * 2),
                    "sim": loss_sim / (self.val_iterations * 2)}
        self.logger_interface.logger.log_loss(val_loss, train_iteration, "validation")
        return val_loss

    # ---------------------------------------------------------------- checkpoints
    @property
    def model_torch_attrs(self) -> List[str]:
        return ["generator", "optimizer_G", "lr_scheduler_G", "discriminator", "optimizer_D", "lr_scheduler_D"]

    def save_checkpoint(self, iteration: int):
        state = {"iteration": iteration}
        for attr in self.model_torch_attrs:
            el = getattr(self, attr, None)  # "discriminator" stays None exactly as in the reference
            state[attr] = el if el is None else el.state_dict()
        state["critic_state_dict"] = self.critic.state_dict()
        torch.save(state, self.checkpoint_dir / f"{iteration}.pt")

    def load_checkpoint(self, ckpt_path: Optional[Path]):
        if ckpt_path is not None and Path(ckpt_path).is_file():
            checkpoint: dict = torch.load(ckpt_path, map_location="cpu")
            for k, v in checkpoint.items():
                if k == "critic_state_dict":
                    self.critic.load_state_dict(v)
                elif k in self.model_torch_attrs:
                    if v is not None and hasattr(self, k):
                        getattr(self, k).load_state_dict(v)
                else:
                    setattr(self, k, v)
        logger.info("Starting from iteration %d", self.iteration)

    def _manage_augmenters(self, augmenters, event: str):
        assert event in ["start", "end"]
        for mode, d in augmenters.items():
            if d is None or (mode == "val" and self.val_every is None):
                continue
            for aug in d.values():
                if event == "start" and hasattr(aug, "restart"):
                    aug.restart()
                elif event == "end" and hasattr(aug, "_finish"):
                    aug._finish()

    def maybe_set_log_images_sample_size(self, mode: str, batch_shape):
        name = f"{mode}_log_sample_size"
        if getattr(self, name) is None:
            bs = batch_shape[-1 if len(batch_shape) == 5 else 0]
            setattr(self, name, min(bs, 64))
