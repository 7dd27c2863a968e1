"""Training engine for the fused path: one flat fp32 parameter / gradient buffer for every
network and the camera poses, gradients accumulated by the backward kernels straight into it,
one NCCL all-reduce per step (rays shard across GPUs) and one fused Adam launch
with the reference's per-group exponential learning-rate schedule
(barf/model_interpolation.py:543-584).  It replaces, for the hot path only, what Lightning's
automatic optimisation does for the reference; the modules remain usable under Lightning."""
import ctypes as C
from typing import List, Optional

import torch as th
import torch.nn as nn

from ._lib import check, lib
from .fused_mlp import FlatParams
from .model_interpolation import le_nice_lr, log_decay_factor
from .parallel import allreduce_sum_, global_mean_scale


class TrainEngine:
    def __init__(self, model, device, process_group=None, betas=(0.9, 0.999), eps: float = 1e-5):
        """model: NerfInterpolation (or a subclass carrying `camera_extrinsics` parameters)."""
        self.model = model
        self.device = th.device(device)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (th.distributed.is_available() and th.distributed.is_initialized()):
            self.world = th.distributed.get_world_size(process_group)
        self.betas, self.eps = betas, eps
        self.step_count = 0

        # one flat buffer: [radiance | proposal | poses]; groups follow model.param_groups
        groups = []
        params: List[nn.Parameter] = []
        for g in model.param_groups:
            ps = [p for p in g["parameters"]] if not isinstance(g["parameters"], list) else g["parameters"]
            g["parameters"] = ps          # generators are single-use: keep the list
            begin = sum(p.numel() for p in params)
            params += ps
            groups.append(dict(begin=begin, end=begin + sum(p.numel() for p in ps),
                               lr0=g["learning_rate_start"],
                               logf=log_decay_factor(g["learning_rate_start"], g["learning_rate_stop"],
                                                     g["learning_rate_decay_end"]),
                               n=g["learning_rate_decay_end"], wd=g.get("weight_decay", 0.0)))
        self.groups = groups
        self.flat = FlatParams(params)
        model.to(self.device)
        self.flat.ensure(self.device)
        for net in (model.model_radiance, getattr(model, "model_proposal", None)):
            if net is not None:
                net.fused_field(flat=self.flat)
        self.grad = th.zeros(self.flat.numel, device=self.device, dtype=th.float32)
        self.exp_avg = th.zeros_like(self.grad)
        self.exp_avg_sq = th.zeros_like(self.grad)
        self.flat.grad_sink = self.grad
        cam = getattr(model, "camera_extrinsics", None)
        self.pose_sink = None
        if cam is not None and hasattr(cam, "rotation"):
            o_r, o_t = self.flat.offset_of(cam.rotation), self.flat.offset_of(cam.translation)
            self.pose_sink = (self.grad[o_r:o_r + cam.rotation.numel()].view_as(cam.rotation),
                              self.grad[o_t:o_t + cam.translation.numel()].view_as(cam.translation))
        n = len(groups)
        self._gb = (C.c_longlong * n)(*[g["begin"] for g in groups])
        self._ge = (C.c_longlong * n)(*[g["end"] for g in groups])
        self._gw = (C.c_float * n)(*[g["wd"] for g in groups])

    # -- pieces ------------------------------------------------------------------------------
    def learning_rates(self, step: int):
        """lr of every group for optimiser step number `step` (1-based).  torch's LRScheduler
        performs one scheduler step at construction, so the reference's k-th optimiser step
        runs with the closed form evaluated at _step_count = k."""
        return [le_nice_lr(g["lr0"], g["logf"], g["n"], step) for g in self.groups]

    def forward_loss(self, o, d, target, img_idx=None, pixel_width=None, coarse_weight: float = 1.0):
        m = self.model
        cam = getattr(m, "camera_extrinsics", None)
        if self.pose_sink is not None and img_idx is not None:
            from . import ops
            o, d, _, _ = ops.pose_forward(cam.rotation, cam.translation, img_idx, o, d, self.pose_sink)
        fine, coarse = m.forward(o, d, pixel_width)
        loss_fine = nn.functional.mse_loss(fine, target)
        loss = loss_fine
        if coarse is not None:
            loss = loss + coarse_weight * nn.functional.mse_loss(coarse, target)
        return loss, loss_fine

    def optimizer_step(self):
        self.step_count += 1
        if self.world > 1:
            allreduce_sum_(self.grad, self.pg)   # NCCL over NVLink, one call per step
        lrs = self.learning_rates(self.step_count)
        glr = (C.c_float * len(lrs))(*lrs)
        with th.cuda.device(self.device):
            check(lib().nerfb200_adam_step(self.flat.flat.data_ptr(), self.grad.data_ptr(),
                                           self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                           self.flat.numel, len(lrs), self._gb, self._ge, glr, self._gw,
                                           self.betas[0], self.betas[1], self.eps, self.step_count,
                                           global_mean_scale(self.world), th.cuda.current_stream().cuda_stream), "adam_step")
        self.flat.version += 1     # the packed bf16 weight images are stale now

    # -- the whole step as one CUDA graph ---------------------------------------------------------
    def capture(self, o, d, target, img_idx=None, pixel_width=None, coarse_weight: float = 1.0):
        """Captures `step` (gradient clear, forward, loss, backward, fused Adam) over static copies of
        the given batch into a CUDA graph; `replay(batch)` then costs two launches on the host (a
        copy of the step's schedule + the graph) instead of ~40, which is what bounds small batches
        (1024 rays x 64 samples: 0.79 ms eager). Requirements: single process (the NCCL all-reduce
        stays eager), at least one eager `step` with a batch of the same shape before (lazy
        initialisations), and later batches of that shape. The learning-rate schedule and Adam's bias
        correction are read from device memory (`nerfb200_adam_step_sched`), so replays follow them."""
        if self.world > 1:
            raise RuntimeError("capture: the all-reduce of a multi-process engine is not captured; use step()")
        if self.step_count < 1:
            raise RuntimeError("capture: run one eager step() with a batch of this shape first")
        self._static = [None if t is None else t.detach().clone() for t in (o, d, target, img_idx, pixel_width)]
        n = len(self.groups)
        # the host may run several replays ahead of the device: the schedule of a step sits in its own
        # pinned slot until the copy that reads it has executed
        self._sched_host = [th.zeros(2 + n).pin_memory() for _ in range(4)]
        self._sched_done = [None] * 4
        self._sched_dev = th.zeros(2 + n, device=self.device)
        self.flat.version += 1                    # the captured forward must contain the weight re-pack
        from ._lib import launch_count
        th.cuda.synchronize(self.device)
        before = launch_count()
        self._graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self._graph):
            self.grad.zero_()
            so, sd, st, si, sp = self._static
            loss, loss_fine = self.forward_loss(so, sd, st, si, sp, coarse_weight)
            loss.backward()
            with th.cuda.device(self.device):
                check(lib().nerfb200_adam_step_sched(
                    self.flat.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                    self.flat.numel, n, self._gb, self._ge, self._gw, self._sched_dev.data_ptr(),
                    self.betas[0], self.betas[1], self.eps, global_mean_scale(self.world),
                    th.cuda.current_stream().cuda_stream), "adam_step_sched")
            self._static_loss = loss_fine.detach()
        self.launches_per_replay = launch_count() - before     # kernels of this library inside one replay
        return self

    def replay(self, o, d, target, img_idx=None, pixel_width=None):
        """One captured step on a new batch (same shapes as at capture); returns the (fine) loss tensor,
        which the NEXT replay overwrites."""
        for dst, src in zip(self._static, (o, d, target, img_idx, pixel_width)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        self.step_count += 1
        slot = self.step_count & 3
        if self._sched_done[slot] is not None:
            self._sched_done[slot].synchronize()
        host = self._sched_host[slot]
        host[0] = 1.0 - self.betas[0] ** self.step_count
        host[1] = (1.0 - self.betas[1] ** self.step_count) ** 0.5
        for q, lr in enumerate(self.learning_rates(self.step_count)):
            host[2 + q] = lr
        self._sched_dev.copy_(host, non_blocking=True)
        ev = th.cuda.Event()
        ev.record()
        self._sched_done[slot] = ev
        self._graph.replay()
        self.flat.version += 1                    # for a later eager call: the packed images are stale
        return self._static_loss

    # -- checkpoints in the layout Lightning writes for the reference ----------------------------
    def _param_list(self):
        return [p for g in self.model.param_groups for p in g["parameters"]]

    def checkpoint(self, epoch: int = 0) -> dict:
        """The dictionary `Trainer.save_checkpoint` writes for the reference's modules
        (`ModelCheckpoint`, barf/run_barf.py:142-146): `state_dict` with the reference's keys,
        `optimizer_states` = the state dict of ONE torch.optim.Adam over `param_groups` (per-parameter
        `step` / `exp_avg` / `exp_avg_sq`, the order of `configure_optimizers`,
        barf/model_interpolation.py:543-564) and `lr_schedulers` = SchedulerLeNice's state — so the
        reference (or `configure_optimizers()` of these modules) resumes from it and this engine
        resumes from a checkpoint the reference wrote."""
        params = self._param_list()
        state, groups, idx = {}, [], 0
        lrs = self.learning_rates(self.step_count + 1)      # what the scheduler has set for the NEXT step
        step_t = th.tensor(float(self.step_count))
        for gi, g in enumerate(self.groups):
            ids = []
            for p in self.model.param_groups[gi]["parameters"]:
                o = self.flat.offset_of(p)
                if self.step_count > 0:
                    state[idx] = {"step": step_t.clone(),
                                  "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape).detach().cpu().clone(),
                                  "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape).detach().cpu().clone()}
                ids.append(idx)
                idx += 1
            groups.append({"lr": lrs[gi], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": g["wd"],
                           "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                           "differentiable": False, "fused": None, "initial_lr": g["lr0"], "params": ids})
        assert idx == len(params)
        sched = {"start_LR": [g["lr0"] for g in self.groups],
                 "stop_LR": [mg["learning_rate_stop"] for mg in self.model.param_groups],
                 "number_of_steps": [g["n"] for g in self.groups],
                 "log_decay_factors": [g["logf"] for g in self.groups],
                 "decay_factors": [float(th.exp(th.tensor(g["logf"]))) for g in self.groups],
                 "base_lrs": [g["lr0"] for g in self.groups], "last_epoch": self.step_count,
                 "_step_count": self.step_count + 1, "_get_lr_called_within_step": False, "_last_lr": lrs}
        return {"epoch": epoch, "global_step": self.step_count, "pytorch-lightning_version": "2.0.0",
                "state_dict": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                "optimizer_states": [{"state": state, "param_groups": groups}], "lr_schedulers": [sched],
                "loops": {}, "callbacks": {}, "hparams_name": "kwargs", "hyper_parameters": {}}

    def save_checkpoint(self, path: str, epoch: int = 0) -> None:
        th.save(self.checkpoint(epoch), path)

    def load_checkpoint(self, ckpt) -> None:
        """Resumes from `checkpoint()`'s dictionary, from a file of it, or from a checkpoint
        Lightning wrote for the reference's module (same keys)."""
        if isinstance(ckpt, str):
            ckpt = th.load(ckpt, map_location="cpu", weights_only=False)
        self.model.load_state_dict(ckpt["state_dict"])
        self.flat.ensure(self.device)            # load_state_dict copies in place: the views stay valid
        self.flat.version += 1
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        opt = ckpt["optimizer_states"][0]
        params = self._param_list()
        steps = set()
        for ids, in [(g["params"],) for g in opt["param_groups"]]:
            for i in ids:
                st = opt["state"].get(i)
                if st is None:
                    continue
                p = params[i]
                o = self.flat.offset_of(p)
                self.exp_avg[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise RuntimeError(f"checkpoint holds parameters at different optimiser steps {sorted(steps)}: "
                               "the fused Adam keeps one step count")
        self.step_count = steps.pop() if steps else int(ckpt.get("global_step", 0))

    def step(self, o, d, target, img_idx=None, pixel_width=None, coarse_weight: float = 1.0):
        """One optimisation step on this rank's shard of rays; returns the (fine) loss tensor."""
        self.grad.zero_()
        loss, loss_fine = self.forward_loss(o, d, target, img_idx, pixel_width, coarse_weight)
        loss.backward()
        self.optimizer_step()
        return loss_fine.detach()


class HostStepper:
    """Training from HOST batches without stalling the device: the host->device copy of batch
    i + 1 runs on a copy stream while step i computes, and the loss of every step is copied to
    pinned memory asynchronously and handed back one call later, so the host never waits for the
    step it has just enqueued (the reference's loop copies, computes and reads the loss back to
    back, barf/model_interpolation.py:490-526, 588-597).

        stepper = HostStepper(engine)
        for batch in pinned_host_batches:        # tuples (o, d, target, img_idx, pixel_width)
            loss_of_previous_step = stepper.submit(batch)   # None for the first call
        last_loss = stepper.flush()
    """

    def __init__(self, engine: TrainEngine, coarse_weight: float = 1.0):
        self.engine = engine
        self.coarse_weight = coarse_weight
        self.copy_stream = th.cuda.Stream(device=engine.device)
        self.staging = [None, None]
        self.ready = [th.cuda.Event(), th.cuda.Event()]       # H2D of the slot has landed
        self.done = [None, None]                              # the step that used the slot has finished
        self.loss_host = [th.zeros(1).pin_memory(), th.zeros(1).pin_memory()]
        self.count = 0
        self.h2d_bytes = 0

    def _read(self, slot: int) -> float:
        self.done[slot].synchronize()
        return float(self.loss_host[slot][0])

    def submit(self, host_batch):
        s = self.count & 1
        compute = th.cuda.current_stream(self.engine.device)
        with th.cuda.stream(self.copy_stream):
            if self.done[s] is not None:
                self.copy_stream.wait_event(self.done[s])     # the slot's previous step has consumed it
            if self.staging[s] is None:
                self.staging[s] = tuple(th.empty(t.shape, dtype=t.dtype, device=self.engine.device) for t in host_batch)
            for dst, src in zip(self.staging[s], host_batch):
                dst.copy_(src, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in host_batch)
        compute.wait_event(self.ready[s])
        loss = self.engine.step(*self.staging[s], coarse_weight=self.coarse_weight)
        self.loss_host[s].copy_(loss.reshape(1), non_blocking=True)
        ev = th.cuda.Event()
        ev.record(compute)
        self.done[s] = ev
        self.count += 1
        # read the PREVIOUS step's loss only now: this step is already queued behind it, so the
        # device does not idle while the host waits
        return self._read(s ^ 1) if self.count > 1 else None

    def flush(self):
        return self._read((self.count - 1) & 1) if self.count > 0 else None
