"""Training engine for the fused path: one flat fp32 parameter / gradient buffer for every
network and the camera poses, gradients accumulated by the backward kernels straight into it,
one NCCL all-reduce per step (rays shard across GPUs) and one fused Adam launch
with the reference's per-group exponential learning-rate schedule
(barf/model_interpolation.py:543-584).  It replaces, for the hot path only, what Lightning's
automatic optimisation does for the reference; the modules remain usable under Lightning."""
import ctypes as C
import math
from typing import List, Optional

import torch as th
import torch.nn as nn

from ._lib import LR_EXPONENTIAL, LR_LE_NICE, NbAdamGroup, check, lib
from .fused_mlp import FlatParams
from .model_interpolation import le_nice_lr, log_decay_factor
from .parallel import allreduce_sum_, global_mean_scale


def exponential_lr(lr0: float, log_gamma: float, n: int, step: int) -> float:
    """Learning rate of optimiser step `step` (1-based) under torch's ExponentialLR(gamma,
    last_epoch=-n-1) as GARF constructs it (garf/model_garf.py:365-428), with the scheduler semantics of
    the torch in this image (2.11; pinned by tests/test_host_logic.py against the real scheduler): one
    factor gamma per scheduler step, none at construction. (torch 2.0.0, the reference's pin, also
    multiplied once at construction and skipped the step on which last_epoch passed through 0:
    gamma^(step - [step > n]) instead of gamma^(step - 1).)"""
    return lr0 * math.exp(log_gamma * (step - 1))


class TrainEngine:
    def __init__(self, model, device, process_group=None, betas=(0.9, 0.999), eps: float = 1e-5,
                 loss_fn=None):
        """model: NerfInterpolation (or a subclass carrying `camera_extrinsics` parameters), or any module
        exposing `param_groups` and fused networks (`fused_networks()`).

        loss_fn(*batch) -> (loss, logs): the differentiable loss of one step and a dict of device scalars
        to report (must contain "loss_fine"); default = pose transform + forward + MSE on a
        (o, d, target, img_idx, pixel_width) batch. A model with `training_loss` (BarfModel) passes that.
        It must not synchronise with the host: the step is captured into a CUDA graph."""
        self.model = model
        self.device = th.device(device)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (th.distributed.is_available() and th.distributed.is_initialized()):
            self.world = th.distributed.get_world_size(process_group)
        self.betas, self.eps = betas, eps
        self.step_count = 0
        self.loss_fn = loss_fn
        self.coarse_weight = 1.0

        # one flat buffer: [radiance | proposal | poses]; groups follow model.param_groups
        groups = []
        params: List[nn.Parameter] = []
        for g in model.param_groups:
            ps = [p for p in g["parameters"]] if not isinstance(g["parameters"], list) else g["parameters"]
            g["parameters"] = ps          # generators are single-use: keep the list
            begin = sum(p.numel() for p in params)
            params += ps
            if g.get("schedule", "le_nice") == "exponential":
                # torch ExponentialLR(gamma, last_epoch=-n-1), garf/model_garf.py:365-428
                groups.append(dict(begin=begin, end=begin + sum(p.numel() for p in ps), lr0=g["learning_rate_start"],
                                   logf=math.log(g["gamma"]), n=int(g["learning_rate_decay_end"]),
                                   wd=g.get("weight_decay", 0.0), mode=LR_EXPONENTIAL))
            else:
                logf = log_decay_factor(g["learning_rate_start"], g["learning_rate_stop"], g["learning_rate_decay_end"])
                groups.append(dict(begin=begin, end=begin + sum(p.numel() for p in ps), lr0=g["learning_rate_start"],
                                   logf=logf, n=int(g["learning_rate_decay_end"] or 0) if logf != 0.0 else 0,
                                   wd=g.get("weight_decay", 0.0), mode=LR_LE_NICE))
        self.groups = groups
        self.flat = FlatParams(params)
        model.to(self.device)
        self.flat.ensure(self.device)
        nets = model.fused_networks() if hasattr(model, "fused_networks") else \
            [model.model_radiance, getattr(model, "model_proposal", None)]
        for net in nets:
            if net is not None:
                net.fused_field(flat=self.flat)
        # gradient buffer + one slot behind it for the loss: the slot rides in the same all-reduce, so
        # every rank sees a NaN of any rank and the fused Adam skips the step on all of them
        self.grad_all = th.zeros(self.flat.numel + 4, device=self.device, dtype=th.float32)
        self.grad = self.grad_all[: self.flat.numel]
        self.loss_slot = self.grad_all[self.flat.numel: self.flat.numel + 1]
        self.exp_avg = th.zeros_like(self.grad)
        self.exp_avg_sq = th.zeros_like(self.grad)
        self.state = th.zeros(4, device=self.device, dtype=th.int64)   # [steps, skipped steps, scratch]
        self.flat.grad_sink = self.grad
        cam = getattr(model, "camera_extrinsics", None)
        self.pose_sink = None
        if cam is not None and hasattr(cam, "rotation"):
            o_r, o_t = self.flat.offset_of(cam.rotation), self.flat.offset_of(cam.translation)
            self.pose_sink = (self.grad[o_r:o_r + cam.rotation.numel()].view_as(cam.rotation),
                              self.grad[o_t:o_t + cam.translation.numel()].view_as(cam.translation))
            cam.grad_sink = self.pose_sink
        self._groups_c = (NbAdamGroup * len(groups))(*[
            NbAdamGroup(begin=g["begin"], end=g["end"], n_steps=g["n"], lr0=g["lr0"], log_factor=g["logf"],
                        weight_decay=g["wd"], mode=g["mode"]) for g in groups])
        self._graph = None
        self.sync_replicas()

    # -- pieces ------------------------------------------------------------------------------
    def sync_replicas(self):
        """Rank 0's parameters, moments and step counters to every rank (what DDP's construction-time
        broadcast does for the reference under Lightning): replicas that were built from different
        random states would otherwise train apart silently."""
        if self.world > 1:
            import torch.distributed as dist
            for t in (self.flat.flat, self.exp_avg, self.exp_avg_sq, self.state):
                dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
            counts = self.state[:2].tolist()
            self.step_count = int(counts[0])
            self.flat.version += 1

    def learning_rates(self, step: int):
        """lr of every group for optimiser step number `step` (1-based), as the device evaluates it.
        torch's LRScheduler performs one scheduler step at construction, so the reference's k-th
        optimiser step runs with the closed form evaluated at _step_count = k."""
        out = []
        for g in self.groups:
            if g["mode"] == LR_EXPONENTIAL:
                out.append(exponential_lr(g["lr0"], g["logf"], g["n"], step))
            else:
                out.append(g["lr0"] * math.exp(g["logf"] * min(step, g["n"])))
        return out

    def skipped_steps(self) -> int:
        """Steps the non-finite-loss guard turned into no-ops so far (one device read)."""
        return int(self.state[1].item())

    def _default_loss(self, o, d, target, img_idx=None, pixel_width=None):
        m = self.model
        cam = getattr(m, "camera_extrinsics", None)
        if self.pose_sink is not None and img_idx is not None:
            from . import ops
            o, d, _, _ = ops.pose_forward(cam.rotation, cam.translation, img_idx, o, d, self.pose_sink)
        fine, coarse = m.forward(o, d, pixel_width)
        loss_fine = nn.functional.mse_loss(fine, target)
        loss = loss_fine
        if coarse is not None:
            loss = loss + self.coarse_weight * nn.functional.mse_loss(coarse, target)
        return loss, {"loss_fine": loss_fine.detach()}

    def forward_loss(self, o, d, target, img_idx=None, pixel_width=None, coarse_weight: float = 1.0):
        self.coarse_weight = coarse_weight
        loss, logs = self._default_loss(o, d, target, img_idx, pixel_width)
        return loss, logs["loss_fine"]

    def _loss(self, batch):
        if self.loss_fn is not None:
            return self.loss_fn(*batch)
        return self._default_loss(*batch)

    def _pre_step(self):
        """Host-side schedules of the model (coarse-to-fine alpha, blur weights): small in-place
        device writes issued before the step (outside the captured graph)."""
        upd = getattr(self.model, "update_schedules", None)
        if upd is not None:
            upd(self.step_count)

    def _device_step(self, batch):
        """gradient clear -> loss -> backward -> all-reduce -> fused Adam; no host synchronisation."""
        nvtx = th.cuda.nvtx
        self.grad_all.zero_()
        nvtx.range_push("nerfb200.forward")
        loss, logs = self._loss(batch)
        nvtx.range_pop()
        nvtx.range_push("nerfb200.backward")
        loss.backward()
        self.loss_slot.copy_(loss.detach().reshape(1))
        nvtx.range_pop()
        nvtx.range_push("nerfb200.optimizer")
        self.optimizer_step()
        nvtx.range_pop()
        return logs

    def optimizer_step(self):
        if self.world > 1:
            allreduce_sum_(self.grad_all, self.pg)   # NCCL over NVLink, one call per step (gradients + loss slot)
        with th.cuda.device(self.device):
            check(lib().nerfb200_adam_step_dev(self.flat.flat.data_ptr(), self.grad.data_ptr(),
                                               self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                               self.flat.numel, self._groups_c, len(self.groups),
                                               self.betas[0], self.betas[1], self.eps,
                                               global_mean_scale(self.world), self.loss_slot.data_ptr(),
                                               self.state.data_ptr(), th.cuda.current_stream().cuda_stream),
                  "adam_step_dev")
        self.flat.version += 1     # the packed bf16 weight images are stale now

    def step(self, *batch, coarse_weight: float = 1.0):
        """One optimisation step on this rank's shard of rays; returns the (fine) loss tensor."""
        self.coarse_weight = coarse_weight
        self._pre_step()
        logs = self._device_step(batch)
        self.step_count += 1
        self.last_logs = logs
        return logs["loss_fine"]

    # -- the whole step as one CUDA graph ---------------------------------------------------------
    def capture(self, *batch, coarse_weight: float = 1.0):
        """Captures `step` (gradient clear, forward, loss, backward, the NCCL all-reduce of a multi-process
        engine, fused Adam) over static copies of the given batch into ONE CUDA graph; `replay(batch)`
        then costs a few small copies and one graph launch on the host instead of ~40 launches — what
        bounds small batches (1024 rays x 64 samples: 0.79 ms eager) and what kept eight ranks from
        scaling end to end (every rank's host jitter reaches all ranks through the all-reduce).
        Requirements: at least one eager `step` with a batch of the same shape before (lazy
        initialisations, NCCL communicator), and later batches of that shape. Step counter, learning-
        rate schedules, Adam's bias corrections and the non-finite-loss guard live in device memory
        (`nerfb200_adam_step_dev`), so replays follow them with no per-step host work. By-value launch
        parameters are frozen at capture time: `pixel_width_sigma` of the integrated encodings
        (MipBarf's sigma schedule) and `sigma_bias` — re-capture when they change."""
        if self.step_count < 1:
            raise RuntimeError("capture: run one eager step() with a batch of this shape first")
        self.coarse_weight = coarse_weight
        # static inputs of the graph: views of ONE flat device buffer, so that a host batch arrives with a
        # single host-to-device copy (`replay_packed`) instead of one copy per tensor
        self._static_layout, off = [], 0
        for t in batch:
            if t is None:
                self._static_layout.append(None)
                continue
            nbytes = t.numel() * t.element_size()
            self._static_layout.append((off, nbytes, t.dtype, tuple(t.shape)))
            off += (nbytes + 255) // 256 * 256
        self._static_flat = th.zeros(max(off, 256), device=self.device, dtype=th.uint8)
        self._static = []
        for t, lay in zip(batch, self._static_layout):
            if lay is None:
                self._static.append(None)
                continue
            view = self._static_flat[lay[0]: lay[0] + lay[1]].view(lay[2]).view(lay[3])
            view.copy_(t.detach())
            self._static.append(view)
        from ._lib import launch_count
        import gc
        self.last_logs = None
        gc.collect()
        th.cuda.synchronize(self.device)
        snap = [t.clone() for t in (self.flat.flat, self.exp_avg, self.exp_avg_sq, self.state)]

        def restore():
            for dst, src in zip((self.flat.flat, self.exp_avg, self.exp_avg_sq, self.state), snap):
                dst.copy_(src)
            self.flat.version += 1                # the next forward re-packs the weight images

        # One eager run of the step ON THE CAPTURE STREAM first, then the capture on the same stream:
        # whatever autograd keeps alive between steps (gradient accumulators of parameters a model holds
        # a graph on, e.g. GarfModel's proposal cdf) then belongs to the capture stream. Accumulators
        # left over from eager steps on the default stream make the autograd engine synchronise the
        # capturing stream with it at the end of backward, which invalidates the capture
        # (cudaErrorStreamCaptureIsolation, observed on B200 / torch 2.11).
        cs = th.cuda.Stream(device=self.device)
        cs.wait_stream(th.cuda.current_stream(self.device))
        with th.cuda.stream(cs):
            self._device_step(self._static)
            restore()
        cs.synchronize()
        gc.collect()
        before = launch_count()
        self._graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self._graph, stream=cs, capture_error_mode="thread_local"):
            logs = self._device_step(self._static)
            self._static_logs = {k: v.detach() for k, v in logs.items()}
        self.launches_per_replay = launch_count() - before     # kernels of this library inside one replay
        with th.cuda.stream(cs):
            restore()                             # capture does not execute; keep the engine state exact anyway
        th.cuda.current_stream(self.device).wait_stream(cs)
        return self

    def replay_packed(self, host_flat: th.Tensor):
        """One captured step on a batch that sits in ONE pinned host buffer laid out as `static_layout()`
        says: a single host-to-device copy and a graph launch."""
        self._pre_step()
        self._static_flat.copy_(host_flat, non_blocking=True)
        self._graph.replay()
        self.step_count += 1
        self.flat.version += 1
        self.last_logs = self._static_logs
        return self._static_logs["loss_fine"]

    def pack_batch(self, batch, pin: bool = False) -> th.Tensor:
        """The batch as one flat uint8 buffer in the layout of the graph's static inputs (on the device
        of its tensors, or in pinned host memory for host tensors with `pin`): what `replay_packed` takes."""
        layout, nbytes = self.static_layout()
        first = next(t for t in batch if t is not None)
        flat = th.zeros(nbytes, dtype=th.uint8, device=first.device)
        if pin and not first.is_cuda:
            flat = flat.pin_memory()
        for t, lay in zip(batch, layout):
            if lay is not None:
                flat[lay[0]: lay[0] + lay[1]].view(lay[2]).view(lay[3]).copy_(t)
        return flat

    def static_layout(self):
        """[(byte offset, bytes, dtype, shape) | None] of the captured batch inside the flat input buffer,
        and the buffer's size in bytes."""
        return list(self._static_layout), int(self._static_flat.numel())

    def release_graph(self):
        """Drops the captured graph. A graph that contains NCCL collectives keeps the communicator busy:
        `torch.distributed.destroy_process_group()` blocks (observed on 2 x B200, torch 2.11 / NCCL 2.28)
        until every such graph has been destroyed — call this before tearing the process group down."""
        if self._graph is not None:
            th.cuda.synchronize(self.device)
            self._graph = None
            self._static_logs = None
            import gc
            gc.collect()
            th.cuda.synchronize(self.device)

    def replay(self, *batch):
        """One captured step on a new batch (same shapes as at capture); returns the (fine) loss tensor,
        which the NEXT replay overwrites (as all of `last_logs`)."""
        self._pre_step()
        for dst, src in zip(self._static, batch):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.step_count += 1
        self.flat.version += 1                    # for a later eager call: the packed images are stale
        self.last_logs = self._static_logs
        return self._static_logs["loss_fine"]

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
        resumes from a checkpoint the reference wrote. (`loops` is left out: Lightning restores loop
        progress only when the key is present, and this engine has no loop state to offer; not verified
        against a real Lightning install, which this image lacks.) Steps the non-finite-loss guard
        skipped count for the schedulers but not for Adam's per-parameter `step`, as in the reference."""
        params = self._param_list()
        state, groups, idx = {}, [], 0
        lrs = self.learning_rates(self.step_count + 1)      # what the scheduler has set for the NEXT step
        step_t = th.tensor(float(self.step_count - self.skipped_steps()))
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
                "callbacks": {}, "hparams_name": "kwargs",
                "hyper_parameters": dict(getattr(self.model, "hparams", None) or {})}

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
        adam_steps = steps.pop() if steps else int(ckpt.get("global_step", 0))
        sched = (ckpt.get("lr_schedulers") or [None])[0]
        self.step_count = max(int(sched["last_epoch"]) if sched and "last_epoch" in sched else adam_steps, adam_steps)
        self.state.copy_(th.tensor([self.step_count, self.step_count - adam_steps, 0, 0], dtype=th.int64))
        self.sync_replicas()


class HostStepper:
    """Training from HOST batches without stalling the device: the host->device copy of batch
    i + 1 runs on a copy stream while step i computes, and the loss of every step is copied to
    pinned memory asynchronously and handed back `depth` calls later, so the host never waits for
    a step it has just enqueued and can run `depth` steps ahead of the device (the reference's loop
    copies, computes and reads the loss back to back, barf/model_interpolation.py:490-526, 588-597).
    With `use_graph` the step is the engine's captured CUDA graph (captured on first use, after one
    eager step), so a step costs the host a handful of copies and one graph launch.

        stepper = HostStepper(engine)
        for batch in pinned_host_batches:        # tuples (o, d, target, img_idx, pixel_width)
            loss_of_an_earlier_step = stepper.submit(batch)   # None for the first `depth` calls
        remaining_losses = stepper.drain()
    """

    def __init__(self, engine: TrainEngine, coarse_weight: float = 1.0, depth: int = 1, use_graph: bool = False):
        self.engine = engine
        self.coarse_weight = coarse_weight
        self.depth = max(int(depth), 1)
        self.use_graph = use_graph
        n = self.depth + 1
        self.copy_stream = th.cuda.Stream(device=engine.device)
        self.staging = [None] * n
        self.ready = [th.cuda.Event() for _ in range(n)]      # H2D of the slot has landed
        self.done = [None] * n                                # the step that used the slot has finished
        self.loss_host = [th.zeros(1).pin_memory() for _ in range(n)]
        self.count = 0
        self.read = 0
        self.h2d_bytes = 0
        self._packed = None
        self._packed_views = None

    def _read(self, slot: int) -> float:
        self.done[slot].synchronize()
        return float(self.loss_host[slot][0])

    def _submit_packed(self, host_batch):
        """Graph mode: the batch is packed into one pinned buffer per slot (host memcpy) and reaches the
        graph's static inputs with ONE host-to-device copy on the compute stream; no copy stream, no
        per-tensor copies, no device-to-device staging."""
        n = self.depth + 1
        s = self.count % n
        eng = self.engine
        compute = th.cuda.current_stream(eng.device)
        layout, nbytes = eng.static_layout()
        if self._packed is None:
            self._packed = [th.zeros(nbytes, dtype=th.uint8).pin_memory() for _ in range(n)]
            self._packed_views = [[None if lay is None else buf[lay[0]: lay[0] + lay[1]].view(lay[2]).view(lay[3])
                                   for lay in layout] for buf in self._packed]
        if self.done[s] is not None:
            self.done[s].synchronize()                        # the copy that read this slot has long finished
        for dst, src in zip(self._packed_views[s], host_batch):
            if dst is not None:
                dst.copy_(src)                                # host memcpy into the slot
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in host_batch if t is not None)
        loss = eng.replay_packed(self._packed[s])
        self.loss_host[s].copy_(loss.reshape(1), non_blocking=True)
        ev = th.cuda.Event()
        ev.record(compute)
        self.done[s] = ev
        self.count += 1
        if self.count - self.read > self.depth:
            out = self._read(self.read % n)
            self.read += 1
            return out
        return None

    def submit(self, host_batch):
        n = self.depth + 1
        s = self.count % n
        eng = self.engine
        if self.use_graph and eng._graph is not None:
            return self._submit_packed(host_batch)
        compute = th.cuda.current_stream(eng.device)
        with th.cuda.stream(self.copy_stream):
            if self.done[s] is not None:
                self.copy_stream.wait_event(self.done[s])     # the slot's previous step has consumed it
            if self.staging[s] is None:
                self.staging[s] = tuple(None if t is None else th.empty(t.shape, dtype=t.dtype, device=eng.device)
                                        for t in host_batch)
            for dst, src in zip(self.staging[s], host_batch):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in host_batch if t is not None)
        compute.wait_event(self.ready[s])
        if self.use_graph and eng.step_count >= 1:
            if eng._graph is None:
                eng.capture(*self.staging[s], coarse_weight=self.coarse_weight)
            loss = eng.replay(*self.staging[s])
        else:
            loss = eng.step(*self.staging[s], coarse_weight=self.coarse_weight)
        self.loss_host[s].copy_(loss.reshape(1), non_blocking=True)
        ev = th.cuda.Event()
        ev.record(compute)
        self.done[s] = ev
        self.count += 1
        # read an EARLIER step's loss only now: `depth` steps are queued behind it, so the device does
        # not idle while the host waits
        if self.count - self.read > self.depth:
            out = self._read(self.read % n)
            self.read += 1
            return out
        return None

    def drain(self):
        """Losses of the steps whose read-back is still outstanding, oldest first."""
        out = []
        n = self.depth + 1
        while self.read < self.count:
            out.append(self._read(self.read % n))
            self.read += 1
        return out

    def flush(self):
        """Loss of the most recent step (reads back everything outstanding)."""
        rest = self.drain()
        return rest[-1] if rest else None
