"""NerfInterpolation: the render module (uniform t-sampling -> field -> alpha compositing ->
pdf resampling -> field -> compositing) behind the reference's constructor, attributes, forward
signature and Lightning hooks (reference barf/model_interpolation.py:30-597).  Every stage runs
as a hand-written CUDA kernel through the C ABI; there are no host syncs on the path."""
import math
import warnings
from typing import Literal, Optional

import torch as th
import torch.nn as nn

from . import ops
from ._lightning_compat import LightningModule
from .field_function import field_rays
from .model_interpolation_architecture import NerfModel

uniform_sampling_strategies = Literal["stratified_uniform", "equidistant"]
integration_strategies = Literal["left", "middle"]

MAGIC_NUMBER = 1 / 3  # barf/magic.py:2 (the compositing kernel applies it, COMPOSITE_BARF flavour)


class DummyCamEx(nn.Module):
    def forward(self, i, o, d):
        return o, d, None, None


class SchedulerLeNice(th.optim.lr_scheduler.LRScheduler):
    """Per-group exponential decay in closed form (barf/model_interpolation.py:30-67):
    lr_i(step) = start_i * exp(min(step, n_i) * (ln stop_i - ln start_i) / n_i), constant when
    n_i is 0/None or start_i is 0."""

    def __init__(self, optimizer, start_LR, stop_LR=None, number_of_steps=None, verbose=False):
        self.start_LR = start_LR
        self.stop_LR = stop_LR
        self.number_of_steps = number_of_steps
        self.log_decay_factors = [log_decay_factor(start_LR[i], None if stop_LR is None else stop_LR[i],
                                                   None if number_of_steps is None else number_of_steps[i])
                                  for i, _ in enumerate(optimizer.param_groups)]
        self.decay_factors = [math.exp(f) for f in self.log_decay_factors]
        super().__init__(optimizer)   # `verbose` no longer exists in current torch

    def get_lr(self):
        return self._get_closed_form_lr()

    def _get_closed_form_lr(self):
        return [le_nice_lr(self.start_LR[i], self.log_decay_factors[i],
                           None if self.number_of_steps is None else self.number_of_steps[i], self._step_count)
                for i in range(len(self.start_LR))]


def log_decay_factor(start: float, stop: Optional[float], n_steps) -> float:
    if stop is None or n_steps in (0, None) or start == 0:
        return 0.0
    return (math.log(stop) - math.log(start)) / n_steps


def le_nice_lr(start: float, log_factor: float, n_steps, step: int) -> float:
    if log_factor == 0.0:
        return start
    return start * math.exp(log_factor * min(step, n_steps))


class NerfInterpolation(LightningModule):
    def __init__(self, near_sphere_normalized: float, far_sphere_normalized: float,
                 model_radiance: NerfModel, samples_per_ray_radiance: int,
                 uniform_sampling_strategy: uniform_sampling_strategies = "stratified_uniform",
                 uniform_sampling_offset_size: float = 0.,
                 integration_strategy: integration_strategies = "middle",
                 model_proposal: Optional[NerfModel] = None, samples_per_ray_proposal: int = 0):
        LightningModule.__init__(self)
        self.save_hyperparameters(ignore=["model_radiance", "model_proposal"])
        self.near_sphere_normalized = near_sphere_normalized
        self.far_sphere_normalized = far_sphere_normalized
        self.samples_per_ray_radiance = samples_per_ray_radiance
        self.samples_per_ray_proposal = samples_per_ray_proposal
        self.uniform_sampling_strategy = uniform_sampling_strategy
        self.uniform_sampling_offset_size = uniform_sampling_offset_size
        self.integration_strategy = integration_strategy
        self.model_radiance = model_radiance
        self.model_proposal = model_proposal
        self.camera_extrinsics = DummyCamEx()
        self.proposal = samples_per_ray_proposal > 0
        models = [model_radiance, model_proposal] if self.proposal else [model_radiance]
        self.param_groups = [group for m in models for group in m.param_groups]

    # ---- sampling ----------------------------------------------------------------------------
    def _get_intervals(self, t: th.Tensor):
        t_end = th.empty_like(t)
        t_end[:, :-1] = t[:, 1:]
        t_end[:, -1] = self.far_sphere_normalized
        return t, t_end

    def _sample_t_stratified_uniform(self, batch_size: int, n_samples: int, strategy: str, offset_size: float):
        """Bin starts / ends (B,S).  The uniforms come from torch's generator in the reference's
        draw order ((B,S) jitter first, then the (B,1) offset), so a seeded run consumes the
        same random numbers as the reference on the same device."""
        if strategy == "stratified_uniform":
            jitter = th.rand((batch_size, n_samples), device=self.device)
        elif strategy == "equidistant":
            jitter = None
        else:
            raise ValueError(f"sampling_strategy must be one of {uniform_sampling_strategies.__args__}, was '{strategy}'")
        offset_u = th.rand((batch_size, 1), device=self.device) if offset_size != 0 else None
        return ops.sample_uniform(self.near_sphere_normalized, self.far_sphere_normalized, batch_size, n_samples,
                                  self.device, jitter, offset_u, offset_size)

    def _sample_t_pdf_weighted(self, t_coarse, weights, distances_coarse, n_samples: int):
        """Deterministic pdf resampling; the reference's whole-batch equidistant fallback (taken
        when any ray violates its postcondition) is applied on the device by a gated kernel, so
        its (B,1) offset uniforms are drawn unconditionally."""
        fallback_u = th.rand((t_coarse.shape[0], 1), device=t_coarse.device)
        return ops.resample_alloc(t_coarse, weights, distances_coarse, n_samples, self.near_sphere_normalized,
                                  self.far_sphere_normalized, fallback_u)

    def _get_t_query(self, t_start, t_end, strategy: str):
        if strategy == "left":
            return t_start
        if strategy == "middle":
            return (t_start + t_end) / 2
        raise ValueError(f"strategy must be one of {integration_strategies.__args__}, was '{strategy}'")

    def _compute_positions(self, origins, directions, t_start, t_end):
        """(B,S,3) positions / directions.  Kept for callers of the reference surface; the hot
        path never materialises them (the fused kernel computes x = o + t d in registers)."""
        t = self._get_t_query(t_start, t_end, self.integration_strategy)
        positions = origins.unsqueeze(1) + t.unsqueeze(2) * directions.unsqueeze(1)
        return positions, directions.unsqueeze(1).expand(-1, positions.shape[1], -1).contiguous()

    # ---- rendering ---------------------------------------------------------------------------
    def _render_rays(self, densities, colors, distances):
        return ops.render_rays(densities, colors, distances)

    def _compute_color(self, model, t_start, t_end, ray_origs, ray_dirs, pixel_width, batch_size, samples_per_ray):
        if self.integration_strategy not in ("left", "middle"):
            raise ValueError(f"strategy must be one of {integration_strategies.__args__}, was '{self.integration_strategy}'")
        sample_dist = t_end - t_start
        density, color = field_rays(model, ray_origs, ray_dirs, t_start, t_end, pixel_width, self.integration_strategy)
        rgb, weights = self._render_rays(density, color, sample_dist)
        return rgb, weights, sample_dist

    def forward(self, ray_origs: th.Tensor, ray_dirs: th.Tensor, pixel_width: th.Tensor = None):
        """(rgb_fine (B,3), rgb_coarse (B,3) | None) — barf/model_interpolation.py:417-486."""
        batch_size = ray_origs.shape[0]
        if self.proposal:
            t_c0, t_c1 = self._sample_t_stratified_uniform(batch_size, self.samples_per_ray_proposal,
                                                           self.uniform_sampling_strategy, self.uniform_sampling_offset_size)
            rgb_coarse, weights, dist_c = self._compute_color(self.model_proposal, t_c0, t_c1, ray_origs, ray_dirs,
                                                              pixel_width, batch_size, self.samples_per_ray_proposal)
            t_f0, t_f1 = self._sample_t_pdf_weighted(t_c0, weights, dist_c, self.samples_per_ray_radiance)
            rgb_fine, _, _ = self._compute_color(self.model_radiance, t_f0, t_f1, ray_origs, ray_dirs, pixel_width,
                                                 batch_size, self.samples_per_ray_radiance)
        else:
            t_f0, t_f1 = self._sample_t_stratified_uniform(batch_size, self.samples_per_ray_radiance,
                                                           self.uniform_sampling_strategy, self.uniform_sampling_offset_size)
            rgb_fine, _, _ = self._compute_color(self.model_radiance, t_f0, t_f1, ray_origs, ray_dirs, pixel_width,
                                                 batch_size, self.samples_per_ray_radiance)
            rgb_coarse = None
        return rgb_fine, rgb_coarse

    # ---- the inner fused op: forward(rays_o, rays_d, near, far) -> rgb, depth, weights ------------
    @th.no_grad()
    def render(self, ray_origs: th.Tensor, ray_dirs: th.Tensor, pixel_width: th.Tensor = None,
               near: float = None, far: float = None):
        """(rgb (B,3), depth (B,), weights (B,S)) of the radiance pass — BASELINE.json's north-star
        signature (near / far default to the module's planes). Each pass is ONE call of the C ABI
        (`nerfb200_render_rays`: sampling -> fused field -> compositing); with a proposal network the
        coarse pass's weights feed the pdf resampler and the fine pass renders on its bins.
        depth = sum w t_mid / max(sum w, eps) (what GarfModel returns through nerfacc, garf/model_garf.py:
        223-236; the barf module itself never computes one)."""
        import ctypes as C
        from ._lib import COMPOSITE_BARF, check, lib
        from .field_function import _prep
        near = self.near_sphere_normalized if near is None else near
        far = self.far_sphere_normalized if far is None else far
        o = ops._f32(ray_origs, "ray_origs")
        d = ops._f32(ray_dirs, "ray_dirs", tuple(o.shape))
        B, dev = o.shape[0], o.device
        pw = None if pixel_width is None else _prep(pixel_width, B, 1, dev)
        t_mode = {"left": 0, "middle": 1}[self.integration_strategy]
        if self.uniform_sampling_strategy not in ("stratified_uniform", "equidistant"):
            raise ValueError(f"sampling_strategy must be one of {uniform_sampling_strategies.__args__}")

        def one_pass(model, S, t0=None, t1=None):
            field = model.fused_field()
            cm = field.prepare(dev)
            cp, cd = field.pe_cfgs()
            jitter = offset_u = None
            if t0 is None:
                if self.uniform_sampling_strategy == "stratified_uniform":
                    jitter = th.rand((B, S), device=dev)
                if self.uniform_sampling_offset_size != 0:
                    offset_u = th.rand((B, 1), device=dev).reshape(-1)
            nbytes = C.c_longlong()
            check(lib().nerfb200_render_rays_workspace_bytes(B, S, C.byref(nbytes)), "render_rays_workspace_bytes")
            ws = th.empty(max(nbytes.value, 4), device=dev, dtype=th.uint8)
            rgb, depth = th.empty((B, 3), device=dev), th.empty((B,), device=dev)
            w = th.empty((B, S), device=dev)
            with th.cuda.device(dev):
                check(lib().nerfb200_render_rays(
                    C.byref(cm.program), field.wpack.data_ptr(), field.bias.data_ptr(), cm.bias_floats, C.byref(cp),
                    C.byref(cd), ops._ptr(field.pe_pos.alpha_tensor()), ops._ptr(field.pe_dir.alpha_tensor()),
                    float(field.sigma_bias), o.data_ptr(), d.data_ptr(), ops._ptr(pw), B, S, float(near), float(far),
                    ops._ptr(t0), ops._ptr(t1), ops._ptr(jitter), ops._ptr(offset_u),
                    float(self.uniform_sampling_offset_size), t_mode, COMPOSITE_BARF, ws.data_ptr(), rgb.data_ptr(),
                    depth.data_ptr(), w.data_ptr(), None, th.cuda.current_stream().cuda_stream), "render_rays")
            return rgb, depth, w, ws

        if not self.proposal:
            rgb, depth, w, _ = one_pass(self.model_radiance, self.samples_per_ray_radiance)
            return rgb, depth, w
        Sc = self.samples_per_ray_proposal
        _, _, w_c, ws = one_pass(self.model_proposal, Sc)
        n = B * Sc
        ws_f = ws.view(th.float32)
        t_c0, dist_c = ws_f[:n].view(B, Sc), ws_f[2 * n: 3 * n].view(B, Sc)
        t_f0, t_f1 = self._sample_t_pdf_weighted(t_c0, w_c, dist_c, self.samples_per_ray_radiance)
        rgb, depth, w, _ = one_pass(self.model_radiance, self.samples_per_ray_radiance, t_f0, t_f1)
        return rgb, depth, w

    # ---- Lightning surface -------------------------------------------------------------------
    def _step_helper(self, batch, batch_idx, purpose: Literal["train", "val"]):
        (ray_origs_raw, ray_origs_pred, ray_dirs_raw, ray_dirs_pred, ray_colors_raw, img_idx, pixel_width) = batch
        pred_fine, pred_coarse = self.forward(ray_origs_pred, ray_dirs_pred, pixel_width)
        loss = nn.functional.mse_loss(pred_fine, ray_colors_raw[:, 0])
        logs = {f"{purpose}_loss_fine": loss, f"{purpose}_psnr": self.psnr_tensor(loss)}
        if self.proposal:
            loss_coarse = nn.functional.mse_loss(pred_coarse, ray_colors_raw[:, 0])
            loss = loss + loss_coarse
            logs[f"{purpose}_loss_coarse"] = loss_coarse
        self.log_dict(logs)
        return self._nan_guard(loss)

    @staticmethod
    def _nan_guard(loss: th.Tensor) -> th.Tensor:
        """The reference's guard, verbatim in effect (barf/model_interpolation.py:522-524): a NaN loss
        is replaced by a fresh leaf, so backward reaches no parameter and the optimiser skips the step.
        This is the path Lightning's automatic optimisation drives, and like the reference it reads
        the loss back (one host sync); engine.TrainEngine applies the same rule on the device instead
        (nerfb200_adam_step_dev skips the step when the all-reduced loss is not finite)."""
        if loss.isnan():
            loss = th.tensor(1., requires_grad=True)
            warnings.warn("loss was nan - no optimization step performed")
        return loss

    def validation_transform_rays(self, ray_origs, ray_dirs, transform_params=None):
        return ray_origs, ray_dirs, transform_params

    def training_step(self, batch, batch_idx):
        return self._step_helper(batch, batch_idx, "train")

    def validation_step(self, batch, batch_idx):
        return self._step_helper(batch, batch_idx, "val")

    def configure_optimizers(self):
        optimizer = th.optim.Adam(
            [{"params": g["parameters"], "lr": g["learning_rate_start"], "weight_decay": g["weight_decay"]}
             for g in self.param_groups], eps=1e-5)
        scheduler = SchedulerLeNice(optimizer,
                                    start_LR=[g["learning_rate_start"] for g in self.param_groups],
                                    stop_LR=[g["learning_rate_stop"] for g in self.param_groups],
                                    number_of_steps=[g["learning_rate_decay_end"] for g in self.param_groups])
        return {"optimizer": optimizer,
                "lr_scheduler": {"scheduler": scheduler, "interval": "step", "frequency": 1,
                                 "name": "le_nice_lr_scheduler"}}

    @staticmethod
    def psnr_tensor(loss: th.Tensor) -> th.Tensor:
        """-10 log10(loss) on the device (NaN where the reference would refuse, loss <= 1e-7)."""
        l = loss.detach()
        return th.where(l > 1e-7, -10.0 * th.log10(l.clamp_min(1e-30)), th.full_like(l, float("nan")))

    def compute_psnr(self, loss: th.Tensor) -> float:
        """The reference's host-side PSNR (one device sync) — barf/model_interpolation.py:588-597."""
        value = float(loss.detach().cpu().item())
        if value <= 1e-7:
            print(f"WARN: Loss was {value} - psnr not computed")
            return float("nan")
        return -10 * math.log10(value)
