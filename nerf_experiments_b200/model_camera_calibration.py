"""Camera-calibration models — module surface of reference barf/model_camera_calibration.py
(CameraCalibrationModel), barf/model_barf.py (BarfModel) and barf/model_mip.py (MipNeRF, MipBarf):
the step glue (a14) around the fused render path. Pose alignment and the pose error run in one
CUDA launch (`ops.kabsch`) instead of ~40 torch launches, an SVD and two host syncs per step.

Lightning is optional here: the models read the training data from `self.trainer.datamodule`
when a trainer is attached (as the reference does) and otherwise from `self.datamodule`, which
may be a `ray_batcher.GpuRayBatcher`-backed object exposing `dataset_train`,
`get_blurred_pixel_colors(batch, sigma)` and, for the schedules, `n_batches` / `current_epoch`."""
from typing import Literal, Optional

import torch as th
import torch.nn as nn

from . import ops
from .model_camera_extrinsics import CameraExtrinsics
from .model_interpolation import NerfInterpolation
from .model_interpolation_architecture import NerfModel


class LoopState:
    """Stand-in for the pieces of a Lightning Trainer the step helpers read."""

    def __init__(self, datamodule, n_batches: int):
        self.datamodule = datamodule
        self.train_dataloader = range(n_batches)
        self.current_epoch = 0


class CameraCalibrationModel(NerfInterpolation):
    def __init__(self, n_training_images: int, camera_learning_rate_start: float,
                 camera_learning_rate_stop: float, camera_learning_rate_decay_end: int,
                 near_sphere_normalized: float, far_sphere_normalized: float, model_radiance: NerfModel,
                 samples_per_ray_radiance: int, model_proposal: Optional[NerfModel] = None,
                 samples_per_ray_proposal: int = 0, max_gaussian_sigma: float = 0.0,
                 uniform_sampling_strategy="stratified_uniform", uniform_sampling_offset_size: float = 0.,
                 integration_strategy="middle"):
        NerfInterpolation.__init__(self, near_sphere_normalized=near_sphere_normalized,
                                   far_sphere_normalized=far_sphere_normalized, model_radiance=model_radiance,
                                   samples_per_ray_radiance=samples_per_ray_radiance,
                                   uniform_sampling_strategy=uniform_sampling_strategy,
                                   uniform_sampling_offset_size=uniform_sampling_offset_size,
                                   integration_strategy=integration_strategy, model_proposal=model_proposal,
                                   samples_per_ray_proposal=samples_per_ray_proposal)
        self.camera_extrinsics = CameraExtrinsics(n_training_images, camera_learning_rate_start,
                                                  camera_learning_rate_stop, camera_learning_rate_decay_end)
        self.param_groups = self.param_groups + self.camera_extrinsics.param_groups
        self.max_gaussian_sigma = max_gaussian_sigma

    # -- data access ---------------------------------------------------------------------------
    def _loop(self):
        tr = getattr(self, "trainer", None)
        if tr is not None and getattr(tr, "datamodule", None) is not None:
            return tr
        loop = getattr(self, "loop", None)
        if loop is None:
            raise RuntimeError("attach a Lightning trainer or set model.loop = LoopState(datamodule, n_batches)")
        return loop

    # -- alignment -----------------------------------------------------------------------------
    def kabsch_algorithm(self, point_cloud_from: th.Tensor, point_cloud_to: th.Tensor, remove_outliers: bool = True):
        """R (3,3), t (1,3), c (1,): R @ point_cloud_from * c + t estimates point_cloud_to."""
        return ops.kabsch(point_cloud_from, point_cloud_to, remove_outliers)

    def validation_transform_rays(self, origs_val, dirs_val, post_transform_params=None):
        if post_transform_params is None:
            post_transform_params = self.compute_post_transform_params()
        R, t, c = post_transform_params
        origs_model = th.matmul(R, origs_val.unsqueeze(-1)).squeeze(-1) * c + t
        dirs_model = th.matmul(R, dirs_val.unsqueeze(-1)).squeeze(-1)
        return origs_model, dirs_model, post_transform_params

    def _train_origins(self):
        dataset = self._loop().datamodule.dataset_train
        cache = getattr(self, "_origin_cache", None)
        if cache is None or cache[0] is not dataset:
            # built once per dataset: the step then issues no host-to-device copy (and can be captured)
            img_idxs = th.tensor([dataset.index_to_index[i] for i in range(dataset.n_images)],
                                 device=self.device, dtype=th.int64)
            cache = (dataset, img_idxs, dataset.camera_origins.to(self.device).contiguous(),
                     dataset.camera_origins_noisy.to(self.device).contiguous())
            self._origin_cache = cache
        _, img_idxs, origs_raw, origs_noisy = cache
        origs_pred, _ = self.camera_extrinsics.forward_origins(img_idxs, origs_noisy)
        return origs_raw, origs_pred

    def compute_post_transform_params(self, from_raw_to_pred=True, return_origs=False, remove_outliers=True):
        origs_raw, origs_pred = self._train_origins()
        if from_raw_to_pred:
            params = self.kabsch_algorithm(origs_raw, origs_pred, remove_outliers=remove_outliers)
        else:
            params = self.kabsch_algorithm(origs_pred, origs_raw, remove_outliers=remove_outliers)
        return (params, origs_raw, origs_pred) if return_origs else params

    # -- batch transformations ------------------------------------------------------------------
    def validation_transform(self, batch):
        o_raw, _, d_raw, _, colors, img_idx, pixel_width = batch
        o_pred, d_pred, _ = self.validation_transform_rays(o_raw, d_raw)
        return o_raw, o_pred, d_raw, d_pred, colors, img_idx, pixel_width

    def training_transform(self, batch):
        o_raw, o_noisy, d_raw, d_noisy, colors, img_idx, pixel_width = batch
        o_pred, d_pred, _, _ = self.camera_extrinsics(img_idx, o_noisy, d_noisy)
        return o_raw, o_pred, d_raw, d_pred, colors, img_idx, pixel_width

    def compute_pose_error(self):
        """Mean distance between the true camera origins and the predicted ones aligned to them
        (one kernel: both fits, the quantile and the error)."""
        origs_raw, origs_pred = self._train_origins()
        return ops.kabsch(origs_pred, origs_raw, True, want_error=True)[3]


class BarfModel(CameraCalibrationModel):
    """reference barf/model_barf.py:12-92."""

    @staticmethod
    def get_sigma_alpha(alpha: th.Tensor, sigma_max: float) -> th.Tensor:
        sigma = sigma_max * 2 ** (-alpha)
        if sigma < 1 / 4:
            return th.tensor([0.], device=alpha.device)
        return sigma

    # -- the training step split for the engine: host-side schedules / device-only loss ------------
    def update_schedules(self, step: int) -> None:
        """Host part of a training step (barf/model_barf.py:36-47): coarse-to-fine alpha of both
        encoders from the fractional epoch and the blur level of the targets, written in place to
        device memory the kernels read — so the device part below can live in a CUDA graph."""
        loop = self._loop()
        epoch = step / len(loop.train_dataloader)
        self.model_radiance.position_encoder.update_alpha(epoch)
        self.model_radiance.direction_encoder.update_alpha(epoch)
        enc = self.model_radiance.position_encoder
        sigma = float(BarfModel.get_sigma_alpha(th.tensor(enc.alpha_value), self.max_gaussian_sigma))
        dm = loop.datamodule
        lo, hi, coef = dm.blur_levels(sigma)
        w = [0.0] * dm.n_sigmas
        if lo == hi:
            w[lo] = 1.0
        else:
            w[lo], w[hi] = coef, 1.0 - coef
        if getattr(self, "_blur_w", None) is None or self._blur_w.numel() != dm.n_sigmas:
            self._blur_w = th.zeros(dm.n_sigmas, device=self.device)
            self._blur_w_host = th.zeros(dm.n_sigmas).pin_memory()
        if getattr(self, "_blur_w_last", None) != w:         # the level changes rarely: no copy otherwise
            self._blur_w_host.copy_(th.tensor(w))
            self._blur_w.copy_(self._blur_w_host, non_blocking=True)
            self._blur_w_last = w
        self._sigma_value = sigma

    def training_loss(self, o_raw, o_noisy, d_raw, d_noisy, colors, img_idx, pixel_width):
        """Device part of BarfModel.training_step (barf/model_barf.py:29-92) on the reference's 7-tuple
        whose colours are the raw blur pyramid (B, n_sigmas, 3): pose transform, blurred targets, render,
        loss, PSNR and the per-step pose error (Kabsch alignment) — no host synchronisation."""
        cam = self.camera_extrinsics
        o_pred, d_pred, _, _ = cam(img_idx, o_noisy, d_noisy)
        blurred = (colors * self._blur_w.view(1, -1, 1)).sum(dim=1) if colors.shape[1] == self._blur_w.numel() \
            else colors[:, 0]
        fine, coarse = self.forward(o_pred, d_pred, pixel_width)
        loss_fine = nn.functional.mse_loss(fine, blurred)
        loss = loss_fine
        logs = {"loss_fine": loss_fine.detach(), "train_psnr": self.psnr_tensor(loss_fine),
                "alpha": self.model_radiance.position_encoder.alpha}
        if self.proposal:
            loss_coarse = nn.functional.mse_loss(coarse, blurred)
            loss = loss_fine + loss_coarse
            logs["train_loss_coarse"] = loss_coarse.detach()
        with th.no_grad():
            logs["pose_error"] = self.compute_pose_error()
        return loss, logs

    def _step_helper(self, batch, batch_idx, purpose: Literal["train", "val"]):
        loop = self._loop()
        if purpose == "train":
            batch = self.training_transform(batch)
            epoch = loop.current_epoch + batch_idx / len(loop.train_dataloader)
            self.model_radiance.position_encoder.update_alpha(epoch)
            self.model_radiance.direction_encoder.update_alpha(epoch)
        elif purpose == "val":
            batch = self.validation_transform(batch)
        enc = self.model_radiance.position_encoder
        # the reference reads alpha back from the device here; update_alpha keeps a host copy
        alpha = th.tensor(enc.alpha_value) if hasattr(enc, "alpha_value") else enc.alpha
        sigma = BarfModel.get_sigma_alpha(alpha, self.max_gaussian_sigma)
        batch = loop.datamodule.get_blurred_pixel_colors(batch, float(sigma))
        _, o_pred, _, d_pred, colors, _, pixel_width = batch
        fine, coarse = self.forward(o_pred, d_pred, pixel_width)
        loss_fine = nn.functional.mse_loss(fine, colors[:, 0])
        log = {f"{purpose}_loss_fine": loss_fine, f"{purpose}_psnr": self.psnr_tensor(loss_fine),
               "alpha": self.model_radiance.position_encoder.alpha, "sigma": sigma}
        loss = loss_fine
        if self.proposal:
            loss_coarse = nn.functional.mse_loss(coarse, colors[:, 0])
            loss = loss_fine + loss_coarse
            log[f"{purpose}_loss_coarse"] = loss_coarse
        if purpose == "train":
            log["pose_error"] = self.compute_pose_error()
        self.log_dict(log)
        return loss


class MipNeRF(NerfInterpolation):
    """reference barf/model_mip.py:17-84: one network used as proposal and radiance model, coarse
    loss weighted 0.1. (At the reference's HEAD the constructor passes `self` twice and raises;
    this is the evident intent.)"""

    def __init__(self, near_sphere_normalized: float, far_sphere_normalized: float, model_radiance: NerfModel,
                 samples_per_ray_radiance: int, uniform_sampling_strategy="stratified_uniform",
                 uniform_sampling_offset_size: float = 0., integration_strategy="middle",
                 samples_per_ray_proposal: int = 0):
        NerfInterpolation.__init__(self, near_sphere_normalized=near_sphere_normalized,
                                   far_sphere_normalized=far_sphere_normalized, model_radiance=model_radiance,
                                   model_proposal=model_radiance if samples_per_ray_proposal > 0 else None,
                                   samples_per_ray_radiance=samples_per_ray_radiance,
                                   uniform_sampling_strategy=uniform_sampling_strategy,
                                   uniform_sampling_offset_size=uniform_sampling_offset_size,
                                   integration_strategy=integration_strategy,
                                   samples_per_ray_proposal=samples_per_ray_proposal)
        self.param_groups = self.model_radiance.param_groups

    def _step_helper(self, batch, batch_idx, purpose: Literal["train", "val"]):
        _, o_pred, _, d_pred, colors, _, pixel_width = batch
        fine, coarse = self.forward(o_pred, d_pred, pixel_width)
        loss = nn.functional.mse_loss(fine, colors[:, 0])
        logs = {f"{purpose}_loss_fine": loss, f"{purpose}_psnr": self.psnr_tensor(loss)}
        if self.proposal:
            loss_coarse = nn.functional.mse_loss(coarse, colors[:, 0])
            loss = loss + loss_coarse * 0.1
            logs[f"{purpose}_loss_coarse"] = loss_coarse
        self.log_dict(logs)
        return self._nan_guard(loss)      # barf/model_mip.py:78-80


class MipBarf(CameraCalibrationModel):
    """reference barf/model_mip.py:87-304: Mip-NeRF integrated encoding + pose refinement with a
    joint schedule for the image blur and the pixel-width (cone) sigma."""

    def __init__(self, model_radiance: NerfModel, samples_per_ray_radiance: int, n_training_images: int,
                 camera_learning_rate_start: float, camera_learning_rate_stop: float,
                 camera_learning_rate_decay_end: int = -1, near_sphere_normalized: float = 2.,
                 far_sphere_normalized: float = 8., uniform_sampling_strategy="stratified_uniform",
                 uniform_sampling_offset_size: float = 0., samples_per_ray_proposal: int = 0,
                 sigma_decay_start_step: int = 0, sigma_decay_end_step: int = 0, start_blur_sigma: float = 0.,
                 start_pixel_width_sigma: float = 0.0):
        CameraCalibrationModel.__init__(
            self, model_radiance=model_radiance,
            model_proposal=model_radiance if samples_per_ray_proposal > 0 else None,
            samples_per_ray_radiance=samples_per_ray_radiance, n_training_images=n_training_images,
            camera_learning_rate_start=camera_learning_rate_start, camera_learning_rate_stop=camera_learning_rate_stop,
            camera_learning_rate_decay_end=camera_learning_rate_decay_end, max_gaussian_sigma=None,
            near_sphere_normalized=near_sphere_normalized, far_sphere_normalized=far_sphere_normalized,
            uniform_sampling_strategy=uniform_sampling_strategy,
            uniform_sampling_offset_size=uniform_sampling_offset_size, integration_strategy="middle",
            samples_per_ray_proposal=samples_per_ray_proposal)
        self.start_blur_sigma = float(start_blur_sigma)
        self.start_pixel_width_sigma = float(start_pixel_width_sigma)
        self.sigma_decay_start_step = sigma_decay_start_step
        self.sigma_decay_end_step = sigma_decay_end_step
        self.sigma_schedule = 1.
        self.param_groups = [g for m in (self.model_radiance, self.camera_extrinsics) for g in m.param_groups]
        self.model_radiance.position_encoder.pixel_width_sigma = self.start_pixel_width_sigma

    def update_sigma_schedule(self, current_step):
        if current_step < self.sigma_decay_start_step:
            s = 1.
        elif self.sigma_decay_start_step <= current_step <= self.sigma_decay_end_step:
            s = (0.25 / max(self.start_blur_sigma, self.start_pixel_width_sigma)) ** (
                (self.sigma_decay_start_step - current_step) / (self.sigma_decay_start_step - self.sigma_decay_end_step))
        else:
            s = 0.
        self.sigma_schedule = s

    @property
    def current_blur_sigma(self):
        sigma = self.sigma_schedule * self.start_blur_sigma
        return 0.0 if sigma < 0.25 else sigma

    @property
    def current_pixel_width_sigma(self):
        sigma = self.sigma_schedule * self.start_pixel_width_sigma
        return 0.0 if sigma < 0.25 else sigma

    def _step_helper(self, batch, batch_idx, purpose: Literal["train", "val"]):
        loop = self._loop()
        if purpose == "train":
            current_step = loop.current_epoch * len(loop.train_dataloader) + batch_idx
            self.update_sigma_schedule(current_step)
            self.model_radiance.position_encoder.pixel_width_sigma = self.current_pixel_width_sigma
            batch = self.training_transform(batch)
        elif purpose == "val":
            batch = self.validation_transform(batch)
        else:
            raise ValueError(f"purpose={purpose} is invalid")
        _, o_pred, _, d_pred, colors, _, pixel_width = loop.datamodule.get_blurred_pixel_colors(
            batch, self.current_blur_sigma)
        fine, coarse = self.forward(o_pred, d_pred, pixel_width)
        loss = nn.functional.mse_loss(fine, colors[:, 0])
        logs = {f"{purpose}_loss_fine": loss, f"{purpose}_psnr": self.psnr_tensor(loss),
                "PE_sigma": self.model_radiance.position_encoder.pixel_width_sigma,
                "blur_sigma": self.current_blur_sigma}
        if self.proposal:
            loss_coarse = nn.functional.mse_loss(coarse, colors[:, 0])
            loss = loss + loss_coarse * 0.1
            logs[f"{purpose}_loss_coarse"] = loss_coarse
        if (purpose == "train" and batch_idx % 100 == 0) or (purpose == "val" and batch_idx == 0):
            logs["pose_error"] = self.compute_pose_error()
        self.log_dict(logs)
        return self._nan_guard(loss)      # barf/model_mip.py:300-302
