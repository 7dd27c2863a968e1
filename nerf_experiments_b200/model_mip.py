"""MipNeRF / MipBarf — module surface of reference barf/model_mip.py:17-304: the Mip-NeRF integrated
encoding with one network used as proposal and radiance model, and its pose-refining variant with the
joint schedule of image blur and pixel-width (cone) sigma."""
from typing import Literal

import torch as th
import torch.nn as nn

from .model_camera_calibration import CameraCalibrationModel, LoopState  # noqa: F401
from .model_interpolation import NerfInterpolation
from .model_interpolation_architecture import NerfModel


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
