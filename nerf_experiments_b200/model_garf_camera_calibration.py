"""GARF with camera-pose refinement — module surface of reference garf/model_camera_calibration.py:14-479:
`CameraCalibrationModel(GarfModel)` over 6-tuple batches (o_raw, o_noisy, d_raw, d_noisy, colors,
img_idx), `training_transform` / `validation_transform`, the Kabsch alignment helpers and manual
optimisation with THREE Adam optimisers + ExponentialLR schedulers (proposal, radiance, camera).

The camera path is the same kernels as in barf/: `CameraExtrinsics.forward` (pose_fwd / pose_bwd) in front of
the fused GARF field, whose backward returns d(ray origin) / d(ray direction) in fp32; `engine_for(model)`
runs the step with all five parameter groups in one flat buffer and one fused Adam."""
from typing import Optional

import torch as th

from . import ops
from .model_camera_extrinsics import CameraExtrinsics
from .model_garf import GarfModel


class CameraCalibrationModel(GarfModel):
    def __init__(self, n_training_images: int, camera_learning_rate_start: float,
                 camera_learning_rate_stop: float, camera_learning_rate_decay_end: int,
                 pose_error_logging_period: int = 10, *inner_model_args, **inner_model_kwargs):
        super().__init__(*inner_model_args, **inner_model_kwargs)
        # garf/model_camera_extrinsics.py:4 takes the image count only; the learning rates of the pose group
        # live on this module (garf/model_camera_calibration.py:31-35)
        self.camera_extrinsics = CameraExtrinsics(n_training_images, camera_learning_rate_start,
                                                  camera_learning_rate_stop, camera_learning_rate_decay_end)
        self.camera_learning_rate = camera_learning_rate_start
        self.camera_learning_rate_start = camera_learning_rate_start
        self.camera_learning_rate_stop = camera_learning_rate_stop
        self.camera_learning_rate_decay_end = camera_learning_rate_decay_end
        self.pose_error_logging_period = pose_error_logging_period

    # -- alignment (one kernel launch: csrc/kabsch.cu) ---------------------------------------------
    def kabsch_algorithm(self, point_cloud_from: th.Tensor, point_cloud_to: th.Tensor, remove_outliers: bool = True):
        return ops.kabsch(point_cloud_from, point_cloud_to, remove_outliers)

    def validation_transform_rays(self, origs_val, dirs_val, post_transform_params=None):
        if post_transform_params is None:
            raise RuntimeError("validation_transform_rays: pass the Kabsch parameters (R, t, c) of the training origins")
        R, t, c = post_transform_params
        origs_model = th.matmul(R, origs_val.unsqueeze(-1)).squeeze(-1) * c + t
        dirs_model = th.matmul(R, dirs_val.unsqueeze(-1)).squeeze(-1)
        return origs_model, dirs_model, post_transform_params

    # -- batch transformations ---------------------------------------------------------------------
    def training_transform(self, batch):
        o_raw, o_noisy, d_raw, d_noisy, colors, img_idx = batch
        o_pred, d_pred, _, _ = self.camera_extrinsics(img_idx, o_noisy, d_noisy)
        return o_raw, o_pred, d_raw, d_pred, colors, img_idx

    def validation_transform(self, batch, post_transform_params=None):
        o_raw, _, d_raw, _, colors, img_idx = batch
        o_pred, d_pred, _ = self.validation_transform_rays(o_raw, d_raw, post_transform_params)
        return o_raw, o_pred, d_raw, d_pred, colors, img_idx

    def _forward_loss(self, batch, u_rays=None):
        if len(batch) == 3:
            return super()._forward_loss(batch, u_rays)
        _, o_pred, _, d_pred, colors, _ = batch
        return super()._forward_loss((o_pred, d_pred, colors), u_rays)

    # -- steps -----------------------------------------------------------------------------------------
    def _opt_and_sched(self):
        if getattr(self, "trainer", None) is not None:      # real Lightning
            return self.optimizers(use_pl_optimizer=False), self.lr_schedulers()
        if not hasattr(self, "_camera_optimizer"):
            self.configure_optimizers()
        return ([self._proposal_optimizer, self._radiance_optimizer, self._camera_optimizer],
                [self._proposal_learning_rate_scheduler, self._radiance_learning_rate_scheduler,
                 self._camera_learning_rate_scheduler])

    def training_step(self, batch, batch_idx: int, u_rays=None):
        return super().training_step(self.training_transform(batch), batch_idx, u_rays)

    def validation_step(self, batch, batch_idx: int, post_transform_params=None):
        return super().validation_step(self.validation_transform(batch, post_transform_params), batch_idx)

    def configure_optimizers(self):
        optimizers, schedulers = super().configure_optimizers()
        lr = self.camera_learning_rate_start
        self._camera_optimizer = th.optim.Adam([{"params": self.camera_extrinsics.parameters(), "lr": lr, "initial_lr": lr}])
        self._camera_learning_rate_scheduler = th.optim.lr_scheduler.ExponentialLR(
            self._camera_optimizer,
            gamma=self._calculate_decay_factor(self.camera_learning_rate_start, self.camera_learning_rate_stop,
                                               self.camera_learning_rate_decay_end),
            last_epoch=-self.camera_learning_rate_decay_end - 1)
        return optimizers + [self._camera_optimizer], schedulers + [self._camera_learning_rate_scheduler]

    # -- engine surface ----------------------------------------------------------------------------
    @property
    def param_groups(self):
        if getattr(self, "_param_groups_cam", None) is None:
            self._param_groups_cam = list(super().param_groups) + [{
                "parameters": list(self.camera_extrinsics.parameters()),
                "learning_rate_start": self.camera_learning_rate_start, "learning_rate_stop": self.camera_learning_rate_stop,
                "learning_rate_decay_end": self.camera_learning_rate_decay_end, "weight_decay": 0.0,
                "schedule": "exponential",
                "gamma": self._calculate_decay_factor(self.camera_learning_rate_start, self.camera_learning_rate_stop,
                                                      self.camera_learning_rate_decay_end)}]
        return self._param_groups_cam

    def training_loss(self, o_raw, o_noisy, d_raw, d_noisy, colors, img_idx, u_prop=None, u_rad=None):
        """Device part of training_step on the reference's 6-tuple (no host synchronisation; capturable)."""
        o_pred, d_pred, _, _ = self.camera_extrinsics(img_idx, o_noisy, d_noisy)
        return super().training_loss(o_pred, d_pred, colors, u_prop, u_rad)
