"""BarfModel — module surface of reference barf/model_barf.py:12-92: coarse-to-fine positional-encoding
schedule, blurred targets and per-step pose error around CameraCalibrationModel, split for the training
engine into the host-side schedules (`update_schedules`) and the capturable device part
(`training_loss`)."""
from typing import Literal

import torch as th
import torch.nn as nn

from .model_camera_calibration import CameraCalibrationModel, LoopState  # noqa: F401


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
        # The per-step pose error (one single-block kernel, ~0.1 ms of latency) depends on the pose
        # parameters only: it is forked onto a side stream (also inside a captured graph) and joins at the
        # end, so it runs in the shadow of the field kernels instead of in front of them.
        main = th.cuda.current_stream(o_noisy.device)
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != o_noisy.device:
            self._side_stream = th.cuda.Stream(device=o_noisy.device)
        side = self._side_stream
        side.wait_stream(main)
        with th.cuda.stream(side), th.no_grad():
            pose_error = self.compute_pose_error()
            pose_error.record_stream(main)
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
        main.wait_stream(side)
        logs["pose_error"] = pose_error
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
