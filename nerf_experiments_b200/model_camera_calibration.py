"""CameraCalibrationModel — module surface of reference barf/model_camera_calibration.py: the step
glue (a14) around the fused render path shared by BarfModel (model_barf.py) and MipBarf (model_mip.py). Pose alignment and the pose error run in one
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


def __getattr__(name):
    """`BarfModel`, `MipBarf`, `MipNeRF` live in model_barf.py / model_mip.py as in the reference; older
    imports from this module keep working."""
    if name == "BarfModel":
        from .model_barf import BarfModel
        return BarfModel
    if name in ("MipBarf", "MipNeRF"):
        from . import model_mip
        return getattr(model_mip, name)
    raise AttributeError(name)
