"""CameraExtrinsics: per-image so(3) rotation + translation refinement parameters with the
reference's constructor, parameter names and methods (barf/model_camera_extrinsics.py:7-85);
the rotation exponential, the ray transform and its backward run as CUDA kernels."""
import torch as th
import torch.nn as nn

from . import ops
from .model_interpolation_architecture import NerfBaseModel


class CameraExtrinsics(NerfBaseModel):
    def __init__(self, n_train_images: int, learning_rate_start: float, learning_rate_stop: float,
                 learning_rate_decay_end: int = -1) -> None:
        super().__init__()
        self.size = n_train_images
        self.rotation = nn.Parameter(th.zeros((n_train_images, 3)))      # so(3) Lie algebra
        self.translation = nn.Parameter(th.zeros((n_train_images, 3)))
        self._add_param_group(self.parameters(), learning_rate_start, learning_rate_stop, learning_rate_decay_end)

    @staticmethod
    def so3_to_SO3(so3: th.Tensor) -> th.Tensor:
        """(N,3) (or anything holding 3 numbers per rotation) -> (N,3,3) rotation matrices."""
        return ops.so3_to_SO3(so3)

    def get_rotations(self, img_idx: th.Tensor) -> th.Tensor:
        return ops.so3_to_SO3(self.rotation[img_idx.long()])

    def forward_origins(self, i: th.Tensor, o: th.Tensor):
        t = self.translation[i.long()]          # / MAGIC_NUMBER_THE_SECOND (= 1, barf/magic.py:1)
        return o + t, t

    def forward(self, i: th.Tensor, o: th.Tensor, d: th.Tensor):
        """(o + t_i, R_i d, R_i, t_i)."""
        # `grad_sink` (set by engine.TrainEngine): the backward kernel accumulates the pose gradients
        # straight into the engine's flat gradient buffer
        return ops.pose_forward(self.rotation, self.translation, i, o, d, getattr(self, "grad_sink", None))
