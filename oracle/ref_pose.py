"""Oracle: camera extrinsics (a13).  Test infrastructure only."""
import torch as th


def so3_to_SO3(so3: th.Tensor) -> th.Tensor:
    """reference barf/model_camera_extrinsics.py:22-43."""
    return th.matrix_exp(th.cross(-th.eye(3).view(1, 3, 3), so3.view(-1, 3, 1), dim=1))


def pose_forward(rotation, translation, img_idx, o, d):
    """reference barf/model_camera_extrinsics.py:46-85 (MAGIC_NUMBER_THE_SECOND = 1)."""
    t = translation[img_idx] / 1
    new_o = o + t
    R = so3_to_SO3(rotation)[img_idx]
    new_d = th.matmul(R, d.unsqueeze(-1)).squeeze(-1)
    return new_o, new_d, R, t


def validation_transform_rays(o, d, R, t, c):
    """reference barf/model_camera_calibration.py:186-193."""
    return th.matmul(R, o.unsqueeze(-1)).squeeze(-1) * c + t, th.matmul(R, d.unsqueeze(-1)).squeeze(-1)
