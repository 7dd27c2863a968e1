"""Oracle: pose alignment and pose error in plain PyTorch on the CPU,
restating CameraCalibrationModel.kabsch_algorithm / compute_pose_error (reference
barf/model_camera_calibration.py:69-156, :340-346), and the BARF / Mip-BARF blur schedules
(barf/model_barf.py:15-24, barf/model_mip.py:168-225).  Test infrastructure only.  Pinned by
tests/golden/calibration.npz (outputs of the unmodified reference methods)."""
import torch as th


def kabsch(point_cloud_from: th.Tensor, point_cloud_to: th.Tensor, remove_outliers: bool = True):
    def align_rotation(P, Q):
        H = P.T @ Q
        U, S, Vh = th.linalg.svd(H.float())
        d = th.linalg.det((Vh.T @ U.T).float())
        K = th.eye(3)
        K[-1, -1] = d
        return Vh.T @ K @ U.T

    mean_from = point_cloud_from.mean(dim=0, keepdim=True)
    mean_to = point_cloud_to.mean(dim=0, keepdim=True)
    cf, ct = point_cloud_from - mean_from, point_cloud_to - mean_to
    c = th.sqrt((ct ** 2).sum()) / th.sqrt((cf ** 2).sum())
    R = align_rotation(cf, ct)
    t = mean_to - (th.matmul(R, mean_from.T) * c).T
    if remove_outliers:
        hat = th.matmul(R, point_cloud_from.unsqueeze(-1)).squeeze(-1) * c + t
        dist = th.linalg.norm(hat - point_cloud_to, dim=1)
        keep = dist < th.quantile(dist, 0.9)
        return kabsch(point_cloud_from[keep], point_cloud_to[keep], remove_outliers=False)
    return R, t, c


def pose_error(origs_raw: th.Tensor, origs_pred: th.Tensor, remove_outliers: bool = True):
    R, t, c = kabsch(origs_pred, origs_raw, remove_outliers)
    aligned = th.matmul(R.unsqueeze(0), origs_pred.unsqueeze(2)).squeeze(2) * c + t
    return (((origs_raw - aligned) ** 2).sum(dim=1) ** 0.5).mean()


def barf_sigma(alpha: float, sigma_max: float) -> float:
    """reference barf/model_barf.py:15-24."""
    sigma = sigma_max * 2 ** (-alpha)
    return 0.0 if sigma < 0.25 else sigma


def mip_sigma_schedule(step: int, start: int, end: int, blur0: float, pw0: float) -> float:
    """reference barf/model_mip.py:168-198."""
    if step < start:
        return 1.0
    if start <= step <= end:
        return (0.25 / max(blur0, pw0)) ** ((start - step) / (start - end))
    return 0.0
