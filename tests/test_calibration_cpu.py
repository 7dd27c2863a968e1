"""Oracle of the pose alignment / pose error / blur schedules against the outputs of the
unmodified reference (tests/golden/calibration.npz), and the host-side schedule logic of the
BarfModel / MipBarf module surface (no CUDA needed)."""
import os
import types

import numpy as np
import pytest
import torch as th

from oracle import ref_calibration

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calibration.npz")
CASES = ("clean", "noisy", "outliers", "small")


def _g():
    z = np.load(G)
    return {k: th.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("remove_outliers", [False, True])
def test_oracle_kabsch_matches_reference(tag, remove_outliers):
    g = _g()
    R, t, c = ref_calibration.kabsch(g[f"{tag}_from"], g[f"{tag}_to"], remove_outliers)
    k = int(remove_outliers)
    assert th.allclose(R, g[f"{tag}_R_{k}"], atol=1e-6)
    assert th.allclose(t, g[f"{tag}_t_{k}"], atol=1e-5)
    assert th.allclose(c.reshape(1), g[f"{tag}_c_{k}"], atol=1e-6)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_pose_error_matches_reference(tag):
    g = _g()
    # compute_pose_error aligns the predicted origins (from) to the raw ones (to)
    err = ref_calibration.pose_error(g[f"{tag}_to"], g[f"{tag}_from"])
    assert th.allclose(err.reshape(1), g[f"{tag}_err"], atol=1e-6)


def test_kabsch_recovers_a_known_similarity():
    gen = th.Generator().manual_seed(0)
    Q, _ = th.linalg.qr(th.randn((3, 3), generator=gen))
    if th.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    src = th.randn((50, 3), generator=gen)
    dst = (Q @ src.T).T * 1.7 + th.tensor([[0.3, -2.0, 1.0]])
    R, t, c = ref_calibration.kabsch(src, dst, remove_outliers=False)
    assert th.allclose(R, Q, atol=1e-5) and abs(float(c) - 1.7) < 1e-5
    assert th.allclose(t, th.tensor([[0.3, -2.0, 1.0]]), atol=1e-5)


def test_sigma_schedules_match_reference():
    from nerf_experiments_b200.model_camera_calibration import BarfModel, MipBarf
    g = _g()
    alphas = (0.0, 2.5, 4.9, 5.1, 9.0)
    ours = [float(BarfModel.get_sigma_alpha(th.tensor(a), 8.0)) for a in alphas]
    assert np.allclose(ours, g["barf_sigma"].numpy(), rtol=1e-6)
    assert np.allclose([ref_calibration.barf_sigma(a, 8.0) for a in alphas], g["barf_sigma"].numpy(), rtol=1e-6)
    mip = types.SimpleNamespace(sigma_decay_start_step=100, sigma_decay_end_step=1100, start_blur_sigma=8.0,
                                start_pixel_width_sigma=4.0, sigma_schedule=1.0)
    steps = (0, 100, 600, 1100, 1101)
    sched = []
    for s in steps:
        MipBarf.update_sigma_schedule(mip, s)
        sched.append(mip.sigma_schedule)
    assert np.allclose(sched, g["mip_schedule"].numpy(), rtol=1e-12)
    assert np.allclose([ref_calibration.mip_sigma_schedule(s, 100, 1100, 8.0, 4.0) for s in steps],
                       g["mip_schedule"].numpy(), rtol=1e-12)
    # SURVEY KAT: at the end of the decay the larger sigma has reached 1/4, after it both are 0
    MipBarf.update_sigma_schedule(mip, 1100)
    assert abs(mip.sigma_schedule * 8.0 - 0.25) < 1e-12
    assert MipBarf.current_blur_sigma.fget(mip) == pytest.approx(0.25)
    assert MipBarf.current_pixel_width_sigma.fget(mip) == 0.0     # 4 * 1/32 < 1/4


def test_module_surface_of_the_calibration_models():
    """Constructor arguments, attributes and param_groups as the reference's (CPU construction)."""
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.model_barf import BarfModel
    from nerf_experiments_b200.model_interpolation_architecture import NerfModel
    from nerf_experiments_b200.model_mip import MipBarf
    ep = pe.IntegratedFourierFeatures(levels=10, include_identity=True, scale=1., distribute_variance=False)
    ed = pe.BarfPositionalEncoding(0, 1, 0, 1, True)
    net = NerfModel(2, 64, True, False, 2, ep, ed, 5e-4, 1e-5, 1000)
    m = MipBarf(model_radiance=net, samples_per_ray_radiance=48, n_training_images=6, camera_learning_rate_start=1e-3,
                camera_learning_rate_stop=1e-5, camera_learning_rate_decay_end=1000, samples_per_ray_proposal=16,
                sigma_decay_start_step=10, sigma_decay_end_step=100, start_blur_sigma=8., start_pixel_width_sigma=1.5)
    assert m.model_proposal is m.model_radiance and m.proposal
    assert len(m.param_groups) == 2 and ep.pixel_width_sigma == 1.5
    assert set(m.camera_extrinsics.state_dict()) == {"rotation", "translation"}
    assert m.camera_extrinsics.rotation.shape == (6, 3)
    eb = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, True, 1.0)
    net2 = NerfModel(2, 64, True, False, 2, eb, pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, True, 1.0))
    b = BarfModel(n_training_images=4, camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
                  camera_learning_rate_decay_end=100, near_sphere_normalized=2., far_sphere_normalized=8.,
                  model_radiance=net2, samples_per_ray_radiance=32, max_gaussian_sigma=8.0)
    assert len(b.param_groups) == 2 and b.max_gaussian_sigma == 8.0
    opt = b.configure_optimizers()
    assert opt["optimizer"].defaults["eps"] == 1e-5 and opt["lr_scheduler"]["interval"] == "step"
