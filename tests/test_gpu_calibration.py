"""Pose alignment kernel (C ABI `nerfb200_kabsch`) and the BarfModel / MipBarf module surface on the
GPU against the outputs of the unmodified reference (tests/golden/calibration.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch as th

from oracle import ref_calibration

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calibration.npz")


def _g():
    z = np.load(G)
    return {k: th.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("tag", ["clean", "noisy", "outliers", "small"])
@pytest.mark.parametrize("remove_outliers", [False, True])
def test_kabsch_kernel_matches_reference(cuda, tag, remove_outliers):
    from nerf_experiments_b200 import ops
    g = _g()
    R, t, c, err = ops.kabsch(g[f"{tag}_from"].to(cuda), g[f"{tag}_to"].to(cuda), remove_outliers, want_error=True)
    k = int(remove_outliers)
    assert R.shape == (3, 3) and t.shape == (1, 3) and c.shape == (1,)
    assert (R.cpu() - g[f"{tag}_R_{k}"]).abs().max() < 2e-6
    assert (t.cpu() - g[f"{tag}_t_{k}"]).abs().max() < 2e-5
    assert (c.cpu() - g[f"{tag}_c_{k}"]).abs().max() < 2e-6
    if remove_outliers:
        assert abs(float(err) - float(g[f"{tag}_err"])) < 1e-5 * max(1.0, float(g[f"{tag}_err"]))


def test_kabsch_kernel_sizes_and_reflection(cuda):
    """1..2048 points against the oracle; a reflected target must still give a proper rotation."""
    from nerf_experiments_b200 import ops
    gen = th.Generator().manual_seed(4)
    for n in (4, 33, 500, 2048):
        src = th.randn((n, 3), generator=gen) * 2
        Q, _ = th.linalg.qr(th.randn((3, 3), generator=gen))
        dst = (Q @ src.T).T * 0.8 + th.randn((1, 3), generator=gen) + 0.01 * th.randn((n, 3), generator=gen)
        for ro in (False, True):
            R, t, c = ops.kabsch(src.to(cuda), dst.to(cuda), ro)
            Rr, tr, cr = ref_calibration.kabsch(src, dst, ro)
            assert (R.cpu() - Rr).abs().max() < 1e-5 and (t.cpu() - tr).abs().max() < 1e-4
            assert abs(float(c) - float(cr)) < 1e-5
            assert abs(float(th.linalg.det(R.cpu())) - 1.0) < 1e-5
    with pytest.raises(RuntimeError):
        ops.kabsch(th.zeros((4000, 3), device=cuda), th.zeros((4000, 3), device=cuda))
    with pytest.raises(ValueError):
        ops.kabsch(th.zeros((4, 2), device=cuda), th.zeros((4, 2), device=cuda))


def _mip_model(cuda, g, fixed_offset=True):
    from nerf_experiments_b200 import ops
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.model_interpolation_architecture import NerfModel
    from nerf_experiments_b200.model_mip import MipBarf
    ep = pe.IntegratedFourierFeatures(levels=10, include_identity=True, scale=1., distribute_variance=False)
    ed = pe.BarfPositionalEncoding(0, 1, 0, 1, True)
    net = NerfModel(n_hidden=2, hidden_dim=64, delayed_direction=True, delayed_density=False, n_segments=2,
                    position_encoder=ep, direction_encoder=ed, learning_rate_start=5e-4, learning_rate_stop=1e-5,
                    learning_rate_decay_end=1000)
    sd = {k[len("mip_sd."):]: v for k, v in g.items() if k.startswith("mip_sd.")}
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    m = MipBarf(model_radiance=net, samples_per_ray_radiance=48, n_training_images=6, camera_learning_rate_start=1e-3,
                camera_learning_rate_stop=1e-5, camera_learning_rate_decay_end=1000,
                uniform_sampling_strategy="equidistant", uniform_sampling_offset_size=-1., samples_per_ray_proposal=16,
                sigma_decay_start_step=10, sigma_decay_end_step=100, start_blur_sigma=8., start_pixel_width_sigma=1.5)
    m.camera_extrinsics.load_state_dict({"rotation": g["mip_rotation"], "translation": g["mip_translation"]})
    m = m.to(cuda)
    if fixed_offset:     # the offset uniforms the reference run drew
        m._sample_t_stratified_uniform = lambda B, S, strat, off: ops.sample_uniform(
            2.0, 8.0, B, S, cuda, None, g["mip_offset"].to(cuda), off)
    return m, net


def test_mip_barf_forward_backward_against_reference(cuda):
    """C3 path: integrated encoding with the cone-sigma term, ONE network as proposal and radiance
    model (its gradients accumulate over both passes), pose refinement in front."""
    g = _g()
    m, net = _mip_model(cuda, g)
    o2, d2, _, _ = m.camera_extrinsics(g["mip_idx"].to(cuda), g["mip_o"].to(cuda), g["mip_d"].to(cuda))
    fine, coarse = m(o2, d2, g["mip_pw"].to(cuda))
    assert (fine.detach().cpu() - g["mip_fine"]).abs().max() < 1e-2        # north_star: bf16-MLP rgb 1e-2 abs
    assert (coarse.detach().cpu() - g["mip_coarse"]).abs().max() < 1e-2
    target = g["mip_target"].to(cuda)
    loss = th.nn.functional.mse_loss(fine, target) + 0.1 * th.nn.functional.mse_loss(coarse, target)
    assert abs(float(loss.detach()) - float(g["mip_loss"])) < 2e-3
    loss.backward()

    def rel(a, b):
        return float((a.cpu() - b).norm() / (b.norm() + 1e-12))
    worst = 0.0
    for k, p in net.named_parameters():
        r = rel(p.grad, g["mip_grad." + k])
        worst = max(worst, r)
        assert r < 0.25, (k, r)                                              # bf16 operands vs fp32
    assert rel(m.camera_extrinsics.translation.grad, g["mip_d_translation"]) < 0.25
    assert rel(m.camera_extrinsics.rotation.grad, g["mip_d_rotation"]) < 0.25


def _scene_batcher(cuda, n_images=6, hw=32, sigmas=(8.0, 4.0, 2.0, 1.0, 0.0)):
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher
    gen = th.Generator().manual_seed(2)
    images = th.rand((n_images, hw, hw, len(sigmas), 3), generator=gen)
    c2w = th.eye(4).repeat(n_images, 1, 1)
    for i in range(n_images):
        a = 2 * np.pi * i / n_images
        pos = th.tensor([4 * np.cos(a), 4 * np.sin(a), 1.0], dtype=th.float32)
        z = th.nn.functional.normalize(pos, dim=0)
        x = th.nn.functional.normalize(th.linalg.cross(th.tensor([0., 0., 1.]), z), dim=0)
        y = th.linalg.cross(z, x)
        c2w[i, :3, 0], c2w[i, :3, 1], c2w[i, :3, 2], c2w[i, :3, 3] = x, y, z, pos
    noisy = c2w.clone()
    noisy[:, :3, 3] += 0.1 * th.randn((n_images, 3), generator=gen)
    return GpuRayBatcher(images, c2w, 40.0, noisy, list(sigmas), None, cuda)


def test_barf_and_mip_step_helpers_run_on_the_batcher(cuda):
    """training_step / validation_step of BarfModel and MipBarf on batches of the GPU ray batcher:
    losses finite, pose error == oracle, logged keys as the reference's, gradients reach networks
    and poses, no Lightning needed."""
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.model_barf import BarfModel
    from nerf_experiments_b200.model_camera_calibration import LoopState
    from nerf_experiments_b200.model_interpolation_architecture import NerfModel
    b = _scene_batcher(cuda)
    th.manual_seed(0)
    ep = pe.BarfPositionalEncoding(10, 0.0, 0.2, 0.8, True, 1.0)
    ed = pe.BarfPositionalEncoding(4, 0.0, 0.2, 0.8, True, 1.0)
    net = NerfModel(2, 64, True, False, 2, ep, ed, 5e-4, 1e-5, 1000)
    model = BarfModel(n_training_images=6, camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
                      camera_learning_rate_decay_end=1000, near_sphere_normalized=2., far_sphere_normalized=8.,
                      model_radiance=net, samples_per_ray_radiance=32, max_gaussian_sigma=8.0,
                      uniform_sampling_strategy="stratified_uniform", uniform_sampling_offset_size=-1.).to(cuda)
    with th.no_grad():
        model.camera_extrinsics.translation.copy_(0.02 * th.randn(6, 3))
    model.loop = LoopState(b, n_batches=10)
    idx = b.epoch_permutation()[:256]
    loss = model.training_step(b.batch(idx), 5)          # epoch 0.5 -> alpha in the middle of its ramp
    assert th.isfinite(loss)
    assert ep.alpha_value == pytest.approx(10 * (0.5 - 0.2) / 0.6, rel=1e-5)
    logged = model._logged
    assert {"train_loss_fine", "train_psnr", "alpha", "sigma", "pose_error"} <= set(logged)
    want = ref_calibration.pose_error(b.camera_origins.cpu(), (b.camera_origins_noisy + model.camera_extrinsics.translation.detach()).cpu())
    assert abs(float(logged["pose_error"]) - float(want)) < 1e-5
    loss.backward()
    assert all(p.grad is not None and th.isfinite(p.grad).all() for p in net.parameters())
    assert model.camera_extrinsics.rotation.grad.abs().sum() > 0
    vloss = model.validation_step(b.batch(idx), 0)
    assert th.isfinite(vloss) and "val_psnr" in model._logged

    # validation transform == oracle Kabsch applied to the rays
    o = th.randn((20, 3), device=cuda)
    d = th.nn.functional.normalize(th.randn((20, 3), device=cuda), dim=1)
    o_m, d_m, params = model.validation_transform_rays(o, d)
    pred = (b.camera_origins_noisy + model.camera_extrinsics.translation.detach()).cpu()
    R, t, c = ref_calibration.kabsch(b.camera_origins.cpu(), pred, True)
    assert (o_m.cpu() - ((R @ o.cpu().T).T * c + t)).abs().max() < 1e-4
    assert (d_m.cpu() - (R @ d.cpu().T).T).abs().max() < 1e-5

    g = _g()
    mip, mnet = _mip_model(cuda, g, fixed_offset=False)
    mip.loop = LoopState(b, n_batches=10)
    mip.loop.current_epoch = 3                             # step 35: inside the sigma decay
    loss = mip.training_step(b.batch(idx), 5)
    assert th.isfinite(loss)
    s = ref_calibration.mip_sigma_schedule(35, 10, 100, 8.0, 1.5)
    assert mip.sigma_schedule == pytest.approx(s)
    assert mnet.position_encoder.pixel_width_sigma == pytest.approx(1.5 * s if 1.5 * s >= 0.25 else 0.0)
    assert "pose_error" not in mip._logged                 # only every 100 batches
    loss.backward()
    assert all(p.grad is not None and th.isfinite(p.grad).all() for p in mnet.parameters())
