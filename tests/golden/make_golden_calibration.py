"""Generates tests/golden/calibration.npz from the UNMODIFIED reference (this container only):
CameraCalibrationModel.kabsch_algorithm with and without outlier rejection, the pose error, and
the BARF / Mip-BARF sigma schedules.  Run: python tests/golden/make_golden_calibration.py"""
import os
import sys
import types

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
sys.path.insert(0, "/root/reference/barf")


def mip_barf_golden(model_mip, g):
    """MipBarf (integrated encoding, ONE network as proposal and radiance model, pose refinement):
    forward of the unmodified reference + the loss of its _step_helper (fine + 0.1 coarse,
    barf/model_mip.py:283-292) + gradients, with the offset uniforms it drew."""
    import model_interpolation_architecture as arch
    import positional_encodings as pe
    th.manual_seed(99)
    ep = pe.IntegratedFourierFeatures(levels=10, include_identity=True, scale=1., distribute_variance=False)
    ed = pe.BarfPositionalEncoding(0, 1, 0, 1, True)
    net = arch.NerfModel(n_hidden=2, hidden_dim=64, delayed_direction=True, delayed_density=False, n_segments=2,
                         position_encoder=ep, direction_encoder=ed, learning_rate_start=5e-4,
                         learning_rate_stop=1e-5, learning_rate_decay_end=1000)
    m = model_mip.MipBarf(model_radiance=net, samples_per_ray_radiance=48, n_training_images=6,
                          camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
                          camera_learning_rate_decay_end=1000, uniform_sampling_strategy="equidistant",
                          uniform_sampling_offset_size=-1., samples_per_ray_proposal=16,
                          sigma_decay_start_step=10, sigma_decay_end_step=100, start_blur_sigma=8.,
                          start_pixel_width_sigma=1.5)
    with th.no_grad():
        m.camera_extrinsics.rotation.copy_(th.randn((6, 3), generator=g) * 0.05)
        m.camera_extrinsics.translation.copy_(th.randn((6, 3), generator=g) * 0.05)
    B = 24
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    target = th.rand((B, 3), generator=g)
    idx = th.randint(0, 6, (B,), generator=g)
    pw = th.full((B,), 1 / 55.0)     # wide pixels: the cone radius matters at this scale
    o2, d2, _, _ = m.camera_extrinsics(idx, o, d)
    th.manual_seed(5)
    fine, coarse = m(o2, d2, pw)
    th.manual_seed(5)
    offset = th.rand((B, 1))
    loss = th.nn.functional.mse_loss(fine, target) + 0.1 * th.nn.functional.mse_loss(coarse, target)
    loss.backward()
    out = {"mip_o": o, "mip_d": d, "mip_target": target, "mip_idx": idx, "mip_pw": pw, "mip_offset": offset,
           "mip_fine": fine, "mip_coarse": coarse, "mip_loss": loss.reshape(1),
           "mip_rotation": m.camera_extrinsics.rotation, "mip_translation": m.camera_extrinsics.translation,
           "mip_d_rotation": m.camera_extrinsics.rotation.grad, "mip_d_translation": m.camera_extrinsics.translation.grad}
    for k, v in net.state_dict().items():
        out["mip_sd." + k] = v
    for k, p in net.named_parameters():
        out["mip_grad." + k] = p.grad
    return out


def main():
    import model_barf
    import model_camera_calibration as mcc
    import model_mip
    g = th.Generator().manual_seed(21)
    out = {}
    fake = types.SimpleNamespace()
    fake.kabsch_algorithm = lambda a, b, remove_outliers=True: mcc.CameraCalibrationModel.kabsch_algorithm(fake, a, b, remove_outliers)
    for tag, n, noise, outliers in (("clean", 100, 0.0, 0), ("noisy", 100, 0.05, 0), ("outliers", 60, 0.02, 5),
                                    ("small", 12, 0.01, 1)):
        A = th.randn((3, 3), generator=g)
        Q, _ = th.linalg.qr(A)
        if th.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        c = float(th.rand(1, generator=g)) + 0.5
        t = th.randn((1, 3), generator=g)
        src = th.randn((n, 3), generator=g) * 3
        dst = (Q @ src.T).T * c + t + noise * th.randn((n, 3), generator=g)
        if outliers:
            dst[:outliers] += 3.0 * th.randn((outliers, 3), generator=g)
        out[f"{tag}_from"], out[f"{tag}_to"] = src, dst
        for ro in (False, True):
            R, tt, cc = fake.kabsch_algorithm(src, dst, ro)
            out[f"{tag}_R_{int(ro)}"], out[f"{tag}_t_{int(ro)}"], out[f"{tag}_c_{int(ro)}"] = R, tt, cc.reshape(1)
        # compute_pose_error: align pred (=from) to raw (=to), error over all points
        R, tt, cc = fake.kabsch_algorithm(src, dst, True)
        aligned = th.matmul(R.unsqueeze(0), src.unsqueeze(2)).squeeze(2) * cc + tt
        out[f"{tag}_err"] = (((dst - aligned) ** 2).sum(dim=1) ** 0.5).mean().reshape(1)
    out["barf_sigma"] = np.array([float(model_barf.BarfModel.get_sigma_alpha(th.tensor(a), 8.0)) for a in (0.0, 2.5, 4.9, 5.1, 9.0)])
    mip = types.SimpleNamespace(sigma_decay_start_step=100, sigma_decay_end_step=1100, start_blur_sigma=8.0,
                                start_pixel_width_sigma=4.0, sigma_schedule=1.0)
    sched = []
    for step in (0, 100, 600, 1100, 1101):
        model_mip.MipBarf.update_sigma_schedule(mip, step)
        sched.append(mip.sigma_schedule)
    out["mip_schedule"] = np.array(sched)
    out.update(mip_barf_golden(model_mip, g))
    arrays = {k: (v.detach().numpy() if isinstance(v, th.Tensor) else np.asarray(v)) for k, v in out.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "calibration.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
