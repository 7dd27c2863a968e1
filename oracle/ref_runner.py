"""Drives the UNMODIFIED reference modules (oracle/_ref/barf, vendored by oracle/make_ref.py, or
/root/reference/barf when present) on the CPU: `BarfModel.training_step` + `loss.backward()` on the
bench workload, with the minimal fake Lightning trainer the step reads.

BASELINE INFRASTRUCTURE ONLY — used by `bench.py --impl reference` / `cpu_baseline` and by tests.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(HERE, "_stubs")


def reference_dir():
    """Directory holding the unmodified barf/*.py, or None."""
    for cand in (os.path.join(HERE, "_ref", "barf"),
                 os.path.join(os.environ.get("NERF_REFERENCE_ROOT", "/root/reference"), "barf")):
        if os.path.isfile(os.path.join(cand, "model_barf.py")):
            return cand
    return None


def load_reference():
    """Imports the reference's flat modules (the way its own scripts do, CWD = barf/)."""
    d = reference_dir()
    if d is None:
        raise RuntimeError("no reference sources: run oracle/make_ref.py where /root/reference exists")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    for p in (d, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, _STUBS)
    sys.path.insert(0, d)
    import importlib
    mods = {n: importlib.import_module(n) for n in
            ("positional_encodings", "model_interpolation_architecture", "model_camera_extrinsics",
             "model_interpolation", "data_module", "model_camera_calibration", "model_barf")}
    return types.SimpleNamespace(**mods, directory=d)


def build_barf(ref, n_images: int, samples: int, near: float, far: float, n_batches: int,
               camera_origins, camera_origins_noisy, blur_sigmas, max_sigma: float, seed: int = 1337,
               alpha_epochs=(0.5, 2.5)):
    """The reference's BarfModel of barf/run_barf.py:151-196 (PE 10/4 + identity, 4x256x2 segments,
    equidistant sampling with offset -1) wired to a fake trainer."""
    import torch as th
    th.manual_seed(seed)
    pe, arch = ref.positional_encodings, ref.model_interpolation_architecture
    ep = pe.BarfPositionalEncoding(levels=10, alpha_start=0, alpha_increase_start_epoch=alpha_epochs[0],
                                   alpha_increase_end_epoch=alpha_epochs[1], include_identity=True, scale=1.)
    ed = pe.BarfPositionalEncoding(levels=4, alpha_start=0, alpha_increase_start_epoch=alpha_epochs[0],
                                   alpha_increase_end_epoch=alpha_epochs[1], include_identity=True, scale=1.)
    net = arch.NerfModel(n_hidden=4, hidden_dim=256, delayed_direction=True, delayed_density=False, n_segments=2,
                         position_encoder=ep, direction_encoder=ed, learning_rate_start=5e-4,
                         learning_rate_stop=1e-5, learning_rate_decay_end=200000)
    model = ref.model_barf.BarfModel(
        n_training_images=n_images, camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
        camera_learning_rate_decay_end=200000, near_sphere_normalized=near, far_sphere_normalized=far,
        samples_per_ray_radiance=samples, samples_per_ray_proposal=0, model_radiance=net,
        uniform_sampling_strategy="equidistant", uniform_sampling_offset_size=-1., max_gaussian_sigma=max_sigma)
    dm = object.__new__(ref.data_module.ImagePoseDataModule)     # get_blurred_pixel_colors only reads this field
    dm.gaussian_blur_sigmas = list(blur_sigmas)
    dm.dataset_train = types.SimpleNamespace(camera_origins=camera_origins, camera_origins_noisy=camera_origins_noisy,
                                             index_to_index={i: i for i in range(n_images)}, n_images=n_images)
    model.trainer = types.SimpleNamespace(train_dataloader=range(n_batches), current_epoch=0, global_step=0,
                                          datamodule=dm, logger=types.SimpleNamespace(experiment=types.SimpleNamespace(dir="/tmp")))
    return model


def training_step(model, batch, batch_idx: int):
    """loss of the reference's `training_step` followed by `backward()` (what Lightning's automatic
    optimisation runs before `optimizer.step()`)."""
    for p in model.parameters():
        p.grad = None
    loss = model.training_step(batch, batch_idx)
    loss.backward()
    return loss
