"""Converged BARF parity run (BASELINE.json north_star: end-of-training PSNR within 0.1 dB; VERDICT r1
item 7: plus Kabsch-aligned pose error within 5 %): the BARF configuration of barf/run_barf.py:44-59,
151-196 (coarse-to-fine mask, blurred targets, pose noise 0.15 with pose refinement, equidistant
sampling with offset -1, lr 5e-4 -> 1e-5 / 1e-3 -> 1e-5) on the synthetic 400 x 400 SDF scene, trained
twice from the same weights on the same ray batches with the same sampling uniforms:

  arm "b200"       this repo: BarfModel through engine.TrainEngine (captured CUDA graph, fused bf16
                   tcgen05 field kernels, fused Adam with the device-side schedule);
  arm "reference"  the UNMODIFIED reference modules (oracle/_ref/barf/*.py: BarfModel.training_step,
                   fp32 PyTorch with TF32 matmuls as barf/run_barf.py:101 sets them) on the same GPU,
                   with torch.optim.Adam(eps=1e-5) and SchedulerLeNice's closed form (the reference's
                   own scheduler class does not construct under torch >= 2.2: `verbose=`).

Both are then evaluated the way the reference evaluates: held-out views rendered from their TRUE poses
mapped into the model's frame by the Kabsch alignment of the training origins
(validation_transform_rays, barf/model_camera_calibration.py:159-193), and compute_pose_error.
The oracle / reference is used here as the checker of a test run, never as a product path.

usage: python scripts/barf_parity.py [--steps 20000] [--rays 1024] [--size 400] [--images 40]
prints one JSON line (kept under profiles/)."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th

NEAR, FAR, SAMPLES = 2.0, 8.0, 128
BLUR = [4.0, 1.41, 0.5, 0.0]


def psnr(mse: float) -> float:
    return -10.0 * math.log10(max(mse, 1e-12))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--size", type=int, default=400)
    ap.add_argument("--images", type=int, default=40)
    ap.add_argument("--val-images", type=int, default=4)
    ap.add_argument("--noise", type=float, default=0.15)
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--seed", type=int, default=7, help="batch order and sampling-offset stream (same on both arms)")
    args = ap.parse_args()
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200 import scene
    from nerf_experiments_b200.engine import TrainEngine
    from nerf_experiments_b200.model_camera_calibration import BarfModel, LoopState
    from nerf_experiments_b200.model_interpolation import le_nice_lr, log_decay_factor
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher

    dev = th.device("cuda:0")
    th.set_float32_matmul_precision("high")          # barf/run_barf.py:101
    n_all = args.images + args.val_images
    sc = scene.make_scene(n_all, args.size, args.size, dev, rotation_noise=args.noise, translation_noise=args.noise,
                          blur_sigmas=BLUR)
    full = sc.batcher
    # training views: the first `images`; held-out views: the rest (their noisy poses are never used)
    n_tr, H, W = args.images, args.size, args.size
    train = GpuRayBatcher(full.images[:n_tr], sc.c2w[:n_tr], sc.focal, sc.c2w_noisy[:n_tr], BLUR, device=dev)
    n_batches = len(train) // args.rays
    steps, decay_end = args.steps, args.steps
    a0, a1 = 0.1 * steps / n_batches, 0.5 * steps / n_batches      # c2f ramp over steps 10 % .. 50 % (run_barf: 20k..100k of 200k)
    g = th.Generator(device=dev).manual_seed(args.seed)
    idx = th.randint(0, len(train), (steps, args.rays), device=dev, generator=g)

    def build_ours():
        th.manual_seed(1337)
        ep = pe.BarfPositionalEncoding(10, 0.0, a0, a1, True, 1.0)
        ed = pe.BarfPositionalEncoding(4, 0.0, a0, a1, True, 1.0)
        net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, decay_end)
        m = BarfModel(n_training_images=n_tr, camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
                      camera_learning_rate_decay_end=decay_end, near_sphere_normalized=NEAR, far_sphere_normalized=FAR,
                      model_radiance=net, samples_per_ray_radiance=SAMPLES, samples_per_ray_proposal=0,
                      max_gaussian_sigma=BLUR[0], uniform_sampling_strategy="equidistant",
                      uniform_sampling_offset_size=-1.0).to(dev)
        m.loop = LoopState(train, n_batches)
        return m

    val_rays = []
    for k in range(n_tr, n_all):
        first = k * H * W
        ridx = th.arange(first, first + H * W, device=dev)
        o_r, _, d_r, _, colors, _, pw = full.batch(ridx)
        val_rays.append((o_r, d_r, colors[:, -1], pw))

    train_views = []
    for k in range(0, n_tr, max(n_tr // 4, 1))[:4]:
        ridx = th.arange(k * H * W, (k + 1) * H * W, device=dev)
        _, o_n, _, d_n, colors, img_idx, pw = train.batch(ridx)
        train_views.append((o_n, d_n, colors[:, -1], img_idx, pw))

    @th.no_grad()
    def evaluate_train_views(model, forward):
        """PSNR of training views rendered from their REFINED poses (no alignment involved)."""
        vals = []
        for (o_n, d_n, c, img_idx, pw) in train_views:
            mse, n = 0.0, 0
            for s in range(0, o_n.shape[0], 16384):
                o_p, d_p, _, _ = model.camera_extrinsics(img_idx[s:s + 16384], o_n[s:s + 16384], d_n[s:s + 16384])
                rgb = forward(o_p, d_p, pw[s:s + 16384])
                mse += float(((rgb - c[s:s + 16384]) ** 2).sum())
                n += rgb.numel()
            vals.append(psnr(mse / n))
        return sum(vals) / len(vals)

    @th.no_grad()
    def evaluate(model, forward):
        """mean PSNR over the held-out views (true poses -> model frame by Kabsch), pose error."""
        params = model.compute_post_transform_params()
        vals = []
        for (o, d, c, pw) in val_rays:
            o_m, d_m, _ = model.validation_transform_rays(o, d, params)
            mse, n = 0.0, 0
            for s in range(0, o.shape[0], 16384):
                rgb = forward(o_m[s:s + 16384], d_m[s:s + 16384], pw[s:s + 16384])
                mse += float(((rgb - c[s:s + 16384]) ** 2).sum())
                n += rgb.numel()
            vals.append(psnr(mse / n))
        return sum(vals) / len(vals), float(model.compute_pose_error())

    out = {"seed": args.seed, "config": {"steps": steps, "rays": args.rays, "samples": SAMPLES, "image_size": args.size, "train_images": n_tr,
                      "val_images": args.val_images, "pose_noise": args.noise, "blur_sigmas": BLUR,
                      "alpha_ramp_epochs": [a0, a1], "lr": "5e-4 -> 1e-5 (network), 1e-3 -> 1e-5 (poses)"}}

    # ---------------- arm 1: this repo ----------------
    m = build_ours()
    init_state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    eng = TrainEngine(m, dev, loss_fn=m.training_loss)
    pose0 = float(m.compute_pose_error())
    th.cuda.synchronize()
    t0 = time.time()
    for s in range(steps):
        th.manual_seed(100000 * args.seed + s)                     # the sampling offsets of step s, on both arms
        batch = train.batch(idx[s])
        if s == 0:
            eng.step(*batch)
        else:
            if eng._graph is None:
                eng.capture(*batch)
            eng.replay(*batch)
    th.cuda.synchronize()
    out["b200"] = {"train_seconds": time.time() - t0, "final_train_loss": float(eng.last_logs["loss_fine"]),
                   "skipped_steps": eng.skipped_steps()}
    eng.release_graph()
    p, e = evaluate(m, lambda o, d, pw: m.forward(o, d, pw)[0])
    out["b200"].update(psnr=p, pose_error=e, train_view_psnr=evaluate_train_views(m, lambda o, d, pw: m.forward(o, d, pw)[0]))
    out["initial_pose_error"] = pose0
    print("b200 arm:", out["b200"], flush=True, file=sys.stderr)

    # ---------------- arm 2: the unmodified reference modules ----------------
    if not args.skip_reference:
        from oracle import ref_runner
        ref = ref_runner.load_reference()
        cam_o, cam_on = train.camera_origins.clone(), train.camera_origins_noisy.clone()
        rm = ref_runner.build_barf(ref, n_tr, SAMPLES, NEAR, FAR, n_batches, cam_o, cam_on, BLUR, BLUR[0],
                                   alpha_epochs=(a0, a1)).to(dev)
        # same initial weights (state-dict keys are the reference's own)
        missing = rm.load_state_dict(init_state, strict=False)
        assert not missing.missing_keys, missing
        # build_barf's 200000-step decay -> this run's
        groups = [dict(params=list(rm.model_radiance.parameters()), lr0=5e-4, lr1=1e-5),
                  dict(params=list(rm.camera_extrinsics.parameters()), lr0=1e-3, lr1=1e-5)]
        opt = th.optim.Adam([{"params": gr["params"], "lr": gr["lr0"]} for gr in groups], eps=1e-5)
        logf = [log_decay_factor(gr["lr0"], gr["lr1"], decay_end) for gr in groups]
        th.cuda.synchronize()
        t0 = time.time()
        for s in range(steps):
            th.manual_seed(100000 * args.seed + s)
            batch = train.batch(idx[s])
            for gi, gr in enumerate(groups):
                opt.param_groups[gi]["lr"] = le_nice_lr(gr["lr0"], logf[gi], decay_end, s + 1)
            opt.zero_grad(set_to_none=True)
            loss = rm.training_step(batch, s)            # epoch = s / n_batches, as Lightning would pass it
            loss.backward()
            opt.step()
        th.cuda.synchronize()
        out["reference"] = {"train_seconds": time.time() - t0, "final_train_loss": float(loss)}
        rp, re = evaluate(rm, lambda o, d, pw: rm.forward(o, d, pw)[0])
        out["reference"].update(psnr=rp, pose_error=re,
                                train_view_psnr=evaluate_train_views(rm, lambda o, d, pw: rm.forward(o, d, pw)[0]))
        out["delta_train_view_db"] = out["b200"]["train_view_psnr"] - out["reference"]["train_view_psnr"]
        out["delta_db"] = out["b200"]["psnr"] - rp
        out["pose_error_ratio"] = out["b200"]["pose_error"] / max(re, 1e-12)
        # the trained bf16-path weights evaluated by the reference's fp32 forward
        rm.load_state_dict({k: v for k, v in m.state_dict().items()}, strict=False)
        xp, xe = evaluate(rm, lambda o, d, pw: rm.forward(o, d, pw)[0])
        out["b200_weights_in_reference_forward"] = {"psnr": xp, "pose_error": xe,
                                                    "train_view_psnr": evaluate_train_views(rm, lambda o, d, pw: rm.forward(o, d, pw)[0])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
