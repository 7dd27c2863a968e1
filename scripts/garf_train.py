"""GARF with pose refinement end to end on this repo's kernels (garf/main.py's configuration: near 2 / far 7,
64 proposal + 192 radiance samples per ray, Gaussian widths in [0.5, 2], width learning-rate factor 16,
pose noise 0.15, lr 2e-4 -> 2e-5 / 5e-4 -> 5e-5 / 4e-3 -> 8e-4 with the decay ends scaled to the run length)
on the synthetic SDF scene: the captured training graph (fused GARF kernels, PropNet chain, pose kernels,
one fused Adam over five groups) for --steps steps, then the Kabsch-aligned pose error and the PSNR of
held-out views rendered from their true poses mapped into the model's frame.

There is no reference arm for this model: garf/model_garf.py needs the third-party `nerfacc` package, which
is neither in /root/reference nor installed (DESIGN.md, a12 "parity unpinned"). This run shows that the
fused path trains — poses converge and the held-out PSNR rises — not parity.

usage: python scripts/garf_train.py [--steps 3000] [--rays 4096] [--size 200] [--images 30]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--size", type=int, default=200)
    ap.add_argument("--images", type=int, default=30)
    ap.add_argument("--val-images", type=int, default=3)
    ap.add_argument("--noise", type=float, default=0.15)
    ap.add_argument("--camera-lr", type=float, nargs=2, default=[4e-3, 8e-4], help="start stop (garf/main.py: 4e-3 8e-4)")
    args = ap.parse_args()
    from nerf_experiments_b200 import ops, scene
    from nerf_experiments_b200.model_garf import garf_engine
    from nerf_experiments_b200.model_garf_camera_calibration import CameraCalibrationModel

    dev = th.device("cuda:0")
    n_tr, n_all, H, W = args.images, args.images + args.val_images, args.size, args.size
    sc = scene.make_scene(n_all, H, W, dev, rotation_noise=args.noise, translation_noise=args.noise)
    n_train_rays = n_tr * H * W
    steps = args.steps
    th.manual_seed(1337)
    m = CameraCalibrationModel(n_tr, args.camera_lr[0], args.camera_lr[1], max(steps // 3, 1), 10,
                               2.0, 7.0, 64, 192, 0.5, 2.0, 16.0,
                               5e-4, 5e-5, max(2 * steps // 3, 1), 0.0,
                               2e-4, 2e-5, steps, 0.0).to(dev)
    m.train()
    eng = garf_engine(m, dev)
    cam_true = sc.c2w[:n_tr, :3, 3].to(dev)
    cam_noisy = sc.c2w_noisy[:n_tr, :3, 3].to(dev)

    @th.no_grad()
    def pose_error():
        """mean distance between the true camera origins and the refined ones after the Kabsch alignment
        (garf/model_camera_calibration.py:196-230)."""
        idx = th.arange(n_tr, device=dev)
        o_pred, _, _, _ = m.camera_extrinsics(idx, cam_noisy, th.zeros_like(cam_noisy))
        R, t, c = ops.kabsch(o_pred, cam_true, True)
        aligned = (o_pred @ R.T) * c + t
        return float((aligned - cam_true).norm(dim=1).mean()), (R, t, c)

    @th.no_grad()
    def val_psnr():
        # true poses of the held-out views -> model frame: the inverse of the alignment above
        _, (R, t, c) = pose_error()
        vals = []
        for k in range(n_tr, n_all):
            sl = slice(k * H * W, (k + 1) * H * W)
            o = ((sc.origins_true[sl] - t) / c) @ R
            d = sc.directions_true[sl] @ R
            mse, n = 0.0, 0
            for s in range(0, H * W, 8192):
                rgb = m(o[s:s + 8192].contiguous(), d[s:s + 8192].contiguous())[0]
                mse += float(((rgb - sc.colors[sl][s:s + 8192]) ** 2).sum())
                n += rgb.numel()
            vals.append(-10 * math.log10(max(mse / n, 1e-12)))
        return sum(vals) / len(vals)

    e0, _ = pose_error()
    m.eval(); p0 = val_psnr(); m.train()
    g = th.Generator(device=dev).manual_seed(3)
    idx = th.randint(0, n_train_rays, (steps, args.rays), device=dev, generator=g)

    def batch(s):
        i = idx[s]
        return (sc.origins_true[i], sc.origins[i], sc.directions_true[i], sc.directions[i], sc.colors[i],
                sc.image_index[i].long())

    th.cuda.synchronize()
    t0 = time.time()
    losses = []
    for s in range(steps):
        b = batch(s)
        if s == 0:
            eng.step(*b)
        else:
            if eng._graph is None:
                eng.capture(*b)
            eng.replay(*b)
        if s % max(steps // 10, 1) == 0 or s == steps - 1:
            losses.append((s, float(eng.last_logs["loss_fine"])))
    th.cuda.synchronize()
    secs = time.time() - t0
    eng.release_graph()
    e1, _ = pose_error()
    m.eval(); p1 = val_psnr()
    print(json.dumps({"config": {"steps": steps, "rays": args.rays, "image_size": args.size, "train_images": n_tr,
                                 "val_images": args.val_images, "pose_noise": args.noise, "camera_lr": args.camera_lr,
                                 "samples": "64 proposal + 192 radiance"},
                      "train_seconds": secs, "rays_per_s": steps * args.rays / secs, "skipped_steps": eng.skipped_steps(),
                      "pose_error_initial": e0, "pose_error_final": e1, "val_psnr_initial": p0, "val_psnr_final": p1,
                      "loss_curve": losses,
                      "reference_arm": "unavailable: garf/model_garf.py needs nerfacc (absent); see the BARF run for reference parity"}))


if __name__ == "__main__":
    main()
