"""End-of-training PSNR parity (BASELINE.json north_star: within 0.1 dB of the reference path).

Trains the same NeRF twice on the synthetic SDF scene, from the same initial weights, on the same
ray batches with the same sampling uniforms:
  arm "b200"      this repo's path (TrainEngine: fused bf16 tcgen05 MLP kernels, compositing,
                  fused Adam), and
  arm "reference" the reference's arithmetic in PyTorch fp32 (oracle/ref_step.py restates
                  NerfInterpolation.forward + MSE; torch.optim.Adam(eps=1e-5) with the
                  SchedulerLeNice closed form), executed with torch on the same GPU,
then renders held-out views with both and reports PSNR.  The oracle is used here as the checker
of a test run, never as a product path.

usage: python scripts/psnr_parity.py [--steps 2000] [--rays 1024] [--samples 64] [--size 64]
prints one JSON line."""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

NEAR, FAR = 2.0, 8.0


def build(samples: int, decay_end: int):
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    th.manual_seed(1337)
    ep = pe.BarfPositionalEncoding(10, 10.0, 0.0, 1.0, False, 1.0)
    ed = pe.BarfPositionalEncoding(4, 4.0, 0.0, 1.0, False, 1.0)
    ep.alpha.fill_(10.0)
    ed.alpha.fill_(4.0)     # vanilla-as-BARF: alpha = levels (reference barf/run_vanilla_as_barf.py:150-166)
    net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 5e-5, decay_end)
    return mi.NerfInterpolation(NEAR, FAR, net, samples, "equidistant", -1.0, "middle", None, 0)


def psnr(mse: float) -> float:
    return -10.0 * math.log10(max(mse, 1e-12))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--images", type=int, default=24)
    ap.add_argument("--val-images", type=int, default=4)
    args = ap.parse_args()
    from nerf_experiments_b200 import scene
    from nerf_experiments_b200.engine import TrainEngine
    from nerf_experiments_b200.model_interpolation import le_nice_lr, log_decay_factor
    from oracle import ref_step
    dev = th.device("cuda:0")
    th.backends.cuda.matmul.allow_tf32 = False      # the reference arm is true fp32
    sc = scene.make_scene(args.images + args.val_images, args.size, args.size, dev)
    n_px = args.size * args.size
    train_rays = args.images * n_px
    val = slice(train_rays, train_rays + args.val_images * n_px)
    perm_gen = th.Generator().manual_seed(7)
    batches = [th.randint(0, train_rays, (args.rays,), generator=perm_gen).to(dev) for _ in range(args.steps)]

    # ---- arm "b200" ---------------------------------------------------------------------------
    model = build(args.samples, args.steps)
    init = {k: v.detach().clone() for k, v in model.model_radiance.state_dict().items()}
    eng = TrainEngine(model, dev)
    th.cuda.manual_seed(99)
    t0 = time.time()
    for idx in batches:
        o, d, c, _, pw = sc.batch(idx)
        eng.step(o, d, c, None, pw)
    th.cuda.synchronize()
    t_ours = time.time() - t0

    def render_val(fn):
        th.cuda.manual_seed(1234)
        out = []
        with th.no_grad():
            for s in range(val.start, val.stop, 4096):
                sl = slice(s, min(s + 4096, val.stop))
                out.append(fn(sc.origins[sl], sc.directions[sl]))
        pred = th.cat(out)
        return float(th.nn.functional.mse_loss(pred, sc.colors[val]))

    model.eval()
    pw_val = th.full((4096, 1), sc.pixel_width, device=dev)
    mse_ours = render_val(lambda o, d: model(o, d, pw_val[: o.shape[0]])[0])

    # ---- arm "reference": the reference's arithmetic, PyTorch fp32 on the same GPU ----------------
    cfg = dict(n_hidden=4, n_segments=2, delayed_direction=True, delayed_density=False)
    pe_cfg = dict(pos_levels=10, dir_levels=4, scale=1.0, identity=False,
                  alpha_pos=th.tensor(10.0, device=dev), alpha_dir=th.tensor(4.0, device=dev))
    sd = {k: v.to(dev).clone().requires_grad_(v.dim() > 0) for k, v in init.items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = th.optim.Adam(params, lr=5e-4, eps=1e-5)
    logf = log_decay_factor(5e-4, 5e-5, args.steps)
    th.cuda.manual_seed(99)
    t0 = time.time()
    with th.device(dev):      # the oracle's factory calls (th.ones, th.linspace, ...) land on the GPU
        for k, idx in enumerate(batches):
            o, d, c, _, _ = sc.batch(idx)
            u = th.rand((args.rays, 1), device=dev)      # same draw as NerfInterpolation's offset
            for gparam in opt.param_groups:
                gparam["lr"] = le_nice_lr(5e-4, logf, args.steps, k + 1)
            opt.zero_grad(set_to_none=True)
            loss, _ = ref_step.barf_step(sd, cfg, pe_cfg, None, None, None, o, d, c, NEAR, FAR, args.samples,
                                         "middle", {"offset": u}, None, 0, "equidistant", -1.0, False)
            loss.backward()
            opt.step()
        th.cuda.synchronize()
        t_ref = time.time() - t0

        def ref_render(o, d):
            u = th.rand((o.shape[0], 1), device=dev)
            rgb, _, _ = ref_step.render(sd, cfg, pe_cfg, o, d, NEAR, FAR, args.samples, "middle", {"offset": u},
                                        None, 0, "equidistant", -1.0, False)
            return rgb
        mse_ref = render_val(ref_render)
        # cross-check: the b200-trained weights rendered by the fp32 reference arithmetic
        sd_ours = {k: v.detach() for k, v in model.model_radiance.state_dict().items()}
        sd_keep, sd = sd, sd_ours
        mse_cross = render_val(ref_render)
        sd = sd_keep

    out = {"steps": args.steps, "rays": args.rays, "samples": args.samples, "image": args.size,
           "train_images": args.images, "val_images": args.val_images,
           "psnr_b200": round(psnr(mse_ours), 4), "psnr_reference_fp32": round(psnr(mse_ref), 4),
           "psnr_b200_weights_in_fp32_reference": round(psnr(mse_cross), 4),
           "delta_db": round(psnr(mse_ours) - psnr(mse_ref), 4),
           "train_s_b200": round(t_ours, 2), "train_s_reference_torch_gpu": round(t_ref, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
