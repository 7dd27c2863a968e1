"""One small pass through every kernel family of libnerfb200.so, meant to be run under
`compute-sanitizer --tool memcheck` (and plain first): hierarchical NeRF step with pose refinement
(pose, sampling, fused field fwd / bwd / wgrad, compositing, resampling, Adam), a GARF step with
pose refinement (lindisp intervals, fused GARF fwd / bwd / wgrad + column sums, transmittance cdf,
icdf resampling, proposal loss), the Kabsch kernel, the inference render op and the ray batcher.
Eager launches only (no CUDA graph), tiny batches: the sanitizer serialises everything."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th


def main():
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import ops, positional_encodings as pe, _lib
    from nerf_experiments_b200.engine import TrainEngine
    from nerf_experiments_b200.model_camera_extrinsics import CameraExtrinsics
    from nerf_experiments_b200.model_garf import garf_engine
    from nerf_experiments_b200.model_garf_camera_calibration import CameraCalibrationModel
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher
    dev = th.device("cuda:0")
    th.manual_seed(0)
    g = th.Generator().manual_seed(1)
    B, n_img = 200, 4                                  # not a multiple of the 128-sample tile
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    target = th.rand((B, 3), generator=g)
    idx = th.randint(0, n_img, (B,), generator=g).int()
    pw = th.full((B, 1), 1 / 555.0)

    def net():
        ep = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, True, 1.0)
        ed = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, True, 1.0)
        m = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 1000)
        ep.alpha.fill_(7.25); ed.alpha.fill_(4.0)
        return m

    # 1. hierarchical NeRF (proposal 24 + radiance 48 samples) with pose refinement, two steps
    model = mi.NerfInterpolation(2.0, 8.0, net(), 48, "stratified_uniform", 1.0, "middle", net(), 24)
    cam = CameraExtrinsics(n_img, 1e-3, 1e-5, 1000)
    model.camera_extrinsics = cam
    model.param_groups = model.param_groups + cam.param_groups
    model = model.to(dev)
    eng = TrainEngine(model, dev)
    for _ in range(2):
        loss = eng.step(*(t.to(dev) for t in (o, d, target, idx, pw)))
    print("nerf step loss", float(loss))
    rgb, depth, w = model.render(o.to(dev), d.to(dev), pw.to(dev), 2.0, 8.0)
    print("render", tuple(rgb.shape), float(rgb.mean()))

    # 2. GARF with pose refinement, two steps
    gm = CameraCalibrationModel(n_img, 1e-3, 1e-5, 40, 10, 2.0, 7.0, 16, 24, 0.5, 1.5, 2.0,
                                1e-3, 1e-4, 50, 0.0, 2e-3, 1e-4, 60, 0.0).to(dev)
    gm.train()
    ge = garf_engine(gm, dev)
    o_n = (o + 0.05 * th.randn((B, 3), generator=g)).to(dev)
    d_n = th.nn.functional.normalize(d + 0.05 * th.randn((B, 3), generator=g), dim=1).to(dev)
    for _ in range(2):
        ge.step(o.to(dev), o_n, d.to(dev), d_n, target.to(dev), idx.to(dev).long())
    print("garf step loss", float(ge.last_logs["loss_fine"]))

    # 3. Kabsch
    a = th.randn((40, 3), generator=g).to(dev)
    R, t, c = ops.kabsch(a, a * 1.1 + 0.2, True)
    print("kabsch scale", float(c))

    # 4. ray batcher + blur pyramid
    imgs = th.rand((n_img, 24, 24, 3), generator=g).to(dev)
    from nerf_experiments_b200.scene import gaussian_blur_pyramid
    c2w = th.eye(4).repeat(n_img, 1, 1).contiguous().to(dev)
    rb = GpuRayBatcher(gaussian_blur_pyramid(imgs, [1.0, 0.0]), c2w, 30.0, c2w.clone(), [1.0, 0.0], device=dev)
    batch = rb.batch(th.randint(0, len(rb), (100,), device=dev))
    print("ray batch", [tuple(x.shape) for x in batch])
    th.cuda.synchronize()
    print("sanitize case done, kernels launched", _lib.launch_count())


if __name__ == "__main__":
    main()
