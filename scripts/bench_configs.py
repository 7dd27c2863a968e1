"""Throughput of the BASELINE.json configurations other than the one bench.py reports (C2), on one
B200: C1 vanilla 1024 x 64, C3 Mip-BARF 8192 rays x (64 + 256) hierarchical with one shared
network, C4 GARF 1024..16384 rays x (64 + 192), C5 800 x 800 full-image render in 16384-ray
chunks.  CUDA events on the launching stream, 3 warm-up + 10 timed steps, synthetic rays.
usage: python scripts/bench_configs.py [c1 c3 c4 c5]  -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from nerf_experiments_b200 import model_interpolation as mi
from nerf_experiments_b200 import model_interpolation_architecture as arch
from nerf_experiments_b200 import positional_encodings as pe
from nerf_experiments_b200.engine import TrainEngine
from nerf_experiments_b200.model_camera_extrinsics import CameraExtrinsics

dev = th.device("cuda:0")
which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
out = {}


def rays(B, n_images=20, seed=0):
    g = th.Generator().manual_seed(seed)
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    return (o.to(dev), d.to(dev), th.rand((B, 3), generator=g).to(dev),
            th.randint(0, n_images, (B,), generator=g).int().to(dev), th.full((B, 1), 1 / 555.0, device=dev))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if "c1" in which:   # vanilla NeRF as run_vanilla_as_barf.py: PE 10/4 without identity, no pose refinement
    th.manual_seed(1337)
    net = arch.NerfModel(4, 256, True, False, 2, pe.BarfPositionalEncoding(10, 10.0, 0.0, 0.0, False, 1.0),
                         pe.BarfPositionalEncoding(4, 4.0, 0.0, 0.0, False, 1.0), 5e-4, 1e-5, 200000)
    model = mi.NerfInterpolation(2.0, 8.0, net, 64, "stratified_uniform", -1.0, "middle", None, 0)
    eng = TrainEngine(model, dev)
    o, d, tgt, idx, pw = rays(1024)
    ms = timeit(lambda: eng.step(o, d, tgt, None, pw))
    out["c1_vanilla_1024x64"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(1024 / ms * 1e3)}
    eng.capture(o, d, tgt, None, pw)          # the same step as one CUDA graph
    ms = timeit(lambda: eng.replay(o, d, tgt, None, pw))
    out["c1_vanilla_1024x64_cuda_graph"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(1024 / ms * 1e3),
                                            "library_launches_per_step": int(eng.launches_per_replay)}
    print(out, flush=True)
    # C2 (the bench.py workload) eager vs graph, for reference
    import bench
    m2 = bench.build_model(20)
    e2 = TrainEngine(m2, dev)
    o2, d2, t2, i2, p2 = rays(4096)
    ms_e = timeit(lambda: e2.step(o2, d2, t2, i2, p2))
    e2.capture(o2, d2, t2, i2, p2)
    ms_g = timeit(lambda: e2.replay(o2, d2, t2, i2, p2))
    out["c2_barf_4096x128_eager_vs_graph"] = {"ms_eager": round(ms_e, 4), "ms_graph": round(ms_g, 4)}
    print(out, flush=True)
    del e2, m2

if "c3" in which:   # Mip-BARF: integrated encoding, ONE network as proposal and radiance model, poses
    from nerf_experiments_b200.model_mip import MipBarf
    th.manual_seed(1337)
    ep = pe.IntegratedFourierFeatures(levels=10, include_identity=True, scale=1., distribute_variance=False)
    ed = pe.BarfPositionalEncoding(0, 1, 0, 1, True)
    net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 200000)
    model = MipBarf(model_radiance=net, samples_per_ray_radiance=256, n_training_images=20,
                    camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5, camera_learning_rate_decay_end=200000,
                    uniform_sampling_strategy="equidistant", uniform_sampling_offset_size=-1., samples_per_ray_proposal=64,
                    sigma_decay_start_step=0, sigma_decay_end_step=100000, start_blur_sigma=8., start_pixel_width_sigma=1.5)
    eng = TrainEngine(model, dev)
    o, d, tgt, idx, pw = rays(8192)
    ms = timeit(lambda: eng.step(o, d, tgt, idx, pw, coarse_weight=0.1), iters=5)
    out["c3_mipbarf_8192x(64+256)"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(8192 / ms * 1e3),
                                       "samples_per_s": round(8192 * 320 / ms * 1e3)}
    print(out, flush=True)
    del eng, model, net

if "c4" in which:   # GARF: Gaussian-activation radiance + proposal network, inverse-CDF resampling (garf/main.py shape)
    from nerf_experiments_b200.model_garf import GarfModel
    th.manual_seed(1337)
    m = GarfModel(2.0, 7.0, 64, 192, 0.5, 1.5, 1.0, 1e-3, 1e-4, 100000, 0.0, 1e-3, 1e-4, 100000, 0.0).to(dev)
    m.train()
    for B in (1024, 4096, 16384):
        o, d, tgt, idx, pw = rays(B)
        step = [0]

        def one():
            m.training_step((o, d, tgt), step[0])
            step[0] += 1
        ms = timeit(one, iters=5)
        out[f"c4_garf_{B}x(64+192)"] = {"ms_per_step": round(ms, 4), "rays_per_s": round(B / ms * 1e3),
                                        "note": "Linear layers as cuBLAS GEMMs between the activation kernels (not yet fused)"}
        print(out, flush=True)
    del m

if "c5" in which:   # 800 x 800 render: 640 000 rays in 16384-ray chunks, C2 network, no gradient
    from nerf_experiments_b200.ray_batcher import render_image
    import bench
    model = bench.build_model(20).to(dev)
    o, d, _, _, _ = rays(640000)
    ms = timeit(lambda: render_image(model, o, d, 800, 800, 1 / 555.0, chunk=16384), iters=3, warm=1)
    out["c5_render_800x800_128spp"] = {"ms_per_image": round(ms, 3), "rays_per_s": round(640000 / ms * 1e3)}
    # hierarchical render (64 coarse + 128 fine samples, two networks)
    th.manual_seed(1)
    def net():
        return arch.NerfModel(4, 256, True, False, 2, pe.BarfPositionalEncoding(10, 10.0, 0.0, 0.0, True, 1.0),
                              pe.BarfPositionalEncoding(4, 4.0, 0.0, 0.0, True, 1.0))
    hm = mi.NerfInterpolation(2.0, 8.0, net(), 128, "equidistant", -1.0, "middle", net(), 64).to(dev)
    ms = timeit(lambda: render_image(hm, o, d, 800, 800, 1 / 555.0, chunk=16384), iters=3, warm=1)
    out["c5_render_800x800_64+128spp"] = {"ms_per_image": round(ms, 3), "rays_per_s": round(640000 / ms * 1e3)}

print(json.dumps(out))
