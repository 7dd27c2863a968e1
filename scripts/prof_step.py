"""Small driver for ncu: a few training steps (no timing).
usage: prof_step.py barf [rays] [steps]   BarfModel.training_step semantics of the bench workload (eager)
       prof_step.py garf [rays] [steps]   GARF step (garf/main.py shape: 64 + 192 samples)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "barf"
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = th.device("cuda:0")
if which == "barf":
    from nerf_experiments_b200 import scene
    from nerf_experiments_b200.engine import TrainEngine
    # random images / seeded poses instead of the sphere-traced scene: a profiler run should not spend
    # its launch budget on ~17 000 scene-generation kernels
    from types import SimpleNamespace
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher
    gcpu = th.Generator().manual_seed(0)
    c2w = scene.look_at_poses(bench.N_IMAGES, 4.0, gcpu)
    noisy = c2w.clone()
    noisy[:, :3, 3] += 0.1 * th.randn((bench.N_IMAGES, 3), generator=gcpu)
    batcher = GpuRayBatcher(th.rand((bench.N_IMAGES, 64, 64, len(bench.BLUR_SIGMAS), 3), generator=gcpu), c2w,
                            64 / 2 / 0.36, noisy, bench.BLUR_SIGMAS, device=dev)
    sc = SimpleNamespace(batcher=batcher, n_images=bench.N_IMAGES)
    model = bench.build_barf_model(sc, len(sc.batcher) // rays)
    eng = TrainEngine(model, dev, loss_fn=model.training_loss)
    g = th.Generator(device=dev).manual_seed(0)
    idx = th.randint(0, len(sc.batcher), (steps, rays), device=dev, generator=g)
    for s in range(steps):
        loss = eng.step(*sc.batcher.batch(idx[s]))
else:
    from nerf_experiments_b200.model_garf import GarfModel, garf_engine
    th.manual_seed(1337)
    m = GarfModel(2.0, 7.0, 64, 192, 0.5, 1.5, 1.0, 1e-3, 1e-4, 100000, 0.0, 1e-3, 1e-4, 100000, 0.0).to(dev)
    m.train()
    eng = garf_engine(m, dev)
    g = th.Generator().manual_seed(0)
    o = (th.nn.functional.normalize(th.randn((rays, 3), generator=g), dim=1) * 4.0).to(dev)
    d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((rays, 3), generator=g), dim=1).to(dev)
    tgt = th.rand((rays, 3), generator=g).to(dev)
    for s in range(steps):
        loss = eng.step(o, d, tgt)
th.cuda.synchronize()
print("loss", float(loss))
