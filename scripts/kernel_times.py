"""Per-kernel times of the bench workload's training step (CUDA events on the launching stream):
python scripts/kernel_times.py [rays] [iters].  With NERFB200_LIB=<other .so> the same script times
another build of the library, so two builds can be compared on the same box in one gpurun call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
from nerf_experiments_b200.engine import TrainEngine

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = th.device("cuda:0")
model = bench.build_model(20)
eng = TrainEngine(model, dev)
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((rays, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((rays, 3), generator=g), dim=1).to(dev)
target = th.rand((rays, 3), generator=g).to(dev)
idx = th.randint(0, 20, (rays,), generator=g).int().to(dev)
pw = th.full((rays, 1), 1 / 555.0, device=dev)
for s in range(3):
    loss = eng.step(o, d, target, idx, pw)
ff = model.model_radiance.fused_field()
ff.timers = {}
th.cuda.synchronize()
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record()
for s in range(iters):
    loss = eng.step(o, d, target, idx, pw)
e1.record()
th.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / iters
macs = ff.macs_per_sample()
S = model.samples_per_ray_radiance
line = [f"lib={os.environ.get('NERFB200_LIB', 'default')}", f"step {step_ms:.3f} ms ({rays / step_ms * 1e3:.0f} rays/s)"]
for k, v in ff.timers.items():
    ms = sorted(a.elapsed_time(b) for a, b in v)[len(v) // 2]
    key = {"mlp_fwd_train": "fwd", "mlp_bwd_inputs": "bwd_inputs", "mlp_bwd": "bwd", "mlp_wgrad": "wgrad"}[k]
    line.append(f"{k} {ms:.3f} ms {2 * macs[key] * rays * S / ms / 1e9:.0f} TF/s")
from nerf_experiments_b200 import ops
from nerf_experiments_b200.field_function import field_rays
t0, t1 = ops.sample_uniform(2.0, 8.0, rays, S, dev, None, th.rand((rays, 1), device=dev), -1.0)
ff.timers = {}
with th.no_grad():
    for s in range(iters + 2):
        field_rays(model.model_radiance, o, d, t0, t1, pw, "middle")
th.cuda.synchronize()
assert list(ff.timers) == ["mlp_fwd"], list(ff.timers)
v = ff.timers["mlp_fwd"]
ms = sorted(a.elapsed_time(b) for a, b in v)[len(v) // 2]
line.append(f"mlp_fwd(infer) {ms:.3f} ms {2 * macs['fwd'] * rays * S / ms / 1e9:.0f} TF/s")
print(" | ".join(line), f"| loss {loss.item():.5f}")
