"""Small driver for ncu: a few BARF training steps of the bench workload (no timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
from nerf_experiments_b200.engine import TrainEngine

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = th.device("cuda:0")
model = bench.build_model(20)
eng = TrainEngine(model, dev)
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((rays, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((rays, 3), generator=g), dim=1).to(dev)
target = th.rand((rays, 3), generator=g).to(dev)
idx = th.randint(0, 20, (rays,), generator=g).int().to(dev)
pw = th.full((rays, 1), 1 / 555.0, device=dev)
for s in range(steps):
    loss = eng.step(o, d, target, idx, pw)
th.cuda.synchronize()
print("loss", loss.item())
