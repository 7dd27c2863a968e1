"""Small driver for ncu: the two resampling kernels at render scale (no timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from nerf_experiments_b200 import ops
dev = th.device("cuda:0")
Br = 262144
tc0, tc1 = ops.sample_uniform(2.0, 8.0, Br, 64, dev, None, th.rand((Br, 1), device=dev), -1.0)
w = th.rand((Br, 64), device=dev) ** 4
dl = (tc1 - tc0).contiguous()
edges = th.linspace(0, 1, 65, device=dev).repeat(Br, 1)
cdf = th.cat((th.zeros(Br, 1, device=dev), th.cumsum(w, 1)), 1)
cdf = cdf / cdf[:, -1:]
u = th.rand((Br,), device=dev)
for _ in range(2):
    ops.resample_alloc(tc0, w, dl, 256, 2.0, 8.0)
    ops.resample_icdf(edges, cdf, 192, u)
th.cuda.synchronize()
print("ok")
