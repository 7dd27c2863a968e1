"""Small driver for ncu: compositing forward / backward at render scale (no timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from nerf_experiments_b200 import ops
dev = th.device("cuda:0")
B, S = 262144, 128
sigma = th.nn.functional.softplus(th.randn((B, S), device=dev))
delta = th.full((B, S), 6.0 / S, device=dev)
rgb = th.rand((B, S, 3), device=dev)
g_rgb = th.randn((B, 3), device=dev)
g_w = th.randn((B, S), device=dev)
for _ in range(2):
    ops.composite_fwd(sigma, delta, rgb)
    ops.composite_bwd(sigma, delta, rgb, g_rgb, g_w)
th.cuda.synchronize()
print("ok")
