import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
from nerf_experiments_b200 import ops, _lib
from nerf_experiments_b200.field_function import field_rays
dev = th.device("cuda:0")
model = bench.build_model(20).to(dev)
net = model.model_radiance
B, S = 4096, 128
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(dev).requires_grad_()
d = th.nn.functional.normalize(-o.detach().cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(dev).requires_grad_()
t0, t1 = ops.sample_uniform(2.0, 8.0, B, S, dev, None, th.rand((B, 1), device=dev), -1.0)
pw = th.full((B, 1), 1 / 555.0, device=dev)
L = _lib.lib()
L.nerfb200_debug_trace_bwd.argtypes = [ctypes.c_void_p]
trace = th.zeros(64 * 8, dtype=th.int64, device=dev)
for it in range(2):
    sigma, rgb = field_rays(net, o, d, t0, t1, pw, "middle")
    if it == 1:
        L.nerfb200_debug_trace_bwd(trace.data_ptr())
    (sigma.sum() + rgb.sum()).backward()
th.cuda.synchronize()
n_ops = net.fused_field().bwd[True].program.n_ops
t = trace.cpu()[:256].view(64, 4)[:n_ops]
base = t[0, 0].item()
print("op  a_ready  mma_issued  acc_full_seen  epi_done   | mma_issue  mma_exec_wait  epilogue  handoff")
prev = None
for i in range(n_ops):
    a, b, c, dd = [x.item() - base for x in t[i]]
    hand = (a - prev) if prev is not None else 0
    print(f"{i:2d} {a:8d} {b:8d} {c:8d} {dd:8d}   | {b-a:6d} {c-b:6d} {dd-c:6d} {hand:6d}")
    prev = dd
