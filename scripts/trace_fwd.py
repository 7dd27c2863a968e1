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
o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(dev)
t0, t1 = ops.sample_uniform(2.0, 8.0, B, S, dev, None, th.rand((B, 1), device=dev), -1.0)
pw = th.full((B, 1), 1 / 555.0, device=dev)
with th.no_grad():
    field_rays(net, o, d, t0, t1, pw, "middle")
trace = th.zeros(64 * 4, dtype=th.int64, device=dev)
L = _lib.lib()
L.nerfb200_debug_trace_fwd.argtypes = [ctypes.c_void_p]
L.nerfb200_debug_trace_fwd(trace.data_ptr())
with th.no_grad():
    field_rays(net, o, d, t0, t1, pw, "middle")
th.cuda.synchronize()
t = trace.cpu().view(64, 4)[:12]
base = t[0, 0].item()
print("op  a_ready  mma_issued  acc_full_seen  epi_done   | mma_issue  mma_exec_wait  epilogue  handoff")
prev_epi = None
for i in range(12):
    a, b, c, dd = [x.item() - base for x in t[i]]
    hand = (a - prev_epi) if prev_epi is not None else 0
    print(f"{i:2d} {a:8d} {b:8d} {c:8d} {dd:8d}   | {b-a:6d} {c-b:6d} {dd-c:6d} {hand:6d}")
    prev_epi = dd
tt = trace.cpu()
print("producer empty-acquired (op, chunk):", [[(tt[64 + oi * 8 + c].item() - base) for c in range(5)] for oi in range(4)])
print("mma full-seen (op, chunk):         ", [[(tt[128 + oi * 8 + c].item() - base) for c in range(5)] for oi in range(4)])
