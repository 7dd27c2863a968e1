"""Cycle trace of the fused MLP kernels (block 0, its second tile): per op the MMA warp's and the
row threads' time stamps, and per K chunk when the A slab / the weight image became available.
usage: python scripts/trace_mlp.py [fwd|bwd]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
from nerf_experiments_b200 import ops, _lib
from nerf_experiments_b200.field_function import field_rays
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
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
for f in (L.nerfb200_debug_trace_fwd, L.nerfb200_debug_trace_bwd):
    f.argtypes = [ctypes.c_void_p]
trace = th.zeros(512, dtype=th.int64, device=dev)
for it in range(2):
    if it == 1 and which == "fwd":
        L.nerfb200_debug_trace_fwd(trace.data_ptr())
    sigma, rgb = field_rays(net, o, d, t0, t1, pw, "middle")
    if it == 1 and which == "bwd":
        L.nerfb200_debug_trace_bwd(trace.data_ptr())
    (sigma.sum() + rgb.sum()).backward()
th.cuda.synchronize()
ff = net.fused_field()
prog = ff.compiled.program if which == "fwd" else ff.bwd[True].program
n_ops = prog.n_ops
t = trace.cpu()
base = t[0].item()
print(f"{which}: op | buf_free mma_issued acc_seen epi_done | issue exec_wait epilogue | per chunk (slab_ready, weights_full)")
for i in range(min(n_ops, 16)):
    a, b, c, dd = [t[i * 4 + k].item() - base for k in range(4)]
    ch = [(t[128 + i * 8 + k].item() - base, t[256 + i * 8 + k].item() - base) for k in range(min(prog.ops[i].n_chunks, 8))]
    print(f"{i:2d} | {a:7d} {b:7d} {c:7d} {dd:7d} | {b-a:6d} {c-b:6d} {dd-c:6d} | " + " ".join(f"({x},{y})" for x, y in ch))
if which == "bwd":
    e = [t[392 + k].item() - base for k in range(5)]
    print("tile start (thread 0): tile_done seen", e[0], "| drain waited +", e[1] - e[0], "| head gradient +", e[2] - e[1],
          "| published +", e[3] - e[2], "| ops done (ray-gradient reduction starts) at", e[4])
if which == "fwd":
    e = [t[392 + k].item() - base for k in range(4)]
    print("tile start (thread 0): tile_done seen", e[0], "| drain waited +", e[1] - e[0], "| encoded +", e[2] - e[1], "| published +", e[3] - e[2])
    a3 = t[3 * 4 + 2].item() - base
    print(f"op 3 epilogue, relative to its acc_full seen by thread 0 ({a3}); drain wait done at +{t[399].item() - base - a3}")
    for name, off in (("thread 0 (warp 0)", 400), ("thread 480 (warp 15)", 424)):
        for j in range(4):
            e = [t[off + 6 * j + k].item() - base - a3 for k in range(5)]
            print(f"  {name} slab {j}: tmem-ld done {e[0]:6d} | math +{e[1]-e[0]:4d} | STS +{e[2]-e[1]:4d} | proxy fence +{e[3]-e[2]:4d} | arrive +{e[4]-e[3]:4d} -> {e[4]:6d}")
    print("  op 4 MMA warp, chunk MMAs issued + committed at:", [t[448 + c].item() - base - a3 for c in range(4)],
          " op 4 chunks (slab_ready seen, weights seen):", [(t[128 + 32 + c].item() - base - a3, t[256 + 32 + c].item() - base - a3) for c in range(4)],
          " op 4 acc_full seen:", t[4 * 4 + 2].item() - base - a3)
