"""Kernel micro-benchmarks on one B200 (CUDA events, warm-up, inputs larger than L2):
fused MLP forward with / without the activation stash, compositing fwd/bwd and resampling at
render scale with achieved HBM GB/s against MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
from nerf_experiments_b200 import ops
from nerf_experiments_b200.field_function import field_rays

dev = th.device("cuda:0")
peaks = bench.measured_peaks()
out = {}


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


# ---- fused MLP forward: inference vs training (stash) ----
model = bench.build_model(20).to(dev)
net = model.model_radiance
B, S = 4096, 128
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(dev)
t0, t1 = ops.sample_uniform(2.0, 8.0, B, S, dev, None, th.rand((B, 1), device=dev), -1.0)
pw = th.full((B, 1), 1 / 555.0, device=dev)
macs = None


def fwd_infer():
    with th.no_grad():
        field_rays(net, o, d, t0, t1, pw, "middle")


ms = timeit(fwd_infer)
macs = net.fused_field().macs_per_sample()
out["mlp_fwd_infer"] = {"ms": ms, "tflops": 2 * macs["fwd"] * B * S / ms / 1e9}
print(out["mlp_fwd_infer"], flush=True)

# ---- compositing at render scale: 65536 rays x 128/256 samples (inputs >> L2) ----
for (Br, Sr) in ((262144, 128), (131072, 256), (4096, 128)):
    sigma = th.nn.functional.softplus(th.randn((Br, Sr), device=dev))
    delta = th.full((Br, Sr), 6.0 / Sr, device=dev)
    rgb = th.rand((Br, Sr, 3), device=dev)
    g_rgb = th.randn((Br, 3), device=dev)
    g_w = th.randn((Br, Sr), device=dev)
    ms_f = timeit(lambda: ops.composite_fwd(sigma, delta, rgb))
    ms_b = timeit(lambda: ops.composite_bwd(sigma, delta, rgb, g_rgb, g_w))
    bytes_f = Br * Sr * 24 + Br * 12
    bytes_b = Br * Sr * (20 + 4 + 16) + Br * 12
    out[f"composite_fwd_{Br}x{Sr}"] = {"ms": ms_f, "GBps": bytes_f / ms_f / 1e6, "frac": bytes_f / ms_f / 1e6 / peaks["hbm_gbs"]}
    out[f"composite_bwd_{Br}x{Sr}"] = {"ms": ms_b, "GBps": bytes_b / ms_b / 1e6, "frac": bytes_b / ms_b / 1e6 / peaks["hbm_gbs"]}
    print(Br, Sr, out[f"composite_fwd_{Br}x{Sr}"], out[f"composite_bwd_{Br}x{Sr}"], flush=True)
    del sigma, delta, rgb, g_rgb, g_w

# ---- resampling: allocator 64 -> 256 and inverse-CDF 64 -> 192 ----
Br = 524288
tc0, tc1 = ops.sample_uniform(2.0, 8.0, Br, 64, dev, None, th.rand((Br, 1), device=dev), -1.0)
w = th.rand((Br, 64), device=dev) ** 4
ms_a = timeit(lambda: ops.resample_alloc(tc0, w, tc1 - tc0, 256, 2.0, 8.0))
bytes_a = Br * (3 * 4 * 64 + 2 * 4 * 256)
out["resample_alloc_64to256"] = {"ms": ms_a, "GBps": bytes_a / ms_a / 1e6, "frac": bytes_a / ms_a / 1e6 / peaks["hbm_gbs"],
                                 "note": "two launches: allocator + gated fallback"}
edges = th.linspace(0, 1, 65, device=dev).repeat(Br, 1)
cdf = th.cat((th.zeros(Br, 1, device=dev), th.cumsum(w, 1)), 1)
cdf = cdf / cdf[:, -1:]
u = th.rand((Br,), device=dev)
ms_i = timeit(lambda: ops.resample_icdf(edges, cdf, 192, u))
bytes_i = Br * (4 * 65 * 2 + 4 + 4 * 193)
out["resample_icdf_64to192"] = {"ms": ms_i, "GBps": bytes_i / ms_i / 1e6, "frac": bytes_i / ms_i / 1e6 / peaks["hbm_gbs"]}
print(out["resample_alloc_64to256"], out["resample_icdf_64to192"], flush=True)
print(json.dumps(out))
