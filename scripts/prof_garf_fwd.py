"""Forward-only timing of the fused GARF radiance / proposal kernels with and without the stashes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
from nerf_experiments_b200.model_garf_proposal import ProposalNetwork

dev = th.device("cuda:0")
B, S = 4096, 192
th.manual_seed(0)
rad = RadianceNetwork(0.5, 1.5).to(dev)
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(dev)
t = th.sort(th.rand((B, S + 1), generator=g) * 5 + 2, dim=1).values.to(dev)
t0, t1 = t[:, :-1].contiguous(), t[:, 1:].contiguous()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def infer():
    with th.no_grad():
        rad.forward_rays(o, d, t0, t1)


out = {"rad_fwd_infer_ms": timeit(infer), "rad_fwd_train_ms": timeit(lambda: rad.forward_rays(o, d, t0, t1))}
n = B * S
out["infer_tflops"] = 2 * 596096 * n / out["rad_fwd_infer_ms"] / 1e9
out["cycles_per_tile_infer"] = out["rad_fwd_infer_ms"] * 1e-3 * 1.9e9 / (n / 128 / 148)
out["cycles_per_tile_train"] = out["rad_fwd_train_ms"] * 1e-3 * 1.9e9 / (n / 128 / 148)
print(json.dumps(out, indent=1))
