"""Per-kernel times of one GARF training step (garf/main.py shape) on one B200: CUDA events around the
fused kernels of both networks, the whole step eager and as a CUDA graph. usage: prof_garf.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from nerf_experiments_b200.model_garf import GarfModel, garf_engine

dev = th.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
th.manual_seed(1337)
m = GarfModel(2.0, 7.0, 64, 192, 0.5, 1.5, 1.0, 1e-3, 1e-4, 100000, 0.0, 1e-3, 1e-4, 100000, 0.0).to(dev)
m.train()
eng = garf_engine(m, dev)
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(dev)
tgt = th.rand((B, 3), generator=g).to(dev)
for _ in range(3):
    eng.step(o, d, tgt)
fields = {"prop": m.proposal_network.fused_field(), "rad": m.radiance_network.fused_field()}
for f in fields.values():
    f.timers = {}
th.cuda.synchronize()
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record()
N = 10
for _ in range(N):
    eng.step(o, d, tgt)
e1.record()
th.cuda.synchronize()
out = {"B": B, "ms_per_step_eager": e0.elapsed_time(e1) / N}
for name, f in fields.items():
    cg = f.compiled
    n = B * (64 if name == "prop" else 192)
    for k, evs in f.timers.items():
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        flops = 2 * cg.macs_per_sample * n
        n_tiles = (n + 127) // 128
        y, z, dy = (cg.fwd.y_slabs_per_tile * 16384 * n_tiles, cg.fwd.z_slabs_per_tile * 16384 * n_tiles,
                    cg.bwd.y_slabs_per_tile * 16384 * n_tiles)
        byt = {"garf_fwd_train": y + z, "garf_bwd": z + dy, "garf_wgrad": y + dy + z}[k]
        out[f"{name}.{k}"] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "stash_GBps": round(byt / ms / 1e6, 1),
                              "stash_MB": round(byt / 1e6, 1)}
        if k == "garf_wgrad":     # what the kernel actually streams: every unit's dY + X + Z slabs (dY slabs are re-read per X block)
            rd = sum(u.n_dy_slabs + u.n_x_slabs + u.n_z_slabs for u in cg.units) * 16384 * n_tiles
            out[f"{name}.{k}"].update(read_MB=round(rd / 1e6, 1), read_GBps=round(rd / ms / 1e6, 1))
    f.timers = None
eng.capture(o, d, tgt)
for _ in range(3):
    eng.replay(o, d, tgt)
th.cuda.synchronize()
e0.record()
for _ in range(N):
    eng.replay(o, d, tgt)
e1.record()
th.cuda.synchronize()
out["ms_per_step_graph"] = e0.elapsed_time(e1) / N
out["rays_per_s_graph"] = B / out["ms_per_step_graph"] * 1e3
out["launches_per_step"] = int(eng.launches_per_replay)
print(json.dumps(out, indent=1))
