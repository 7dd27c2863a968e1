"""800 x 800 render sharded over the ranks (run under torchrun): every rank renders a block of image
rows with the fused no_grad path, rank 0 assembles the image; checks it against rank 0's own
single-GPU render and reports rays/s.  torchrun --nproc-per-node N scripts/bench_render_sharded.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.distributed as dist
import bench
from nerf_experiments_b200.ray_batcher import render_image, render_image_sharded

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = th.device("cuda", local)
th.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
model = bench.build_model(20).to(dev)            # same seed on every rank: replicated weights
H = W = 800
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((H * W, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((H * W, 3), generator=g), dim=1).to(dev)


def run():
    th.manual_seed(3)                            # the sampling offsets: same stream on every rank ...
    return render_image_sharded(model, o, d, H, W, 1 / 555.0)


img = run()
th.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    img = run()
e1.record()
th.cuda.synchronize()
ms = th.tensor([e0.elapsed_time(e1) / 3], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    assert img.shape == (H, W, 3) and bool(th.isfinite(img).all())
    # ... but a different one per row block than in a single-GPU render: compare with equidistant offsets off
    model.uniform_sampling_offset_size = 0.0
    ref = render_image(model, o, d, H, W, 1 / 555.0)
if world > 1:
    model.uniform_sampling_offset_size = 0.0
    img0 = render_image_sharded(model, o, d, H, W, 1 / 555.0)
else:
    img0 = render_image(model, o, d, H, W, 1 / 555.0) if rank == 0 else None
if rank == 0:
    err = float((img0 - ref).abs().max())
    print(json.dumps({"n_gpus": world, "ms_per_image": round(float(ms), 3), "rays_per_s": round(H * W / float(ms) * 1e3),
                      "max_abs_diff_vs_single_gpu": err}))
    assert err == 0.0
if world > 1:
    dist.destroy_process_group()
