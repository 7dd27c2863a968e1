"""Two ranks on two GPUs over NCCL: a sharded step (eager and as a captured CUDA graph that contains
the all-reduce) equals the single-GPU step on the concatenated batch. Needs two
devices: skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
import os
import tempfile

import pytest
import torch as th

pytestmark = pytest.mark.gpu


def _build_and_batches(dev, n_steps, B):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_step import _build, _rays
    model, cam = _build(dev, True, 0, 32, seed=11, sampling="equidistant", offset=0.0)
    batches = [tuple(t.to(dev) for t in _rays(B, 5, 200 + s)) for s in range(n_steps)]
    return model, batches


def _rank_main(rank, world, init_file, out_file, n_steps, B):
    import faulthandler
    import torch.distributed as dist
    from nerf_experiments_b200.engine import TrainEngine
    from nerf_experiments_b200.parallel import shard_range
    faulthandler.dump_traceback_later(90, exit=True)      # a hang becomes a stack dump and a failure
    dev = th.device("cuda", rank)
    th.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    model, batches = _build_and_batches(dev, n_steps, B)
    if rank == 1:                                       # replicas that start apart must be pulled together
        with th.no_grad():
            for p in model.parameters():
                p.add_(0.01)
    eng = TrainEngine(model, dev)
    b0, b1 = shard_range(B, rank, world)
    shard = lambda batch: tuple(t[b0:b1].contiguous() for t in batch)
    losses = [eng.step(*shard(batches[0])), eng.step(*shard(batches[1]))]
    eng.capture(*shard(batches[1]))
    for b in batches[2:]:
        losses.append(eng.replay(*shard(b)).clone())
    mean_losses = th.stack([l.reshape(()) for l in losses])
    dist.all_reduce(mean_losses)
    flat_all = [th.empty_like(eng.flat.flat) for _ in range(world)]
    dist.all_gather(flat_all, eng.flat.flat)
    if rank == 0:
        th.save({"flat": eng.flat.flat.cpu(), "flat_other": flat_all[1].cpu(), "losses": (mean_losses / world).cpu(),
                 "steps": eng.state[:2].tolist()}, out_file)
    eng.release_graph()            # a graph holding NCCL collectives would block the teardown
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_step_on_the_concatenated_batch():
    if th.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from nerf_experiments_b200.engine import TrainEngine
    n_steps, B = 5, 128
    dev = th.device("cuda", 0)
    model, batches = _build_and_batches(dev, n_steps, B)
    eng = TrainEngine(model, dev)
    ref_losses = [float(eng.step(*b)) for b in batches]
    ref_flat = eng.flat.flat.cpu()
    with tempfile.TemporaryDirectory() as d:
        init_file, out_file = os.path.join(d, "rdzv"), os.path.join(d, "out.pt")
        mp.spawn(_rank_main, args=(2, init_file, out_file, n_steps, B), nprocs=2, join=True)
        out = th.load(out_file)
    assert th.equal(out["flat"], out["flat_other"])                 # replicas stay bit-identical
    assert out["steps"] == [n_steps, 0]
    assert out["losses"].tolist() == pytest.approx(ref_losses, rel=1e-4)
    diff = (out["flat"] - ref_flat).abs()
    assert float((diff > 2e-5).float().mean()) < 1e-3 and float(diff.max()) < 5e-3
