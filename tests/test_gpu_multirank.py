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


def _garf_model_and_batches(dev, n_steps, B, n_img=6):
    from nerf_experiments_b200.model_garf_camera_calibration import CameraCalibrationModel
    th.manual_seed(5)
    m = CameraCalibrationModel(n_img, 1e-3, 1e-5, 40, 10, 2.0, 7.0, 16, 32, 0.5, 1.5, 2.0,
                               1e-3, 1e-4, 50, 0.0, 2e-3, 1e-4, 60, 0.0).to(dev)
    m.train()
    gen = th.Generator().manual_seed(3)
    batches = []
    for _ in range(n_steps):
        o = th.nn.functional.normalize(th.randn((B, 3), generator=gen), dim=1) * 4.0
        d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=gen), dim=1)
        o_n = o + 0.05 * th.randn((B, 3), generator=gen)
        d_n = th.nn.functional.normalize(d + 0.05 * th.randn((B, 3), generator=gen), dim=1)
        idx = th.randint(0, n_img, (B,), generator=gen)
        batches.append(tuple(x.to(dev) for x in (o, o_n, d, d_n, th.rand((B, 3), generator=gen), idx,
                                                 th.rand(B, generator=gen), th.rand(B, generator=gen))))
    return m, batches


def _garf_rank_main(rank, world, init_file, out_file, n_steps, B):
    import faulthandler
    import torch.distributed as dist
    from nerf_experiments_b200.model_garf import garf_engine
    from nerf_experiments_b200.parallel import shard_range
    faulthandler.dump_traceback_later(90, exit=True)
    dev = th.device("cuda", rank)
    th.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    m, batches = _garf_model_and_batches(dev, n_steps, B)
    eng = garf_engine(m, dev)
    b0, b1 = shard_range(B, rank, world)
    shard = lambda batch: tuple(t[b0:b1].contiguous() for t in batch)
    eng.step(*shard(batches[0]))
    eng.step(*shard(batches[1]))
    eng.capture(*shard(batches[1]))
    for b in batches[2:]:
        eng.replay(*shard(b))
    flat_all = [th.empty_like(eng.flat.flat) for _ in range(world)]
    dist.all_gather(flat_all, eng.flat.flat)
    if rank == 0:
        th.save({"flat": eng.flat.flat.cpu(), "flat_other": flat_all[1].cpu(), "steps": eng.state[:2].tolist()}, out_file)
    eng.release_graph()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_garf_pose_step_equals_single_rank_step():
    """GARF + pose refinement (five parameter groups, both fused networks, PropNet chain) sharded over two
    ranks — eager steps, then the captured graph with the NCCL all-reduce inside — against the single-GPU
    steps on the concatenated batches with the same uniforms."""
    if th.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from nerf_experiments_b200.model_garf import garf_engine
    n_steps, B = 5, 256
    dev = th.device("cuda", 0)
    m, batches = _garf_model_and_batches(dev, n_steps, B)
    eng = garf_engine(m, dev)
    for b in batches:
        eng.step(*b)
    ref_flat = eng.flat.flat.cpu()
    with tempfile.TemporaryDirectory() as d:
        init_file, out_file = os.path.join(d, "rdzv"), os.path.join(d, "out.pt")
        mp.spawn(_garf_rank_main, args=(2, init_file, out_file, n_steps, B), nprocs=2, join=True)
        out = th.load(out_file)
    assert th.equal(out["flat"], out["flat_other"])                 # replicas stay bit-identical
    assert out["steps"] == [n_steps, 0]
    diff = (out["flat"] - ref_flat).abs()
    # floating-point atomics: equal up to summation order, which Adam's normalisation can blow up to an
    # lr-sized step for the odd parameter with a near-zero gradient (as in the single-GPU engine tests)
    assert float((diff > 5e-5).float().mean()) < 2e-3 and float(diff.max()) < 2e-2
