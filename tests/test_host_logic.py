"""Host-side logic that needs no GPU: module surface / state-dict parity with the reference,
tile-program compilation, weight-gradient scheduling, schedules, sharding (gloo, 2 ranks)."""
import math
import os

import numpy as np
import pytest
import torch as th

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _net(n_hidden=4, hidden=256, ddir=True, ddens=False, nseg=2, identity=True):
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    ep = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, identity, 1.0)
    ed = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, identity, 1.0)
    return arch.NerfModel(n_hidden, hidden, ddir, ddens, nseg, ep, ed)


@pytest.mark.parametrize("tag,kw", [
    ("std_small", dict(n_hidden=2, hidden=64, ddir=True, ddens=False, nseg=2)),
    ("nodelay", dict(n_hidden=1, hidden=64, ddir=False, ddens=True, nseg=2)),
    ("flat", dict(n_hidden=0, hidden=64, ddir=True, ddens=False, nseg=1))])
def test_state_dict_keys_and_seeded_init_match_reference(tag, kw):
    """Same constructor + same seed => the same parameters, key for key, as the reference's
    NerfModel (its state dict is stored in the golden fixture): checkpoints are interchangeable."""
    z = np.load(os.path.join(G, f"nerf_model_{tag}.npz"))
    ref_sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    th.manual_seed(7)
    net = _net(**kw)
    sd = net.state_dict()
    assert sorted(sd.keys()) == sorted(ref_sd.keys())
    for k, v in sd.items():
        if "alpha" in k:
            continue
        assert np.array_equal(v.numpy(), ref_sd[k]), k


def test_param_groups_and_optimizer_surface():
    from nerf_experiments_b200 import model_interpolation as mi
    net = _net(1, 64)
    m = mi.NerfInterpolation(2.0, 8.0, net, 32, "equidistant", -1.0, "middle")
    assert m.param_groups[0]["learning_rate_start"] == 5e-4
    cfg = m.configure_optimizers()
    assert isinstance(cfg["optimizer"], th.optim.Adam) and cfg["optimizer"].defaults["eps"] == 1e-5
    assert cfg["lr_scheduler"]["interval"] == "step"
    with pytest.raises(ValueError):
        m._get_t_query(None, None, "right")


def test_scheduler_le_nice_closed_form():
    from nerf_experiments_b200.model_interpolation import SchedulerLeNice
    p = th.nn.Parameter(th.zeros(1))
    q = th.nn.Parameter(th.zeros(1))
    opt = th.optim.Adam([{"params": [p], "lr": 5e-4}, {"params": [q], "lr": 1e-3}])
    s = SchedulerLeNice(opt, [5e-4, 1e-3], [1e-5, 1e-3], [100, 0])
    for step in range(1, 151):
        opt.step()
        s.step()
        # torch's LRScheduler takes one scheduler step at construction: _step_count = step + 1
        want = 5e-4 * math.exp(min(step + 1, 100) * (math.log(1e-5) - math.log(5e-4)) / 100)
        assert opt.param_groups[0]["lr"] == pytest.approx(want, rel=1e-12)
        assert opt.param_groups[1]["lr"] == 1e-3
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-5, rel=1e-9)


def test_alpha_schedule_matches_oracle():
    from nerf_experiments_b200 import positional_encodings as pe
    from oracle import ref_pe
    enc = pe.BarfPositionalEncoding(10, 0.0, 0.5, 2.5, True, 1.0)
    for e in (0.0, 0.5, 1.25, 2.0, 2.5, 3.0):
        enc.update_alpha(e)
        assert float(enc.alpha) == pytest.approx(ref_pe.barf_alpha(e, 10, 0.0, 0.5, 2.5), rel=1e-6)
        assert th.allclose(enc.compute_mask(enc.alpha), ref_pe.barf_mask(enc.alpha, 10), atol=1e-6)


def test_forward_program_of_the_standard_network():
    from nerf_experiments_b200.mlp_program import compile_backward, compile_forward
    net = _net()
    f = net.fused_field()
    cm = compile_forward(f.layers_fn(f.flat))
    assert cm.program.n_ops == 12 and (cm.program.n_slabs, cm.program.n_stages) == (5, 4)
    # every weight element lands in exactly one forward image
    covered = sum(c.n_rows * c.n_cols for c in cm.pack_chunks)
    assert covered == sum(p.numel() for n, p in net.named_parameters() if n.endswith("weight"))
    # every bias once, plus the 256 density weights the two-tile forward kernel reads as fp32
    assert cm.two_tile_ok and cm.density_w_off >= 0
    assert sum(b.n for b in cm.pack_biases) == 256 + sum(p.numel() for n, p in net.named_parameters() if n.endswith("bias"))
    assert len(cm.pack_chunks_k16) == sum(1 for c in cm.pack_chunks if c.rows_padded >= 16 and c.n_rows > 1)
    assert all(c.img_rows in (16, 128, 256) for c in cm.pack_chunks_k16)
    cb = compile_backward(cm, True, f._encoders())
    assert (cb.pos_grad_cols, cb.dir_grad_cols) == (64, 64)   # canonical encoding-gradient layout
    # the transposed images of an encoding-gradient op cover every (output, encoding column) once,
    # each at its canonical row: cos(c, j) -> 2*(10c+j), sin -> +1, identity -> 60..62
    rows = {}
    for c in cb.pack_chunks:
        if c.dst_row_step == 2 or c.dst_row0 == 60:
            rows.setdefault(c.dst_off, set()).update(c.dst_row0 + r * c.dst_row_step for r in range(c.n_rows))
    assert rows and all(len(r) in (63, 27) and max(r) <= 62 for r in rows.values())
    # weight-gradient units tile every weight exactly once
    total = sum(u.m_real * u.n_real for u in cb.units)
    assert total == covered
    # ... and every bias element is the column sum of exactly one unit's dY slabs
    bias_idx = sorted(u.bias_dst + m for u in cb.units if u.bias_dst >= 0 for m in range(u.m_real))
    assert bias_idx == sorted(
        i for n, p in net.named_parameters() if n.endswith("bias")
        for i in range(f.flat.offset_of(p), f.flat.offset_of(p) + p.numel()))


@pytest.mark.parametrize("n_tiles", [1, 7, 4096])
def test_wgrad_schedule_covers_every_tile_once(n_tiles):
    from nerf_experiments_b200.mlp_program import compile_backward, compile_forward, schedule_wgrad
    net = _net()
    f = net.fused_field()
    cb = compile_backward(compile_forward(f.layers_fn(f.flat)), False)
    items = schedule_wgrad(cb.units, n_tiles, 148)
    seen = {}
    for it in items:
        key = (it.dst, it.dy_slab, it.x_slab)
        seen.setdefault(key, []).append((it.tile_begin, it.tile_end))
    assert len(seen) == len(cb.units)
    for ranges in seen.values():
        ranges.sort()
        assert ranges[0][0] == 0 and ranges[-1][1] == n_tiles
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


def test_unsupported_widths_fail_loudly():
    net = _net(1, 320)
    f = net.fused_field()
    from nerf_experiments_b200.mlp_program import compile_forward
    with pytest.raises(RuntimeError, match="not supported"):
        compile_forward(f.layers_fn(f.flat))


def test_no_cpu_fallback():
    net = _net(1, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(th.zeros(4, 3), th.zeros(4, 3))


def test_shard_ranges_partition_the_batch():
    from nerf_experiments_b200.parallel import shard_range
    for n, w in ((4096, 8), (10, 4), (7, 8)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def _gloo_worker(rank, world, init, q):
    import torch.distributed as dist
    from nerf_experiments_b200.parallel import allreduce_sum_, global_mean_scale, shard_range
    from oracle import ref_render
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    # data-parallel gradient of a mean loss over rays == all-reduced shard gradients / world
    g = th.Generator().manual_seed(0)
    B, S = 16, 8
    sigma = th.nn.functional.softplus(th.randn((B, S), generator=g))
    w = th.randn((S,), generator=g).requires_grad_()
    delta = th.full((B, S), 0.1)
    col = th.rand((B, S, 3), generator=g)
    target = th.rand((B, 3), generator=g)
    b, e = shard_range(B, rank, world)
    rgb, _ = ref_render.render_rays(sigma[b:e] * w.abs(), col[b:e], delta[b:e])
    th.nn.functional.mse_loss(rgb, target[b:e]).backward()
    flat = w.grad.clone()
    allreduce_sum_(flat)
    flat *= global_mean_scale(world)
    q.put((rank, flat))
    dist.barrier()
    dist.destroy_process_group()


def _file_rendezvous(tmp_path, name):
    """file:// rendezvous: no TCP port to collide with a neighbour or a socket in TIME_WAIT"""
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    return f"file://{tmp_path / name}"


def test_gloo_two_ranks_gradient_allreduce_equals_full_batch(tmp_path):
    import torch.multiprocessing as mp
    from oracle import ref_render
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    init = _file_rendezvous(tmp_path, "rdzv_allreduce")
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, init, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = th.Generator().manual_seed(0)
    B, S = 16, 8
    sigma = th.nn.functional.softplus(th.randn((B, S), generator=g))
    w = th.randn((S,), generator=g).requires_grad_()
    delta = th.full((B, S), 0.1)
    col = th.rand((B, S, 3), generator=g)
    target = th.rand((B, 3), generator=g)
    rgb, _ = ref_render.render_rays(sigma * w.abs(), col, delta)
    th.nn.functional.mse_loss(rgb, target).backward()
    assert th.allclose(got[0], got[1])
    assert th.allclose(got[0], w.grad, rtol=1e-5, atol=1e-8)


def _sharded_render_worker(rank, world, init, out):
    import torch.distributed as dist
    from nerf_experiments_b200.parallel import render_rows_sharded
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    H, W = 7, 5                                     # 7 rows over 2 ranks: ragged blocks
    full = th.arange(H * W * 3, dtype=th.float32).view(H, W, 3)
    img = render_rows_sharded(lambda a, b: full[a:b].clone(), H, W, th.device("cpu"), None, 0)
    if rank == 0:
        out.put(bool(th.equal(img, full)))
    else:
        out.put(img is None)
    dist.destroy_process_group()


def test_sharded_render_assembles_the_rows_of_all_ranks(tmp_path):
    """2-rank gloo: every rank renders its block of rows, rank 0 gets the image, the others None."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    init = _file_rendezvous(tmp_path, "rdzv_render")
    procs = [ctx.Process(target=_sharded_render_worker, args=(r, 2, init, out)) for r in range(2)]
    for p_ in procs:
        p_.start()
    results = [out.get(timeout=120) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
    assert results == [True, True]


def test_exponential_lr_closed_form_matches_torch_scheduler():
    """The fused Adam evaluates GARF's schedule in closed form on the device (engine.exponential_lr /
    nerfb200_adam_step_dev); this pins the closed form to torch's ExponentialLR constructed the way
    the reference constructs it (last_epoch = -decay_end - 1, garf/model_garf.py:365-428)."""
    import math
    import torch as th
    from nerf_experiments_b200.engine import exponential_lr
    lr0, stop, n = 5e-4, 5e-5, 7
    gamma = 2 ** (math.log2(stop / lr0) / n)
    p = th.nn.Parameter(th.zeros(1))
    opt = th.optim.Adam([{"params": [p], "lr": lr0, "initial_lr": lr0}])
    sched = th.optim.lr_scheduler.ExponentialLR(opt, gamma=gamma, last_epoch=-n - 1)
    for step in range(1, 20):
        assert opt.param_groups[0]["lr"] == pytest.approx(exponential_lr(lr0, math.log(gamma), n, step), rel=1e-9)
        p.grad = th.ones(1)
        opt.step()
        sched.step()


def test_le_nice_closed_form_edge_cases():
    """lr = start * exp(logf * min(step, n)) with logf = 0 for the constant cases — including the
    reference's n = -1 default of CameraExtrinsics, which evaluates to the STOP rate."""
    from nerf_experiments_b200.model_interpolation import le_nice_lr, log_decay_factor
    assert le_nice_lr(1e-3, log_decay_factor(1e-3, 1e-5, 0), 0, 5) == 1e-3
    assert le_nice_lr(1e-3, log_decay_factor(1e-3, 1e-5, None), None, 5) == 1e-3
    assert le_nice_lr(1e-3, log_decay_factor(1e-3, 1e-5, -1), -1, 5) == pytest.approx(1e-5)
    assert le_nice_lr(1e-3, log_decay_factor(1e-3, 1e-5, 100), 100, 1000) == pytest.approx(1e-5)
