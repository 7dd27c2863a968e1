"""End-to-end parity of the render module + pose refinement + training engine on the GPU
against the CPU oracle step (oracle/ref_step.py)."""
import pytest
import torch as th

from oracle import ref_step

pytestmark = pytest.mark.gpu


def _build(cuda, identity, n_prop, n_rad, seed=0, n_images=5, sampling="equidistant", offset=-1.0):
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.model_camera_extrinsics import CameraExtrinsics
    th.manual_seed(seed)

    def net():
        ep = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, identity, 1.0)
        ed = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, identity, 1.0)
        m = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 1000)
        ep.alpha.fill_(7.25); ed.alpha.fill_(4.0)
        return m

    rad = net()
    prop = net() if n_prop > 0 else None
    model = mi.NerfInterpolation(2.0, 8.0, rad, n_rad, sampling, offset, "middle", prop, n_prop)
    cam = CameraExtrinsics(n_images, 1e-3, 1e-5, 1000)
    with th.no_grad():
        cam.rotation.copy_(th.randn(n_images, 3) * 0.05)
        cam.translation.copy_(th.randn(n_images, 3) * 0.05)
    model.camera_extrinsics = cam
    model.param_groups = model.param_groups + cam.param_groups
    return model.to(cuda), cam


def _rays(B, n_images, seed):
    g = th.Generator().manual_seed(seed)
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    target = th.rand((B, 3), generator=g)
    idx = th.randint(0, n_images, (B,), generator=g).int()
    pw = th.full((B, 1), 1 / 555.0)
    return o, d, target, idx, pw


def _oracle(model, cam, o, d, target, idx, uniforms, n_prop, n_rad, identity, emulate):
    sd_r = {k: v.detach().cpu().clone().requires_grad_(v.dim() > 0) for k, v in model.model_radiance.state_dict().items()}
    sd_p = None
    if n_prop > 0:
        sd_p = {k: v.detach().cpu().clone().requires_grad_(v.dim() > 0) for k, v in model.model_proposal.state_dict().items()}
    rot = cam.rotation.detach().cpu().clone().requires_grad_()
    tr = cam.translation.detach().cpu().clone().requires_grad_()
    cfg = dict(n_hidden=4, n_segments=2, delayed_direction=True, delayed_density=False)
    pe_cfg = dict(pos_levels=10, dir_levels=4, scale=1.0, identity=identity, alpha_pos=th.tensor(7.25), alpha_dir=th.tensor(4.0))
    loss, rgb = ref_step.barf_step(sd_r, cfg, pe_cfg, rot, tr, idx, o, d, target, 2.0, 8.0, n_rad, "middle", uniforms,
                                   sd_p, n_prop, "equidistant", -1.0, emulate)
    loss.backward()
    return loss.detach(), rgb.detach(), sd_r, sd_p, rot.grad, tr.grad


def _rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("identity,n_prop,n_rad,B", [(True, 0, 128, 96), (False, 64, 256, 40)])
def test_render_step_matches_oracle(cuda, identity, n_prop, n_rad, B):
    model, cam = _build(cuda, identity, n_prop, n_rad)
    o, d, target, idx, pw = _rays(B, 5, 1)
    # same uniforms on both sides: the module draws from torch's CUDA generator
    th.manual_seed(123)
    off = th.rand((B, 1), device=cuda)
    uniforms = {"offset": off.cpu()}
    th.manual_seed(123)
    o2, d2, _, _ = cam(idx.to(cuda), o.to(cuda), d.to(cuda))
    fine, coarse = model(o2, d2, pw.to(cuda))
    loss = th.nn.functional.mse_loss(fine, target.to(cuda))
    if coarse is not None:
        loss = loss + th.nn.functional.mse_loss(coarse, target.to(cuda))
    loss.backward()

    for emulate, tol_rgb, tol_g in ((False, 1e-2, 0.3), (True, 4e-3, 5e-2)):
        l_ref, rgb_ref, sd_r, sd_p, g_rot, g_tr = _oracle(model, cam, o, d, target, idx, uniforms, n_prop, n_rad, identity, emulate)
        assert (fine.detach().cpu() - rgb_ref).abs().max() < tol_rgb          # north_star: 1e-2 abs vs fp32
        assert abs(loss.item() - l_ref.item()) < tol_rgb
        for name, p in model.model_radiance.named_parameters():
            assert _rel(p.grad.cpu(), sd_r[name].grad) < tol_g, (emulate, name)
        if n_prop > 0:
            for name, p in model.model_proposal.named_parameters():
                assert _rel(p.grad.cpu(), sd_p[name].grad) < tol_g, (emulate, "proposal", name)
        assert _rel(cam.translation.grad.cpu(), g_tr) < tol_g
        assert _rel(cam.rotation.grad.cpu(), g_rot) < tol_g


def test_engine_matches_torch_adam(cuda):
    """TrainEngine (flat buffer, grad sink, fused Adam, closed-form LR) == autograd + torch Adam
    + SchedulerLeNice on the same module, step for step."""
    from nerf_experiments_b200.engine import TrainEngine
    B = 64
    batches = [_rays(B, 5, 10 + s) for s in range(3)]
    losses = {}
    finals = {}
    for mode in ("engine", "torch"):
        model, cam = _build(cuda, True, 0, 64, seed=4)
        if mode == "engine":
            eng = TrainEngine(model, cuda)
        else:
            cfg = model.configure_optimizers()
            opt, sched = cfg["optimizer"], cfg["lr_scheduler"]["scheduler"]
        th.manual_seed(77)
        ls = []
        for (o, d, target, idx, pw) in batches:
            if mode == "engine":
                ls.append(eng.step(o.to(cuda), d.to(cuda), target.to(cuda), idx.to(cuda), pw.to(cuda)).item())
            else:
                opt.zero_grad()
                o2, d2, _, _ = cam(idx.to(cuda), o.to(cuda), d.to(cuda))
                fine, _ = model(o2, d2, pw.to(cuda))
                loss = th.nn.functional.mse_loss(fine, target.to(cuda))
                loss.backward()
                opt.step(); sched.step()
                ls.append(loss.item())
        losses[mode] = ls
        finals[mode] = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    assert losses["engine"] == pytest.approx(losses["torch"], rel=2e-3, abs=1e-5)
    for k in finals["torch"]:
        a, b = finals["engine"][k], finals["torch"][k]
        assert (a - b).abs().max() <= 2e-4 * max(1.0, b.abs().max().item()) + 2e-4, k


def test_host_stepper_matches_direct_steps(cuda):
    """HostStepper (copy stream + deferred loss read-back) gives the losses and parameters of plain
    engine.step calls on the same batches, one call late."""
    from nerf_experiments_b200.engine import HostStepper, TrainEngine
    B = 96
    batches = [_rays(B, 5, 30 + s) for s in range(7)]
    results = {}
    for mode in ("direct", "host", "host_graph"):
        # (no random sampling offsets: the graph and the eager steps must see the same rays)
        model, cam = _build(cuda, True, 0, 32, seed=9, sampling="equidistant", offset=0.0)
        eng = TrainEngine(model, cuda)
        th.manual_seed(5)
        if mode == "direct":
            ls = [eng.step(*(t.to(cuda) for t in (o, d, target, idx, pw))).item() for (o, d, target, idx, pw) in batches]
        else:
            # host_graph: run-ahead of two steps, the captured graph fed by ONE packed copy per step
            st = HostStepper(eng) if mode == "host" else HostStepper(eng, depth=2, use_graph=True)
            ls = []
            for (o, d, target, idx, pw) in batches:
                prev = st.submit(tuple(t.pin_memory() for t in (o, d, target, idx, pw)))
                if prev is not None:
                    ls.append(prev)
            ls += st.drain()
            assert st.h2d_bytes == sum(t.numel() * t.element_size() for t in batches[0])
            if mode == "host_graph":
                assert eng._graph is not None and st._packed is not None
        results[mode] = (ls, eng.flat.flat.detach().clone())
    for mode in ("host", "host_graph"):
        assert len(results[mode][0]) == 7
        assert results["direct"][0] == pytest.approx(results[mode][0], rel=1e-4)   # atomics: summation order varies
        # gradients are flushed with floating-point atomics: equal up to summation order, which Adam's
        # normalisation can blow up to a full lr-sized step for the odd parameter with a near-zero gradient
        diff = (results["direct"][1] - results[mode][1]).abs()
        assert float((diff > 2e-5).float().mean()) < 1e-3 and float(diff.max()) < 5 * 1e-3


@pytest.mark.parametrize("B", [40, 300])
def test_two_tile_forward_matches_one_tile_kernel(cuda, B):
    """The experimental two-tiles-in-flight forward kernel (nerfb200_mlp_fwd2: per-K-step weight
    images in the SWIZZLE_32B layout, density column evaluated by the epilogue) against the
    default kernel: same outputs, and — through the shared stash / sign-bit layout — the same
    gradients from the (unchanged) backward kernels.  Odd tile counts and several tile pairs per
    CTA are covered by the two batch sizes."""
    o, d, target, idx, pw = _rays(B, 5, 3)
    outs = {}
    for two_tile in (False, True):
        model, cam = _build(cuda, True, 0, 64, seed=2)
        field = model.model_radiance.fused_field()
        field.use_two_tile = two_tile
        th.manual_seed(11)
        o2, d2, _, _ = cam(idx.to(cuda), o.to(cuda), d.to(cuda))
        fine, _ = model(o2, d2, pw.to(cuda))
        loss = th.nn.functional.mse_loss(fine, target.to(cuda))
        loss.backward()
        assert (field.k16_units >= 0) and field.compiled.two_tile_ok
        outs[two_tile] = (fine.detach().clone(), [p.grad.detach().clone() for p in model.model_radiance.parameters()],
                          cam.translation.grad.detach().clone())
    assert (outs[True][0] - outs[False][0]).abs().max() < 2e-3          # density: fp32 dot vs MMA accumulation order
    for ga, gb in zip(outs[True][1], outs[False][1]):
        assert (ga - gb).norm() <= 2e-2 * gb.norm() + 1e-7
    assert (outs[True][2] - outs[False][2]).norm() <= 2e-2 * outs[False][2].norm() + 1e-7


def test_checkpoint_round_trip_and_torch_adam_compatibility(cuda, tmp_path):
    """engine.save_checkpoint / load_checkpoint in the layout Lightning writes for the reference:
    (1) 3 steps + save + load into a fresh engine + 2 steps == 5 uninterrupted steps;
    (2) the optimizer state loads into the torch.optim.Adam + SchedulerLeNice of
    configure_optimizers() (what the reference resumes with) and one torch step from it equals one
    engine step."""
    from nerf_experiments_b200.engine import TrainEngine
    B = 64
    batches = [tuple(t.to(cuda) for t in _rays(B, 5, 50 + s)) for s in range(5)]

    def fresh():
        model, cam = _build(cuda, True, 0, 32, seed=6)
        return model, cam, TrainEngine(model, cuda)

    model_a, _, eng_a = fresh()
    th.manual_seed(1)
    for b in batches:
        eng_a.step(*b)
    model_b, _, eng_b = fresh()
    th.manual_seed(1)
    for b in batches[:3]:
        eng_b.step(*b)
    path = str(tmp_path / "ckpt_epoch=00.ckpt")
    eng_b.save_checkpoint(path, epoch=0)
    ck = th.load(path, map_location="cpu", weights_only=False)
    assert ck["global_step"] == 3 and set(ck["state_dict"]) == set(model_b.state_dict())
    assert any(k.startswith("model_radiance.model_segments.0.0.") for k in ck["state_dict"])
    assert {"camera_extrinsics.rotation", "camera_extrinsics.translation"} <= set(ck["state_dict"])
    rng = th.cuda.get_rng_state(cuda)
    model_c, cam_c, eng_c = fresh()
    eng_c.load_checkpoint(path)
    assert eng_c.step_count == 3
    th.cuda.set_rng_state(rng, cuda)
    for b in batches[3:]:
        eng_c.step(*b)
    diff = (eng_c.flat.flat - eng_a.flat.flat).abs()
    assert float((diff > 2e-5).float().mean()) < 1e-3 and float(diff.max()) < 5e-3

    # (2) torch Adam + SchedulerLeNice resume from the same file
    model_d, cam_d = _build(cuda, True, 0, 32, seed=6)      # no engine: plain autograd gradients
    model_d.load_state_dict(ck["state_dict"])
    cfg = model_d.configure_optimizers()
    opt, sched = cfg["optimizer"], cfg["lr_scheduler"]["scheduler"]
    opt.load_state_dict(ck["optimizer_states"][0])
    sched.load_state_dict(ck["lr_schedulers"][0])
    model_e, cam_e, eng_e = fresh()
    eng_e.load_checkpoint(ck)
    # the moments and the step count arrive intact on both sides
    for pd, pe_ in zip([p for g in opt.param_groups for p in g["params"]],
                       [p for g in model_e.param_groups for p in g["parameters"]]):
        off = eng_e.flat.offset_of(pe_)
        assert float(opt.state[pd]["step"]) == 3.0
        assert th.equal(opt.state[pd]["exp_avg"].reshape(-1).to(cuda), eng_e.exp_avg[off:off + pe_.numel()])
        assert th.equal(opt.state[pd]["exp_avg_sq"].reshape(-1).to(cuda), eng_e.exp_avg_sq[off:off + pe_.numel()])
    o, d, target, idx, pw = batches[3]
    th.manual_seed(9)
    opt.zero_grad()
    o2, d2, _, _ = cam_d(idx, o, d)
    fine, _ = model_d(o2, d2, pw)
    th.nn.functional.mse_loss(fine, target).backward()
    opt.step()
    sched.step()
    th.manual_seed(9)
    eng_e.step(o, d, target, idx, pw)
    # one step from them is the engine's step: a parameter moves by at most lr = 5e-4, the two
    # paths agree on all but the odd parameter whose near-zero gradient is dominated by summation order
    bad = total = 0
    for (n1, p1), (n2, p2) in zip(model_d.named_parameters(), model_e.named_parameters()):
        assert n1 == n2
        bad += int(((p1 - p2).abs() > 5e-5).sum())
        total += p1.numel()
        assert (p1 - p2).abs().max() < 6e-4, n1
    assert bad < 2e-3 * total
    assert opt.param_groups[0]["lr"] == pytest.approx(eng_e.learning_rates(5)[0], rel=1e-6)


def test_captured_step_matches_eager_steps(cuda):
    """TrainEngine.capture / replay (the whole optimisation step as one CUDA graph, Adam's schedule read
    from device memory) == eager step(): same losses and parameters over several steps whose learning
    rates differ (no random sampling offsets in this configuration, so the two runs see the same rays)."""
    from nerf_experiments_b200.engine import TrainEngine
    B = 96
    batches = [tuple(t.to(cuda) for t in _rays(B, 5, 70 + s)) for s in range(6)]
    out = {}
    for mode in ("eager", "graph"):
        model, cam = _build(cuda, True, 0, 32, seed=8, sampling="equidistant", offset=0.0)
        eng = TrainEngine(model, cuda)
        losses = [float(eng.step(*batches[0]))]
        if mode == "graph":
            eng.capture(*batches[0])
            assert eng.launches_per_replay >= 7          # pack, pose, sampling, field fwd/bwd/wgrad, compositing, Adam ...
            for b in batches[1:]:
                losses.append(float(eng.replay(*b)))
        else:
            for b in batches[1:]:
                losses.append(float(eng.step(*b)))
        out[mode] = (losses, eng.flat.flat.detach().clone(), eng.step_count)
    assert out["eager"][2] == out["graph"][2] == 6
    assert out["eager"][0] == pytest.approx(out["graph"][0], rel=1e-4)
    diff = (out["eager"][1] - out["graph"][1]).abs()
    assert float((diff > 2e-5).float().mean()) < 1e-3 and float(diff.max()) < 5e-3


def test_two_tile_entry_point_rejects_programs_it_cannot_run(cuda):
    """nerfb200_mlp_fwd2 validates the tile program: a 6-slab program (direction encoded next to the
    position, not delayed) comes back as an argument error with a message, not as a launch."""
    import ctypes as C
    from nerf_experiments_b200 import _lib
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.fused_mlp import make_inputs
    th.manual_seed(0)
    net = arch.NerfModel(2, 256, False, False, 1, pe.BarfPositionalEncoding(10, 10.0, 0.0, 0.0, True, 1.0),
                         pe.BarfPositionalEncoding(4, 4.0, 0.0, 0.0, True, 1.0)).to(cuda)
    f = net.fused_field()
    cm = f.prepare(cuda)
    assert cm.program.n_slabs == 6 and not cm.two_tile_ok and f.k16_units < 0
    n = 256
    pos = th.randn((n, 3), device=cuda)
    inputs = make_inputs(n, 1, 0, pos=pos, dir=pos, pixel_width_per_sample=True)
    sigma, rgb = th.empty(n, device=cuda), th.empty((n, 3), device=cuda)
    cp, cd = f.pe_cfgs()
    rc = _lib.lib().nerfb200_mlp_fwd2(C.byref(cm.program), f.wpack.data_ptr(), f.bias.data_ptr(), C.byref(inputs),
                                      C.byref(cp), C.byref(cd), f.pe_pos.alpha_tensor().data_ptr(),
                                      f.pe_dir.alpha_tensor().data_ptr(), 0.0, sigma.data_ptr(), rgb.data_ptr(),
                                      None, None, cm.bias_floats, cm.density_w_off, th.cuda.current_stream().cuda_stream)
    assert rc != 0 and b"mlp_fwd2" in _lib.lib().nerfb200_last_error()
    # the default entry point runs the same program
    s2, r2 = net(pos, pos, None, None, None)
    assert th.isfinite(s2).all() and th.isfinite(r2).all()


def test_non_finite_loss_makes_the_step_a_no_op(cuda):
    """The reference replaces a NaN loss by a fresh leaf, so the optimiser moves nothing
    (barf/model_interpolation.py:522-524). The engine applies the same rule on the device: parameters and
    both Adam moments stay bit-identical, the schedules advance, Adam's per-parameter step does not."""
    from nerf_experiments_b200.engine import TrainEngine
    B = 64
    model, cam = _build(cuda, True, 0, 32, seed=3)
    eng = TrainEngine(model, cuda)
    good = [tuple(t.to(cuda) for t in _rays(B, 5, 40 + s)) for s in range(3)]
    eng.step(*good[0])
    snap = [t.clone() for t in (eng.flat.flat, eng.exp_avg, eng.exp_avg_sq)]
    o, d, target, idx, pw = [t.clone() for t in good[1]]
    o[7, 1] = float("nan")                                  # one poisoned ray
    loss = eng.step(o, d, target, idx, pw)
    assert th.isnan(loss)
    for a, b in zip(snap, (eng.flat.flat, eng.exp_avg, eng.exp_avg_sq)):
        assert th.equal(a, b)
    assert eng.step_count == 2 and eng.skipped_steps() == 1 and eng.state[:2].tolist() == [2, 1]
    # an out-of-range image index poisons its ray instead of reading / scattering out of bounds
    bad_idx = idx.clone(); bad_idx[3] = 99
    assert th.isnan(eng.step(good[1][0], good[1][1], target, bad_idx, pw))
    for a, b in zip(snap, (eng.flat.flat, eng.exp_avg, eng.exp_avg_sq)):
        assert th.equal(a, b)
    # the next good step is the SECOND Adam step of every parameter, at the FOURTH scheduler step
    ck_before = eng.checkpoint()
    assert float(ck_before["optimizer_states"][0]["state"][0]["step"]) == 1.0
    assert ck_before["lr_schedulers"][0]["last_epoch"] == 3
    assert th.isfinite(eng.step(*good[2]))
    assert not th.equal(snap[0], eng.flat.flat) and th.isfinite(eng.flat.flat).all()
    # reference arithmetic of that step: torch Adam at step 2 with the lr of scheduler step 4
    lr = eng.learning_rates(4)[0]
    g = eng.grad[:1000].clone()
    m = snap[1][:1000] + (g - snap[1][:1000]) * (1 - 0.9)
    v = snap[2][:1000] * 0.999 + (1 - 0.999) * g * g
    expect = snap[0][:1000] - (lr / (1 - 0.9 ** 2)) * (m / (v.sqrt() / (1 - 0.999 ** 2) ** 0.5 + 1e-5))
    assert (eng.flat.flat[:1000] - expect).abs().max() < 1e-7


def test_barf_training_loss_captured_matches_step_helper(cuda):
    """BarfModel: the capturable device part (`training_loss` after `update_schedules`) equals the
    reference-shaped `_step_helper` (loss, alpha, pose error), and an engine driving it through a CUDA
    graph equals the eager engine."""
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200 import scene
    from nerf_experiments_b200.engine import TrainEngine
    from nerf_experiments_b200.model_camera_calibration import BarfModel, LoopState

    sc = scene.make_scene(6, 32, 32, cuda, rotation_noise=0.1, translation_noise=0.1, blur_sigmas=(4.0, 1.0, 0.0))
    n_batches = 50

    def build():
        th.manual_seed(2)
        ep = pe.BarfPositionalEncoding(10, 0.0, 0.2, 0.8, True, 1.0)
        ed = pe.BarfPositionalEncoding(4, 0.0, 0.2, 0.8, True, 1.0)
        net = arch.NerfModel(4, 256, True, False, 2, ep, ed, 5e-4, 1e-5, 1000)
        m = BarfModel(n_training_images=6, camera_learning_rate_start=1e-3, camera_learning_rate_stop=1e-5,
                      camera_learning_rate_decay_end=1000, near_sphere_normalized=2.0, far_sphere_normalized=8.0,
                      model_radiance=net, samples_per_ray_radiance=32, max_gaussian_sigma=4.0,
                      uniform_sampling_strategy="equidistant", uniform_sampling_offset_size=0.0).to(cuda)
        m.loop = LoopState(sc.batcher, n_batches)
        return m

    g = th.Generator(device=cuda).manual_seed(0)
    idx = th.randint(0, len(sc.batcher), (8, 128), device=cuda, generator=g)
    batches = [sc.batcher.batch(idx[i]) for i in range(8)]

    m = build()
    step = 20                                  # epoch 0.4: mask partly open, blur level between two pyramid levels
    m.update_schedules(step)
    loss_dev, logs = m.training_loss(*batches[0])
    loss_ref = m._step_helper(batches[0], step, "train")
    assert float(loss_dev) == pytest.approx(float(loss_ref), rel=1e-6)
    assert float(logs["pose_error"]) == pytest.approx(float(m._logged["pose_error"]), rel=1e-6)
    assert float(logs["alpha"]) == pytest.approx(float(m._logged["alpha"]))

    out = {}
    for mode in ("eager", "graph"):
        m = build()
        eng = TrainEngine(m, cuda, loss_fn=m.training_loss)
        losses = [float(eng.step(*batches[0]))]
        if mode == "graph":
            eng.capture(*batches[0])
        for b in batches[1:]:
            losses.append(float(eng.replay(*b) if mode == "graph" else eng.step(*b)))
        out[mode] = (losses, eng.flat.flat.detach().clone(), float(eng.last_logs["pose_error"]))
    assert out["eager"][0] == pytest.approx(out["graph"][0], rel=1e-4)
    assert out["eager"][2] == pytest.approx(out["graph"][2], rel=1e-3)
    diff = (out["eager"][1] - out["graph"][1]).abs()
    assert float((diff > 2e-5).float().mean()) < 1e-3 and float(diff.max()) < 5e-3


@pytest.mark.parametrize("n_prop,n_rad,sampling", [(0, 128, "equidistant"), (0, 96, "stratified_uniform"), (64, 256, "equidistant")])
def test_render_matches_forward_and_oracle_compositing(cuda, n_prop, n_rad, sampling):
    """The north-star inner op `render(rays_o, rays_d) -> rgb, depth, weights` (ONE C-ABI call per pass):
    rgb equals `forward` bit for bit on the same uniforms, weights / depth equal the oracle's compositing of
    the field's own outputs within 1e-5 (north_star: fp32 compositing 1e-5 abs), rgb within 1e-2 of the
    fp32 oracle render."""
    from oracle import ref_render
    model, cam = _build(cuda, True, n_prop, n_rad, seed=2, sampling=sampling)
    B = 70
    o, d, target, idx, pw = (t.to(cuda) for t in _rays(B, 5, 3))
    th.manual_seed(42)
    with th.no_grad():
        fine, _ = model(o, d, pw)
    th.manual_seed(42)
    rgb, depth, w = model.render(o, d, pw)
    assert rgb.shape == (B, 3) and depth.shape == (B,) and w.shape == (B, n_rad)
    assert th.equal(rgb, fine)
    assert float(w.min()) >= 0 and float(w.sum(1).max()) <= 1 + 1e-5
    assert float(depth.min()) >= 2.0 - 0.1 and float(depth.max()) <= 8.0
    if n_prop == 0 and sampling == "equidistant":
        # recompute the pass in the open: bins -> field -> oracle compositing
        from nerf_experiments_b200 import ops
        from nerf_experiments_b200.field_function import field_rays
        th.manual_seed(42)
        off = th.rand((B, 1), device=cuda)
        t0, t1 = ops.sample_uniform(2.0, 8.0, B, n_rad, cuda, None, off, -1.0)
        with th.no_grad():
            sigma, col = field_rays(model.model_radiance, o, d, t0, t1, pw, "middle")
        r_rgb, r_w = ref_render.render_rays(sigma.cpu(), col.cpu(), (t1 - t0).cpu())
        assert (rgb.cpu() - r_rgb).abs().max() < 1e-5 and (w.cpu() - r_w).abs().max() < 1e-5
        r_depth = (r_w * ((t0 + t1) / 2).cpu()).sum(1) / r_w.sum(1).clamp_min(th.finfo(th.float32).eps)
        assert (depth.cpu() - r_depth).abs().max() < 1e-4
