"""GPU parity of the GARF / SARF / Gabor path: the activation kernels (through the C ABI)
against the fixtures of the unmodified reference, the GARF networks against the reference's
outputs and gradients, and the GARF model chain (inverse-CDF sampling, nerfacc-flavour
compositing, proposal loss) against the oracle restatement with identical uniforms."""
import os

import numpy as np
import pytest
import torch as th

from oracle import ref_garf

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "garf.npz")
THIN = 37


def _g():
    z = np.load(G)
    return {k: th.from_numpy(z[k]) for k in z.files}


def _thin(t):
    return t if t.numel() <= 8192 else t.flatten()[::THIN]


@pytest.mark.parametrize("name", ["gauss", "sarf", "gabor"])
def test_activation_kernels_match_reference(cuda, name):
    from nerf_experiments_b200 import _lib, ops
    g = _g()
    kind = {"gauss": _lib.ACT_GAUSS, "sarf": _lib.ACT_SARF, "gabor": _lib.ACT_GABOR}[name]
    keys = {"gauss": ("gauss_p",), "sarf": ("sarf_p",), "gabor": ("gabor_p0", "gabor_p1")}[name]
    x = g["act_x"].to(cuda).requires_grad_()
    ps = [g[k].to(cuda).requires_grad_() for k in keys]
    y = ops.activation(kind, x, *ps)
    assert th.allclose(y.cpu(), g[name + "_y"], rtol=2e-6, atol=2e-7)
    grads = th.autograd.grad(y, [x] + ps, g["act_up"].to(cuda))
    assert th.allclose(grads[0].cpu(), g[name + "_dx"], rtol=2e-5, atol=2e-6)
    gk = [name + "_dp"] if len(ps) == 1 else [name + "_dp0", name + "_dp1"]
    for gr, k in zip(grads[1:], gk):
        assert th.allclose(gr.cpu(), g[k], rtol=2e-4, atol=2e-5), k


@pytest.mark.parametrize("N,F", [(1, 1), (5, 3), (1000, 129), (4096, 1024), (70000, 256)])
def test_activation_kernels_shapes(cuda, N, F):
    """Ragged shapes (F not a multiple of the block, one row, many row chunks) against the oracle."""
    from nerf_experiments_b200 import _lib, ops
    gen = th.Generator().manual_seed(N + F)
    x = th.randn((N, F), generator=gen).requires_grad_()
    p = (th.rand(F, generator=gen) + 0.5).requires_grad_()
    up = th.randn((N, F), generator=gen)
    y = ref_garf.gauss_act(x, p)
    rx, rp = th.autograd.grad(y, (x, p), up)
    xc, pc = x.detach().to(cuda).requires_grad_(), p.detach().to(cuda).requires_grad_()
    yc = ops.activation(_lib.ACT_GAUSS, xc, pc)
    gx, gp = th.autograd.grad(yc, (xc, pc), up.to(cuda))
    assert th.allclose(yc.cpu(), y, rtol=2e-6, atol=2e-7)
    assert th.allclose(gx.cpu(), rx, rtol=2e-5, atol=2e-6)
    # the parameter gradient is a sum over N rows in a different order
    assert th.allclose(gp.cpu(), rp, rtol=1e-3, atol=1e-4 * max(1.0, float(N) ** 0.5))


def _seeded_nets(cuda):
    from nerf_experiments_b200.model_garf_proposal import ProposalNetwork
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    th.manual_seed(77)
    prop = ProposalNetwork(0.5, 1.5)
    rad = RadianceNetwork(0.5, 1.5)
    return prop.to(cuda), rad.to(cuda)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_garf_networks_match_reference(cuda):
    """The fused GARF kernels against the outputs / gradients of the UNMODIFIED reference modules
    (tests/golden/garf.npz, fp32). Arithmetic of the fused path: fp32 first layer and raw-coordinate
    terms, bf16 operands with fp32 accumulation elsewhere, bf16 stashes. Stated tolerances: rgb 1e-2
    absolute (north_star's bf16-MLP bound), density 2 % of its range, every parameter gradient 4 %
    relative L2 (measured on B200: see the assertion messages / profiles)."""
    g = _g()
    prop, rad = _seeded_nets(cuda)

    def check_grads(net, prefix):
        worst = 0.0
        for n, p in net.named_parameters():
            ref = g[prefix + n]
            got = _thin(p.grad).cpu()
            err = _rel(got, ref)
            worst = max(worst, err)
            assert err < 4e-2, (n, err)
        return worst

    pos, dirs = g["net_pos"].to(cuda), g["net_dir"].to(cuda)
    rgb, dens = rad(pos, dirs)
    assert (rgb.cpu() - g["rad_rgb"]).abs().max() < 1e-2
    assert (dens.cpu() - g["rad_density"]).abs().max() < 2e-2 * (1 + g["rad_density"].abs().max())
    ((rgb * g["up_rgb"].to(cuda)).sum() + (dens * g["up_density"].to(cuda)).sum()).backward()
    w_rad = check_grads(rad, "rad.grad.")
    sp = prop(pos)
    assert sp.shape == (pos.shape[0], 1)
    assert (sp.cpu() - g["prop_sigma"]).abs().max() < 2e-2 * (1 + g["prop_sigma"].abs().max())
    (sp * g["up_prop"].to(cuda)).sum().backward()
    w_prop = check_grads(prop, "prop.grad.")
    print(f"worst relative L2 gradient error: radiance {w_rad:.3e}, proposal {w_prop:.3e}")


@pytest.mark.parametrize("N", [1, 127, 128, 300, 5000])
def test_garf_kernels_match_the_program_interpreter(cuda, N):
    """The CUDA kernels against the CPU interpreter of the same tile programs (tests/garf_sim.py: same
    roundings), ragged sizes included: outputs to 2e-3, parameter and input gradients to 1.5 % relative
    L2 (summation order and exp2 approximations are what is left)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import garf_sim
    prop, rad = _seeded_nets(cuda)
    gen = th.Generator().manual_seed(N)
    pos = th.randn((N, 3), generator=gen) * 1.2
    dirs = th.nn.functional.normalize(th.randn((N, 3), generator=gen), dim=1)
    up_s, up_c = th.randn(N, generator=gen), th.randn((N, 3), generator=gen)
    for net, has_dir in ((rad, True), (prop, False)):
        f = net.fused_field()
        f.prepare(cuda)
        flat = f.flat.flat.detach().cpu()
        pc = pos.to(cuda).requires_grad_()
        dc = dirs.to(cuda).requires_grad_()
        if has_dir:
            rgb, dens = net(pc, dc)
            loss = (rgb * up_c.to(cuda)).sum() + (dens * up_s.to(cuda)).sum()
        else:
            dens = net(pc)[:, 0]
            rgb = None
            loss = (dens * up_s.to(cuda)).sum()
        loss.backward()
        s_sigma, s_rgb, s_grad, s_dpos, s_ddir = garf_sim.run_network(f.compiled, flat, pos, dirs if has_dir else None,
                                                                    up_s, up_c if has_dir else None)
        assert (dens.detach().cpu() - s_sigma).abs().max() < 2e-3 * (1 + s_sigma.abs().max())
        if has_dir:
            assert (rgb.detach().cpu() - s_rgb).abs().max() < 2e-3
        for p in net.parameters():
            o = f.flat.offset_of(p)
            ref = s_grad[o: o + p.numel()].view(p.shape)
            assert _rel(p.grad.cpu(), ref) < 1.5e-2 or float((p.grad.cpu() - ref).abs().max()) < 1e-6, (has_dir, tuple(p.shape))
        assert _rel(pc.grad.cpu(), s_dpos) < 1.5e-2
        if has_dir:
            assert _rel(dc.grad.cpu(), s_ddir) < 1.5e-2


def test_garf_rays_mode_equals_samples_mode(cuda):
    """forward_rays (positions o + (t0 + t1) / 2 d formed in registers) == forward on materialised
    positions (GarfModel._get_positions, garf/model_garf.py:105), values and ray gradients."""
    prop, rad = _seeded_nets(cuda)
    B, S = 37, 20
    gen = th.Generator().manual_seed(9)
    o = (th.nn.functional.normalize(th.randn((B, 3), generator=gen), dim=1) * 4.0).to(cuda).requires_grad_()
    d = th.nn.functional.normalize(-o.detach().cpu() + 0.3 * th.randn((B, 3), generator=gen), dim=1).to(cuda).requires_grad_()
    t = th.sort(th.rand((B, S + 1), generator=gen) * 5 + 2, dim=1).values.to(cuda)
    t0, t1 = t[:, :-1].contiguous(), t[:, 1:].contiguous()
    up = th.randn((B, S, 3), generator=gen).to(cuda)
    rgb_r, dens_r = rad.forward_rays(o, d, t0, t1)
    ((rgb_r * up).sum() + dens_r.sum()).backward()
    go_r, gd_r = o.grad.clone(), d.grad.clone()
    gw_r = rad.model_density_1[2].weight.grad.clone()
    o.grad = d.grad = None
    rad.zero_grad()
    pos = o[:, None] + d[:, None] * (t0 + t1)[..., None] / 2
    rgb_s, dens_s = rad(pos.view(-1, 3), d.repeat_interleave(S, dim=0))
    ((rgb_s.view(B, S, 3) * up).sum() + dens_s.sum()).backward()
    assert th.equal(rgb_r.reshape(-1, 3), rgb_s) and th.equal(dens_r.reshape(-1), dens_s)
    assert _rel(go_r, o.grad) < 1e-4 and _rel(gd_r, d.grad) < 1e-4
    assert _rel(gw_r, rad.model_density_1[2].weight.grad) < 1e-4
    sp = prop.forward_rays(o, d, t0, t1)
    assert th.equal(sp.reshape(-1, 1), prop(pos.view(-1, 3).detach()))


def test_garf_networks_refuse_cpu_tensors():
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    net = RadianceNetwork(0.5, 1.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(th.zeros(4, 3), th.zeros(4, 3))


@pytest.mark.parametrize("training", [True, False])
def test_garf_model_chain_matches_oracle(cuda, training):
    from nerf_experiments_b200.model_garf import GarfModel
    th.manual_seed(3)
    m = GarfModel(2.0, 7.0, 32, 48, 0.5, 1.5, 1.0, 1e-3, 1e-4, 100, 0.0, 1e-3, 1e-4, 100, 0.0).to(cuda)
    m.train(training)
    B = 24
    gen = th.Generator().manual_seed(11)
    o = th.nn.functional.normalize(th.randn((B, 3), generator=gen), dim=1) * 4.0
    d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=gen), dim=1)
    target = th.rand((B, 3), generator=gen)
    u = (th.rand(B, generator=gen), th.rand(B, generator=gen)) if training else None
    sd_p = {k: v.detach().cpu().clone().requires_grad_() for k, v in m.proposal_network.state_dict().items()}
    sd_r = {k: v.detach().cpu().clone().requires_grad_() for k, v in m.radiance_network.state_dict().items()}
    o_r, d_r = o.clone().requires_grad_(), d.clone().requires_grad_()      # the pose refinement's path: d(loss) / d(rays)
    r_rgb, r_op, r_dp, r_lp, (r0, r1) = ref_garf.garf_forward(
        sd_p, sd_r, o_r, d_r, 2.0, 7.0, 32, 48, None if u is None else u[0], None if u is None else u[1])
    uc = None if u is None else (u[0].to(cuda), u[1].to(cuda))
    o_c, d_c = o.to(cuda).requires_grad_(), d.to(cuda).requires_grad_()
    rgb, (lp, lr) = m._forward_loss((o_c, d_c, target.to(cuda)), uc)
    _, opacity, depth, extras = m(o.to(cuda), d.to(cuda), uc)
    # The oracle runs the networks in fp32, the fused kernels with bf16 operands: the proposal densities
    # differ by ~1 %, so the resampled intervals (a continuous function of the proposal cdf) move by a
    # fraction of a bin, and everything downstream inherits that. Stated bounds:
    assert th.allclose(extras["t_starts"].cpu(), r0, rtol=0, atol=3e-2)
    assert th.allclose(rgb.cpu(), r_rgb, atol=1e-2)                      # north_star: bf16-MLP rgb 1e-2 abs
    assert th.allclose(opacity[:, 0].cpu(), r_op, atol=1e-2)
    assert th.allclose(depth[:, 0].cpu(), r_dp, atol=5e-2)
    assert float(lp) == pytest.approx(float(r_lp), rel=0.15, abs=1e-6)
    # gradients of the summed loss reach both networks like in the reference's manual optimisation
    (lp + lr).backward()
    r_loss = r_lp + th.nn.functional.mse_loss(r_rgb, target)
    r_loss.backward()
    for n, p in m.radiance_network.named_parameters():
        assert _rel(p.grad.cpu(), sd_r[n].grad) < 0.25, n               # bf16 operands vs fp32 (as the ReLU network)
    for n, p in m.proposal_network.named_parameters():
        assert _rel(p.grad.cpu(), sd_p[n].grad) < 0.35, n
    # ray gradients (radiance and proposal-loss paths together; sampling is not differentiated on either side)
    assert _rel(o_c.grad.cpu(), o_r.grad) < 0.3, _rel(o_c.grad.cpu(), o_r.grad)
    assert _rel(d_c.grad.cpu(), d_r.grad) < 0.3, _rel(d_c.grad.cpu(), d_r.grad)


@pytest.mark.parametrize("B,S", [(1, 1), (7, 33), (64, 64), (300, 192)])
def test_propnet_chain_kernels_match_oracle(cuda, B, S):
    """lindisp intervals, transmittance -> cdf (forward and backward) and the proposal loss with its
    gradient, each one launch, against the oracle restatement and its autograd (oracle/ref_garf.py)."""
    from nerf_experiments_b200 import ops
    gen = th.Generator().manual_seed(B * 1000 + S)
    s_edges = th.sort(th.rand((B, S + 1), generator=gen), dim=1).values
    s_edges[:, 0], s_edges[:, -1] = 0.0, 1.0
    t, t0, t1, delta, mid = ops.lindisp_intervals(s_edges.to(cuda), 2.0, 7.0, want_mid=True)
    t_ref = 1.0 / (s_edges / 7.0 + (1.0 - s_edges) / 2.0)
    assert th.allclose(t.cpu(), t_ref, rtol=2e-6) and th.equal(t0, t[:, :-1]) and th.equal(t1, t[:, 1:])
    assert th.allclose(delta.cpu(), t_ref[:, 1:] - t_ref[:, :-1], rtol=1e-4, atol=1e-6)
    assert th.allclose(mid.cpu(), (t_ref[:, 1:] + t_ref[:, :-1]) / 2, rtol=2e-6)

    sigma = (th.nn.functional.softplus(th.randn((B, S), generator=gen)) * 2).requires_grad_()
    r0, r1 = t_ref[:, :-1], t_ref[:, 1:]
    trans_ref, cdf_ref = ref_garf.transmittance_cdf(sigma, r0, r1)
    up = th.randn((B, S + 1), generator=gen)
    (cdf_ref * up).sum().backward()
    sc = sigma.detach().to(cuda).requires_grad_()
    cdf = ops.transmittance_cdf(sc, r0.to(cuda), r1.to(cuda))
    (cdf * up.to(cuda)).sum().backward()
    assert th.allclose(cdf.detach().cpu(), cdf_ref.detach(), atol=2e-6)
    assert th.allclose(sc.grad.cpu(), sigma.grad, rtol=1e-4, atol=1e-6)
    tr, cdf2 = ops.transmittance(sc.detach(), r0.to(cuda), r1.to(cuda))
    assert th.allclose(tr.cpu(), trans_ref.detach(), atol=2e-6) and th.equal(cdf2, cdf.detach())

    # proposal loss: a coarser key histogram (every other edge, perturbed cdf) against the fine query one
    Sk = max(S // 2, 1)
    idx = th.linspace(0, S, Sk + 1).round().long()
    t_k = t_ref[:, idx]
    cdf_k = (cdf_ref.detach()[:, idx] * (0.7 + 0.6 * th.rand((B, Sk + 1), generator=gen))).clamp(0, 1)
    cdf_k = th.sort(cdf_k, dim=1).values.requires_grad_()
    loss_ref = ref_garf.pdf_outer_loss(t_ref, cdf_ref.detach(), t_k, cdf_k).mean()
    loss_ref.backward()
    ck = cdf_k.detach().to(cuda).requires_grad_()
    loss = ops.proposal_loss(t_ref.to(cuda), cdf_ref.detach().to(cuda), t_k.to(cuda), ck)
    (loss * 3.0).backward()
    assert float(loss) == pytest.approx(float(loss_ref), rel=1e-4, abs=1e-9)
    assert th.allclose(ck.grad.cpu() / 3.0, cdf_k.grad, rtol=1e-4, atol=1e-8)


def test_garf_engine_matches_torch_optimisers(cuda):
    """garf_engine (one flat buffer, fused Adam with the four ExponentialLR groups, captured graph) ==
    the reference-shaped training_step with its two torch Adam optimisers and schedulers, step for step."""
    from nerf_experiments_b200.model_garf import GarfModel, garf_engine
    B = 192
    gen = th.Generator().manual_seed(2)
    batches = []
    for _ in range(5):
        o = th.nn.functional.normalize(th.randn((B, 3), generator=gen), dim=1) * 4.0
        d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=gen), dim=1)
        batches.append(tuple(x.to(cuda) for x in (o, d, th.rand((B, 3), generator=gen), th.rand(B, generator=gen),
                                                  th.rand(B, generator=gen))))
    out = {}
    for mode in ("torch", "engine", "graph"):
        th.manual_seed(5)
        m = GarfModel(2.0, 7.0, 16, 32, 0.5, 1.5, 2.0, 1e-3, 1e-4, 50, 0.0, 2e-3, 1e-4, 60, 0.0).to(cuda)
        m.train()
        losses = []
        if mode == "torch":
            for i, (o, d, c, u0, u1) in enumerate(batches):
                losses.append(float(m.training_step((o, d, c), i, u_rays=(u0, u1))))
            flat = th.cat([p.detach().reshape(-1) for g in m.param_groups for p in g["parameters"]])
        else:
            eng = garf_engine(m, cuda)
            eng.step(*batches[0])
            losses.append(float(eng.last_logs["loss_fine"] + eng.last_logs["train_proposal_loss"]))
            if mode == "graph":
                eng.capture(*batches[0])
            for b in batches[1:]:
                (eng.replay if mode == "graph" else eng.step)(*b)
                losses.append(float(eng.last_logs["loss_fine"] + eng.last_logs["train_proposal_loss"]))
            flat = eng.flat.flat.detach().clone()
        out[mode] = (losses, flat)
    for mode in ("engine", "graph"):
        assert out[mode][0] == pytest.approx(out["torch"][0], rel=2e-3)
        diff = (out[mode][1] - out["torch"][1]).abs()
        assert float((diff > 5e-5).float().mean()) < 2e-3 and float(diff.max()) < 2e-2


def test_garf_training_step_reduces_loss(cuda):
    from nerf_experiments_b200.model_garf import GarfModel
    th.manual_seed(5)
    m = GarfModel(2.0, 7.0, 16, 32, 0.5, 1.5, 1.0, 1e-3, 1e-4, 1000, 0.0, 1e-3, 1e-4, 1000, 0.0).to(cuda)
    m.train()
    gen = th.Generator().manual_seed(1)
    o = (th.nn.functional.normalize(th.randn((256, 3), generator=gen), dim=1) * 4.0).to(cuda)
    d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((256, 3), generator=gen), dim=1).to(cuda)
    target = th.full((256, 3), 0.25, device=cuda)
    losses = [float(m.training_step((o, d, target), i)) for i in range(30)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])


def test_garf_camera_calibration_engine_matches_torch_optimisers(cuda):
    """garf/model_camera_calibration.py: 6-tuple batches, poses refined in front of the fused field; the
    engine's single fused Adam over five groups == the reference-shaped step with three optimisers."""
    from nerf_experiments_b200.model_garf import garf_engine
    from nerf_experiments_b200.model_garf_camera_calibration import CameraCalibrationModel
    B, n_img = 192, 6
    gen = th.Generator().manual_seed(3)
    batches = []
    for _ in range(5):
        o = th.nn.functional.normalize(th.randn((B, 3), generator=gen), dim=1) * 4.0
        d = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=gen), dim=1)
        o_n = o + 0.05 * th.randn((B, 3), generator=gen)
        d_n = th.nn.functional.normalize(d + 0.05 * th.randn((B, 3), generator=gen), dim=1)
        idx = th.randint(0, n_img, (B,), generator=gen)
        batches.append(tuple(x.to(cuda) for x in (o, o_n, d, d_n, th.rand((B, 3), generator=gen), idx,
                                                  th.rand(B, generator=gen), th.rand(B, generator=gen))))
    out = {}
    for mode in ("torch", "engine", "graph"):
        th.manual_seed(5)
        m = CameraCalibrationModel(n_img, 1e-3, 1e-5, 40, 10, 2.0, 7.0, 16, 32, 0.5, 1.5, 2.0,
                                   1e-3, 1e-4, 50, 0.0, 2e-3, 1e-4, 60, 0.0).to(cuda)
        m.train()
        assert len(m.param_groups) == 5
        losses = []
        if mode == "torch":
            for i, b in enumerate(batches):
                losses.append(float(m.training_step(b[:6], i, u_rays=(b[6], b[7]))))
            flat = th.cat([p.detach().reshape(-1) for g in m.param_groups for p in g["parameters"]])
        else:
            eng = garf_engine(m, cuda)
            eng.step(*batches[0])
            losses.append(float(eng.last_logs["loss_fine"] + eng.last_logs["train_proposal_loss"]))
            if mode == "graph":
                eng.capture(*batches[0])
            for b in batches[1:]:
                (eng.replay if mode == "graph" else eng.step)(*b)
                losses.append(float(eng.last_logs["loss_fine"] + eng.last_logs["train_proposal_loss"]))
            flat = eng.flat.flat.detach().clone()
            eng.release_graph()
        out[mode] = (losses, flat)
        poses = th.cat([p.detach().reshape(-1) for p in m.camera_extrinsics.parameters()])
        assert float(poses.abs().max()) > 1e-4            # the poses moved
    for mode in ("engine", "graph"):
        assert out[mode][0] == pytest.approx(out["torch"][0], rel=2e-3)
        diff = (out[mode][1] - out["torch"][1]).abs()
        assert float((diff > 5e-5).float().mean()) < 2e-3 and float(diff.max()) < 2e-2
        # the pose group itself (last 36 floats): Adam's first steps move every element by ~lr
        assert th.allclose(out[mode][1][-6 * n_img:], out["torch"][1][-6 * n_img:], atol=3e-4)


def test_garf_input_gradients_match_oracle_autograd(cuda):
    """d(position) / d(direction) of the fused kernels — what the pose refinement of
    garf/model_camera_calibration.py trains on — against autograd through the fp32 oracle networks
    (oracle/ref_garf.py, pinned by the reference goldens): 5 % relative L2 (bf16 operands; the raw-coordinate
    paths are fp32 in both)."""
    prop, rad = _seeded_nets(cuda)
    N = 700
    gen = th.Generator().manual_seed(21)
    pos = th.randn((N, 3), generator=gen) * 1.2
    dirs = th.nn.functional.normalize(th.randn((N, 3), generator=gen), dim=1)
    up_s, up_c = th.randn(N, generator=gen), th.randn((N, 3), generator=gen)
    for net, has_dir in ((rad, True), (prop, False)):
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        pr, dr = pos.clone().requires_grad_(), dirs.clone().requires_grad_()
        pc, dc = pos.to(cuda).requires_grad_(), dirs.to(cuda).requires_grad_()
        if has_dir:
            r_rgb, r_dens = ref_garf.radiance_network(sd, pr, dr)
            ((r_rgb * up_c).sum() + (r_dens * up_s).sum()).backward()
            rgb, dens = net(pc, dc)
            ((rgb * up_c.to(cuda)).sum() + (dens * up_s.to(cuda)).sum()).backward()
            assert _rel(dc.grad.cpu(), dr.grad) < 5e-2, _rel(dc.grad.cpu(), dr.grad)
        else:
            (ref_garf.proposal_network(sd, pr)[:, 0] * up_s).sum().backward()
            (net(pc)[:, 0] * up_s.to(cuda)).sum().backward()
        assert _rel(pc.grad.cpu(), pr.grad) < 5e-2, _rel(pc.grad.cpu(), pr.grad)
