"""GPU parity of the memory-bound kernels (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch as th

from oracle import ref_nerfacc, ref_pose, ref_render, ref_resample, ref_sampling

pytestmark = pytest.mark.gpu


def _ops():
    from nerf_experiments_b200 import ops
    return ops


@pytest.mark.parametrize("B,S,jit,off", [(7, 64, False, -1.0), (33, 128, True, -1.0), (5, 13, True, 0.0),
                                         (4, 1, False, 0.5), (1024, 256, True, -1.0)])
def test_sample_uniform_bit_exact(cuda, B, S, jit, off):
    g = th.Generator().manual_seed(B * 1000 + S)
    jitter = th.rand((B, S), generator=g) if jit else None
    u = th.rand((B, 1), generator=g)
    rs, re = ref_sampling.sample_uniform(2.0, 8.0, B, S, jitter, u, off)
    ts, te = _ops().sample_uniform(2.0, 8.0, B, S, cuda, None if jitter is None else jitter.to(cuda),
                                   u.to(cuda), off)
    assert th.equal(ts.cpu(), rs) and th.equal(te.cpu(), re)


@pytest.mark.parametrize("B,S", [(1, 1), (3, 7), (64, 64), (257, 128), (40, 192), (9, 256), (5, 516), (3, 1022)])
def test_composite_fwd_bwd(cuda, B, S):
    g = th.Generator().manual_seed(S)
    sigma = th.nn.functional.softplus(th.randn((B, S), generator=g) * 2).requires_grad_()
    delta = (th.rand((B, S), generator=g) * 0.1 + 0.01)
    rgb = th.rand((B, S, 3), generator=g).requires_grad_()
    out, w = ref_render.render_rays(sigma, rgb, delta)
    g_rgb = th.randn((B, 3), generator=g)
    g_w = th.randn((B, S), generator=g) * 0.1
    (out * g_rgb).sum().backward(retain_graph=True)
    ds_a, dc_a = sigma.grad.clone(), rgb.grad.clone()
    sigma.grad = None; rgb.grad = None
    ((out * g_rgb).sum() + (w * g_w).sum()).backward()
    ds_b, dc_b = sigma.grad.clone(), rgb.grad.clone()

    ops = _ops()
    sg, dl, cg = sigma.detach().to(cuda), delta.to(cuda), rgb.detach().to(cuda)
    o, wg, _, _ = ops.composite_fwd(sg, dl, cg)
    # north_star: fp32 compositing within 1e-5 abs
    assert (o.cpu() - out.detach()).abs().max() < 1e-5
    assert (wg.cpu() - w.detach()).abs().max() < 1e-5
    d_s, d_c = ops.composite_bwd(sg, dl, cg, g_rgb.to(cuda))
    assert (d_s.cpu() - ds_a).abs().max() < 2e-5 * max(1.0, ds_a.abs().max().item())
    assert (d_c.cpu() - dc_a).abs().max() < 1e-5
    d_s, d_c = ops.composite_bwd(sg, dl, cg, g_rgb.to(cuda), g_w.to(cuda))
    assert (d_s.cpu() - ds_b).abs().max() < 2e-5 * max(1.0, ds_b.abs().max().item())
    assert (d_c.cpu() - dc_b).abs().max() < 1e-5


def test_composite_autograd_function(cuda):
    g = th.Generator().manual_seed(3)
    B, S = 50, 128
    sigma = th.nn.functional.softplus(th.randn((B, S), generator=g)).to(cuda).requires_grad_()
    rgb = th.rand((B, S, 3), generator=g).to(cuda).requires_grad_()
    delta = th.full((B, S), 6 / 128).to(cuda)
    out, w = _ops().render_rays(sigma, rgb, delta)
    (out.sum() + (w ** 2).sum()).backward()
    s2 = sigma.detach().cpu().requires_grad_(); c2 = rgb.detach().cpu().requires_grad_()
    o2, w2 = ref_render.render_rays(s2, c2, delta.cpu())
    (o2.sum() + (w2 ** 2).sum()).backward()
    assert (sigma.grad.cpu() - s2.grad).abs().max() < 2e-5
    assert (rgb.grad.cpu() - c2.grad).abs().max() < 1e-5


@pytest.mark.parametrize("B,S", [(4, 64), (31, 192), (6, 50)])
def test_composite_nerfacc_flavour(cuda, B, S):
    from nerf_experiments_b200 import _lib
    g = th.Generator().manual_seed(S + 1)
    sigma = th.nn.functional.softplus(th.randn((B, S), generator=g)).requires_grad_()
    rgb = th.rand((B, S, 3), generator=g).requires_grad_()
    edges = th.sort(th.rand((B, S + 1), generator=g) * 5 + 2, dim=1).values
    t0, t1 = edges[:, :-1].contiguous(), edges[:, 1:].contiguous()
    o, op, dp, w, tr, al = ref_render.render_rays_nerfacc(sigma, rgb, t0, t1)
    g_rgb = th.randn((B, 3), generator=g); g_o = th.randn((B,), generator=g); g_d = th.randn((B,), generator=g)
    ((o * g_rgb).sum() + (op * g_o).sum() + (dp * g_d).sum()).backward()
    ops = _ops()
    tm = ((t0 + t1) / 2).to(cuda)
    args = (sigma.detach().to(cuda), (t1 - t0).to(cuda), rgb.detach().to(cuda))
    r_o, r_w, r_op, r_dp = ops.composite_fwd(*args, t_mid=tm, flavour=_lib.COMPOSITE_NERFACC,
                                              want_opacity=True, want_depth=True)
    assert (r_o.cpu() - o.detach()).abs().max() < 1e-5
    assert (r_op.cpu() - op.detach()).abs().max() < 1e-5
    assert (r_dp.cpu() - dp.detach()).abs().max() < 1e-4
    d_s, d_c = ops.composite_bwd(*args, g_rgb.to(cuda), None, tm, g_o.to(cuda), g_d.to(cuda),
                                 flavour=_lib.COMPOSITE_NERFACC)
    assert (d_s.cpu() - sigma.grad).abs().max() < 1e-4 * max(1.0, sigma.grad.abs().max().item())
    assert (d_c.cpu() - rgb.grad).abs().max() < 1e-5


def _resample_inputs(B, Sc, seed, zero_frac=0.3):
    g = th.Generator().manual_seed(seed)
    ts, te = ref_sampling.sample_uniform(2.0, 8.0, B, Sc, None, th.rand((B, 1), generator=g), -1.0)
    w = th.rand((B, Sc), generator=g) ** 4
    w = w * (th.rand((B, Sc), generator=g) > zero_frac)   # exact zeros => remainder ties
    w[:, 0] += 1e-3
    return ts, te - ts, w


@pytest.mark.parametrize("B,Sc,Sf", [(64, 64, 256), (33, 64, 128), (5, 4, 16), (17, 100, 300), (3, 1, 9), (2, 256, 1024)])
def test_resample_alloc_bit_exact(cuda, B, Sc, Sf):
    ts, dl, w = _resample_inputs(B, Sc, Sc + Sf)
    r_s, r_e, r_n, failed = ref_resample.sample_pdf_weighted(ts.numpy(), w.numpy(), dl.numpy(), Sf, 2.0, 8.0)
    assert not failed
    t0, t1, cnt, flag = _ops().resample_alloc(ts.to(cuda), w.to(cuda), dl.to(cuda), Sf, 2.0, 8.0, want_counts=True)
    assert int(flag.item()) == 0
    assert np.array_equal(cnt.cpu().numpy(), r_n)
    assert np.array_equal(t0.cpu().numpy(), r_s) and np.array_equal(t1.cpu().numpy(), r_e)
    assert int(cnt.sum(dim=1).min()) == Sf and int(cnt.sum(dim=1).max()) == Sf


def test_resample_alloc_known_answer(cuda):
    # SURVEY §8c probe: w=[.1,.6,.2,.05], 4 -> 16 gives counts [2,9,3,2]
    w = th.tensor([[0.1, 0.6, 0.2, 0.05]])
    ts = th.tensor([[2.0, 3.5, 5.0, 6.5]]); dl = th.full((1, 4), 1.5)
    t0, t1, cnt, flag = _ops().resample_alloc(ts.to(cuda), w.to(cuda), dl.to(cuda), 16, 2.0, 8.0, want_counts=True)
    assert cnt.cpu().tolist() == [[2, 9, 3, 2]]


def test_resample_alloc_fallback(cuda):
    # a ray with all-zero weights makes weights/sum NaN: the reference falls back, for the WHOLE
    # batch, to equidistant samples with offset -1 (barf/model_interpolation.py:273-275)
    B, Sc, Sf = 8, 64, 256
    ts, dl, w = _resample_inputs(B, Sc, 5)
    w[3] = 0.0
    u = th.rand((B,), generator=th.Generator().manual_seed(9))
    r_s, r_e, _, failed = ref_resample.sample_pdf_weighted(ts.numpy(), w.numpy(), dl.numpy(), Sf, 2.0, 8.0, u.numpy())
    assert failed
    t0, t1, cnt, flag = _ops().resample_alloc(ts.to(cuda), w.to(cuda), dl.to(cuda), Sf, 2.0, 8.0,
                                              fallback_u=u.to(cuda), want_counts=True)
    assert int(flag.item()) == 1
    assert np.array_equal(t0.cpu().numpy(), r_s) and np.array_equal(t1.cpu().numpy(), r_e)


@pytest.mark.parametrize("B,Sc,Sf,strat", [(16, 1, 64, True), (40, 64, 192, True), (9, 64, 192, False), (5, 33, 7, True)])
def test_resample_icdf_bit_exact(cuda, B, Sc, Sf, strat):
    g = th.Generator().manual_seed(Sc * 7 + Sf)
    edges = th.linspace(0, 1, Sc + 1).repeat(B, 1)
    pdf = th.rand((B, Sc), generator=g) ** 3 * (th.rand((B, Sc), generator=g) > 0.2)
    pdf[:, 0] += 1e-3
    cdf = th.cat((th.zeros(B, 1), th.cumsum(pdf, 1)), 1)
    cdf = cdf / cdf[:, -1:]
    u = th.rand((B,), generator=g) if strat else None
    r_e, r_i = ref_nerfacc.importance_sampling(edges.numpy(), cdf.numpy(), Sf, None if u is None else u.numpy())
    e, i = _ops().resample_icdf(edges.to(cuda), cdf.to(cuda), Sf, None if u is None else u.to(cuda), want_idx=True)
    assert np.array_equal(i.cpu().numpy(), r_i)
    assert np.array_equal(e.cpu().numpy(), r_e)


def test_pose_forward_backward(cuda):
    g = th.Generator().manual_seed(11)
    n_img, B = 37, 500
    rot = (th.randn((n_img, 3), generator=g) * 0.4); rot[0] = 0.0; rot[1] = th.tensor([1e-4, -2e-4, 5e-5])
    tr = th.randn((n_img, 3), generator=g) * 0.2
    idx = th.randint(0, n_img, (B,), generator=g)
    o = th.randn((B, 3), generator=g); d = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1)
    rot_r, tr_r = rot.clone().requires_grad_(), tr.clone().requires_grad_()
    no, nd, R, t = ref_pose.pose_forward(rot_r, tr_r, idx, o, d)
    go, gd = th.randn((B, 3), generator=g), th.randn((B, 3), generator=g)
    ((no * go).sum() + (nd * gd).sum()).backward()

    ops = _ops()
    rot_g, tr_g = rot.to(cuda).requires_grad_(), tr.to(cuda).requires_grad_()
    o2, d2, R2, t2 = ops.pose_forward(rot_g, tr_g, idx.to(cuda), o.to(cuda), d.to(cuda))
    assert (o2.cpu() - no.detach()).abs().max() < 1e-6
    assert (d2.cpu() - nd.detach()).abs().max() < 1e-5
    assert (R2.cpu() - R.detach()).abs().max() < 1e-5
    ((o2 * go.to(cuda)).sum() + (d2 * gd.to(cuda)).sum()).backward()
    assert (tr_g.grad.cpu() - tr_r.grad).abs().max() < 1e-4
    assert (rot_g.grad.cpu() - rot_r.grad).abs().max() < 2e-4 * max(1.0, rot_r.grad.abs().max().item())


def test_so3_orthogonality(cuda):
    # reference notebook barf/bug_hunting_with_Lauge.ipynb cell 30: max |R^T R - I| < 1e-4
    g = th.Generator().manual_seed(0)
    params = th.randn((1000, 3), generator=g) * th.rand((1000, 1), generator=g) * 3
    R = _ops().so3_to_SO3(params.to(cuda)).cpu()
    err = (th.matmul(R.permute(0, 2, 1), R) - th.eye(3).unsqueeze(0)).abs().max()
    assert err < 1e-4
    assert (R - ref_pose.so3_to_SO3(params)).abs().max() < 1e-5
