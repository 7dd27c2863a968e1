"""GPU parity of the positional encodings and the fused MLP forward against the CPU oracle."""
import pytest
import torch as th

from oracle import ref_mlp, ref_pe

pytestmark = pytest.mark.gpu


def _mods():
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    return arch, pe


@pytest.mark.parametrize("levels,identity,alpha,scale", [(10, True, 2.5, 1.0), (4, True, 0.0, 1.0),
                                                        (10, False, 10.0, 1.0), (6, True, 3.0, 6.2831853)])
def test_barf_encoding(cuda, levels, identity, alpha, scale):
    arch, pe = _mods()
    g = th.Generator().manual_seed(levels)
    x = (th.rand((777, 3), generator=g) * 2 - 1) * 2.0
    enc = pe.BarfPositionalEncoding(levels, 0.0, 1.0, 2.0, identity, scale).to(cuda)
    enc.alpha.fill_(alpha)
    out = enc(x.to(cuda)).cpu()
    ref = ref_pe.barf_encoding(x, levels, scale, identity, th.tensor(alpha))
    # double-angle recurrence: ~2^L * 1e-7 absolute
    assert out.shape == ref.shape
    assert (out - ref).abs().max() < 2e-5
    m = enc.compute_mask(enc.alpha).cpu()
    assert (m - ref_pe.barf_mask(th.tensor(alpha), levels)).abs().max() < 1e-6


def test_barf_mask_known_answer(cuda):
    # alpha = 2.5 -> [1, 1, 0.5, 0, ...]
    arch, pe = _mods()
    enc = pe.BarfPositionalEncoding(6, 0.0, 1.0, 2.0, False, 1.0).to(cuda)
    enc.alpha.fill_(2.5)
    x = th.zeros((1, 3), device=cuda)   # cos(0) = 1: the cos block shows the mask
    out = enc(x).cpu()[0]
    assert th.allclose(out[:6], th.tensor([1.0, 1.0, 0.5, 0.0, 0.0, 0.0]), atol=1e-6)


@pytest.mark.parametrize("distribute,pws,masked", [(False, 0.0, False), (True, 1.5, False), (True, 0.0, True)])
def test_integrated_encoding(cuda, distribute, pws, masked):
    arch, pe = _mods()
    g = th.Generator().manual_seed(5)
    n = 500
    pos = (th.rand((n, 3), generator=g) * 2 - 1) * 3
    d = th.nn.functional.normalize(th.randn((n, 3), generator=g), dim=1)
    t0 = th.rand((n, 1), generator=g) * 4 + 2
    t1 = t0 + th.rand((n, 1), generator=g) * 0.2 + 1e-3
    pw = th.full((n, 1), 1 / 555.0)
    if masked:
        enc = pe.IntegratedBarfFourierFeatures(10, 0.0, 1.0, 2.0, True, 1.0, distribute).to(cuda)
        enc.alpha.fill_(4.3)
        alpha = th.tensor(4.3)
    else:
        enc = pe.IntegratedFourierFeatures(10, 1.0, True, distribute).to(cuda)
        alpha = None
    enc.pixel_width_sigma = pws
    out = enc(pos.to(cuda), d.to(cuda), pw.to(cuda), t0.to(cuda), t1.to(cuda)).cpu()
    ref = ref_pe.integrated_encoding(pos, d, pw, t0, t1, 10, 1.0, True, distribute, pws, alpha)
    assert (out - ref).abs().max() < 2e-5


def _run_case(cuda, n, n_hidden, hidden, delayed_dir, delayed_dens, n_segments, identity=True, seed=0):
    arch, pe = _mods()
    th.manual_seed(seed)
    enc_p = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, identity, 1.0)
    enc_d = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, identity, 1.0)
    net = arch.NerfModel(n_hidden, hidden, delayed_dir, delayed_dens, n_segments, enc_p, enc_d).to(cuda)
    enc_p.alpha.fill_(6.5); enc_d.alpha.fill_(4.0)
    g = th.Generator().manual_seed(seed + 1)
    pos = (th.rand((n, 3), generator=g) * 2 - 1) * 1.5
    d = th.nn.functional.normalize(th.randn((n, 3), generator=g), dim=1)
    with th.no_grad():
        sigma, rgb = net(pos.to(cuda), d.to(cuda))
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    cfg = dict(n_hidden=net.n_hidden, n_segments=net.n_segments, delayed_direction=net.delayed_direction,
               delayed_density=net.delayed_density)
    P = ref_pe.barf_encoding(pos, 10, 1.0, identity, th.tensor(6.5))
    D = ref_pe.barf_encoding(d, 4, 1.0, identity, th.tensor(4.0))
    s_ref, c_ref = ref_mlp.nerf_model_forward(sd, cfg, P, D)                      # the fp32 reference
    s_emu, c_emu = ref_mlp.nerf_model_forward(sd, cfg, P, D, emulate_bf16=True)   # same, bf16 operands
    err_s = ((sigma.cpu() - s_ref).abs() / (1 + s_ref.abs())).max().item()
    err_c = (rgb.cpu() - c_ref).abs().max().item()
    # against the bf16-operand restatement the kernel must agree far more tightly
    assert ((sigma.cpu() - s_emu).abs() / (1 + s_emu.abs())).max().item() < 3e-3
    assert (rgb.cpu() - c_emu).abs().max().item() < 3e-3
    return err_s, err_c, sigma.cpu(), rgb.cpu()


def test_fused_forward_standard(cuda):
    # north_star: bf16-MLP rgb within 1e-2 abs of the fp32 reference
    err_s, err_c, sigma, rgb = _run_case(cuda, 128 * 9 + 37, 4, 256, True, False, 2)
    assert th.isfinite(sigma).all() and th.isfinite(rgb).all()
    assert err_c < 1e-2, err_c
    assert err_s < 3e-2, err_s
    assert (sigma >= 0).all() and (rgb >= 0).all() and (rgb <= 1).all()


@pytest.mark.parametrize("n_hidden,hidden,ddir,ddens,nseg,identity", [
    (0, 64, True, False, 1, True), (1, 128, True, False, 1, False), (2, 128, False, False, 2, True),
    (2, 256, True, True, 2, True), (3, 192, False, True, 3, False), (4, 256, True, False, 2, False),
    (1, 100, True, False, 2, True)])
def test_fused_forward_variants(cuda, n_hidden, hidden, ddir, ddens, nseg, identity):
    # reference notebook cells 39-42: every architecture gives finite, in-range outputs
    err_s, err_c, sigma, rgb = _run_case(cuda, 300, n_hidden, hidden, ddir, ddens, nseg, identity, seed=3)
    assert th.isfinite(sigma).all() and th.isfinite(rgb).all()
    assert (sigma >= 0).all() and (rgb >= 0).all() and (rgb <= 1).all()
    assert err_c < 1e-2, err_c
    assert err_s < 3e-2, err_s


def _rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _grad_case(cuda, n, n_hidden, hidden, ddir, ddens, nseg, identity, input_grads, seed=0):
    """relative gradient errors of the CUDA path against (a) the fp32 reference arithmetic and
    (b) the same arithmetic with bf16 operands (what the kernels are built to compute)."""
    arch, pe = _mods()
    th.manual_seed(seed)
    enc_p = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, identity, 1.0)
    enc_d = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, identity, 1.0)
    net = arch.NerfModel(n_hidden, hidden, ddir, ddens, nseg, enc_p, enc_d).to(cuda)
    enc_p.alpha.fill_(6.5); enc_d.alpha.fill_(4.0)
    g = th.Generator().manual_seed(seed + 1)
    pos = (th.rand((n, 3), generator=g) * 2 - 1) * 1.5
    d = th.nn.functional.normalize(th.randn((n, 3), generator=g), dim=1)
    gs = th.randn((n,), generator=g) * 0.1
    gc = th.randn((n, 3), generator=g)

    pos_g = pos.to(cuda).requires_grad_(input_grads)
    d_g = d.to(cuda).requires_grad_(input_grads)
    sigma, rgb = net(pos_g, d_g)
    ((sigma * gs.to(cuda)).sum() + (rgb * gc.to(cuda)).sum()).backward()

    cfg = dict(n_hidden=n_hidden, n_segments=nseg, delayed_direction=ddir, delayed_density=ddens)
    out = {}
    for emulate in (False, True):
        sd = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and v.dim() > 0)
              for k, v in net.state_dict().items()}
        pos_r = pos.clone().requires_grad_(input_grads)
        d_r = d.clone().requires_grad_(input_grads)
        P = ref_pe.barf_encoding(pos_r, 10, 1.0, identity, th.tensor(6.5))
        D = ref_pe.barf_encoding(d_r, 4, 1.0, identity, th.tensor(4.0))
        s_ref, c_ref = ref_mlp.nerf_model_forward(sd, cfg, P, D, emulate_bf16=emulate)
        ((s_ref * gs).sum() + (c_ref * gc).sum()).backward()
        errs = {}
        for name, p in net.named_parameters():
            assert p.grad is not None, name
            assert th.isfinite(p.grad).all(), name
            errs[name] = _rel(p.grad.cpu(), sd[name].grad)
        if input_grads:
            errs["pos"] = _rel(pos_g.grad.cpu(), pos_r.grad)
            errs["dir"] = _rel(d_g.grad.cpu(), d_r.grad)
        out[emulate] = errs
    return out


def _check_grads(out):
    # bf16 operands put the gradients of a 12-layer ReLU MLP ~10 % (relative L2, random upstream
    # gradients) away from fp32 — scripts/emulate_bf16.py reproduces that on the CPU — so the
    # kernels are held to the bf16-operand restatement tightly and to fp32 loosely.
    bad = {k: v for k, v in out[True].items() if v > 2e-2}
    assert not bad, ("vs bf16-operand oracle", bad)
    bad = {k: v for k, v in out[False].items() if v > 0.25}
    assert not bad, ("vs fp32 oracle", bad)


def test_fused_backward_standard(cuda):
    _check_grads(_grad_case(cuda, 128 * 5 + 17, 4, 256, True, False, 2, True, True))


@pytest.mark.parametrize("n_hidden,hidden,ddir,ddens,nseg,identity,inp", [
    (1, 128, True, False, 1, False, False), (2, 128, False, False, 2, True, True),
    (2, 256, True, True, 2, True, True), (3, 192, False, True, 3, False, False),
    (0, 64, True, False, 1, True, True), (1, 100, True, False, 2, True, True)])
def test_fused_backward_variants(cuda, n_hidden, hidden, ddir, ddens, nseg, identity, inp):
    _check_grads(_grad_case(cuda, 300, n_hidden, hidden, ddir, ddens, nseg, identity, inp, seed=2))


@pytest.mark.parametrize("inp", [False, True])
def test_fused_backward_many_tiles_per_cta(cuda, inp):
    """More tiles than SMs (every CTA walks several tiles): the slab hand-off protocol across
    tile boundaries, with and without the encoding-gradient ops (without them the last op of a
    tile publishes slabs no MMA reads — the case that once ran two barrier phases ahead)."""
    _check_grads(_grad_case(cuda, 128 * 400 + 5, 2, 128, True, False, 2, True, inp, seed=4))
