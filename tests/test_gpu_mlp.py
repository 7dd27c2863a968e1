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
    assert (out - ref).abs().max() < 5e-4
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
    assert (out - ref).abs().max() < 5e-4


def _bf16(x):
    return x.to(th.bfloat16).to(th.float32)


def _emulated_forward(net, pe_pos, pe_dir):
    """The oracle forward with operands rounded to bf16 where the kernel rounds them."""
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    sdq = {k: (_bf16(v) if k.endswith("weight") else v) for k, v in sd.items()}
    cfg = dict(n_hidden=net.n_hidden, n_segments=net.n_segments, delayed_direction=net.delayed_direction,
               delayed_density=net.delayed_density)
    return sd, sdq, cfg


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
    sd, sdq, cfg = _emulated_forward(net, None, None)
    P = ref_pe.barf_encoding(pos, 10, 1.0, identity, th.tensor(6.5))
    D = ref_pe.barf_encoding(d, 4, 1.0, identity, th.tensor(4.0))
    s_ref, c_ref = ref_mlp.nerf_model_forward(sd, cfg, P, D)
    err_s = ((sigma.cpu() - s_ref).abs() / (1 + s_ref.abs())).max().item()
    err_c = (rgb.cpu() - c_ref).abs().max().item()
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
