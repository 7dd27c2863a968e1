"""GARF / SARF / Gabor on the CPU: the oracle restatement against the fixtures generated from
the unmodified reference (tests/golden/garf.npz), and the module surface of the host mirror
(state-dict keys, seeded initialisation, parameter groups, optimiser schedule)."""
import os

import numpy as np
import pytest
import torch as th

from oracle import ref_garf

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "garf.npz")
THIN = 37


def _g():
    z = np.load(G)
    return {k: th.from_numpy(z[k]) for k in z.files}


def _thin(t):
    return t if t.numel() <= 8192 else t.flatten()[::THIN]


def _sums(t):
    d = t.detach().double()
    return th.stack((d.sum(), d.abs().sum(), (d * d).sum()))


def test_activation_oracle_matches_reference_outputs_and_gradients():
    g = _g()
    up = g["act_up"]
    for name, fn, params in (("gauss", ref_garf.gauss_act, ("gauss_p",)),
                             ("sarf", ref_garf.sarf_act, ("sarf_p",)),
                             ("gabor", ref_garf.gabor_act, ("gabor_p0", "gabor_p1"))):
        x = g["act_x"].clone().requires_grad_()
        ps = [g[k].clone().requires_grad_() for k in params]
        y = fn(x, *ps)
        assert th.allclose(y, g[name + "_y"], rtol=1e-6, atol=1e-7), name
        grads = th.autograd.grad(y, [x] + ps, up)
        assert th.allclose(grads[0], g[name + "_dx"], rtol=1e-5, atol=1e-6), name
        keys = [name + "_dp"] if len(ps) == 1 else [name + "_dp0", name + "_dp1"]
        for gr, k in zip(grads[1:], keys):
            assert th.allclose(gr, g[k], rtol=1e-4, atol=1e-5), k


def _seeded_nets():
    from nerf_experiments_b200.model_garf_proposal import ProposalNetwork
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    th.manual_seed(77)
    prop = ProposalNetwork(0.5, 1.5)      # creation order of GarfModel.__init__
    rad = RadianceNetwork(0.5, 1.5)
    return prop, rad


def test_garf_modules_share_keys_and_seeded_init_with_the_reference():
    g = _g()
    prop, rad = _seeded_nets()
    for tag, net in (("rad", rad), ("prop", prop)):
        ref_keys = sorted(k[len(tag) + 7:] for k in g if k.startswith(tag + ".sdsum."))
        assert sorted(net.state_dict().keys()) == ref_keys
        for k, v in net.state_dict().items():
            assert th.allclose(_sums(v), g[f"{tag}.sdsum.{k}"], rtol=1e-12, atol=0), k
    assert len(list(rad.parameters_linear())) == 20 and len(list(rad.parameters_gaussian())) == 8
    assert sum(p.numel() for p in rad.parameters()) == 601604
    assert sum(p.numel() for p in prop.parameters()) == 167297


def test_garf_network_oracle_matches_reference_outputs_and_gradients():
    g = _g()
    prop, rad = _seeded_nets()
    sd = {k: v.detach().clone().requires_grad_() for k, v in rad.state_dict().items()}
    rgb, dens = ref_garf.radiance_network(sd, g["net_pos"], g["net_dir"])
    assert th.allclose(rgb, g["rad_rgb"], rtol=1e-5, atol=1e-6)
    assert th.allclose(dens, g["rad_density"], rtol=1e-5, atol=1e-6)
    ((rgb * g["up_rgb"]).sum() + (dens * g["up_density"]).sum()).backward()
    for k, v in sd.items():
        ref = g["rad.grad." + k]
        assert th.allclose(_thin(v.grad), ref, rtol=2e-3, atol=1e-5 * float(ref.abs().max() + 1)), k
    sdp = {k: v.detach().clone().requires_grad_() for k, v in prop.state_dict().items()}
    sp = ref_garf.proposal_network(sdp, g["net_pos"])
    assert th.allclose(sp, g["prop_sigma"], rtol=1e-5, atol=1e-6)
    (sp * g["up_prop"]).sum().backward()
    for k, v in sdp.items():
        ref = g["prop.grad." + k]
        assert th.allclose(_thin(v.grad), ref, rtol=2e-3, atol=1e-5 * float(ref.abs().max() + 1)), k


def test_garf_model_surface_and_schedules():
    from nerf_experiments_b200.model_garf import GarfModel
    m = GarfModel(2.0, 7.0, 64, 192, 0.5, 1.5, 0.25, 1e-3, 1e-4, 100, 0.0, 5e-4, 5e-5, 200, 1e-6)
    assert m.automatic_optimization is False
    (po, ro), (ps, rs) = m.configure_optimizers()
    assert [g["lr"] for g in po.param_groups] == pytest.approx([1e-3, 0.25e-3])
    assert [g["weight_decay"] for g in ro.param_groups] == [1e-6, 1e-6]
    # ExponentialLR with last_epoch = -decay_end - 1 (reference garf/model_garf.py:384-392): gamma^k
    assert ps.gamma == pytest.approx(2 ** (np.log2(1e-4 / 1e-3) / 100))
    assert rs.gamma == pytest.approx(m._calculate_decay_factor(5e-4, 5e-5, 200))
    # no CPU fallback: the activations refuse CPU tensors
    with pytest.raises(RuntimeError):
        m.radiance_network(th.zeros(4, 3), th.zeros(4, 3))


def test_outer_loss_known_answers():
    """A proposal histogram that bounds the radiance histogram costs nothing; one that misses
    mass is penalised by (excess)^2 / w (Mip-NeRF 360 eq. 13 as nerfacc implements it)."""
    pdf_outer_loss = ref_garf.pdf_outer_loss      # the CUDA kernel is held to this oracle in tests/test_gpu_garf.py
    t = th.tensor([[0.0, 1.0, 2.0, 3.0]])
    cdf = th.tensor([[0.0, 0.2, 0.7, 1.0]])
    assert float(pdf_outer_loss(t, cdf, t, cdf).sum()) == 0.0
    flat = th.tensor([[0.0, 1 / 3, 2 / 3, 1.0]])
    loss = pdf_outer_loss(t, cdf, t, flat)
    assert loss[0, 1] == pytest.approx((0.5 - 1 / 3) ** 2 / (0.5 + 1e-7), rel=1e-5)
    assert float(loss[0, 0]) == 0.0 and float(loss[0, 2]) == 0.0
