"""The oracle against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch as th

from oracle import ref_mlp, ref_pe, ref_pose, ref_render, ref_resample, ref_sampling, ref_step

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    return {k: th.from_numpy(z[k]) for k in z.files}


def test_positional_encodings_bit_exact():
    g = load("positional_encodings")
    x, d, t0, t1, pw = g["x"], g["dir"], g["t0"], g["t1"], g["pw"]
    for alpha in (0.0, 2.5, 6.75, 10.0):
        assert th.equal(ref_pe.barf_encoding(x, 10, 1.0, True, th.tensor(alpha)), g[f"barf_id_a{alpha}"])
        assert th.equal(ref_pe.barf_mask(th.tensor(alpha), 10), g[f"mask_a{alpha}"])
    assert th.equal(ref_pe.barf_encoding(d, 4, 2 * th.pi, False, th.tensor(1.5)), g["barf_noid_l4_2pi_a1.5"])
    assert th.equal(ref_pe.barf_encoding(x, 6, 1.0, False, None), g["fourier_l6"])
    for dv in (False, True):
        for pws in (0.0, 1.5):
            out = ref_pe.integrated_encoding(x, d, pw, t0, t1, 10, 1.0, True, dv, pws)
            assert th.equal(out, g[f"ipe_dv{int(dv)}_pws{pws}"])
    out = ref_pe.integrated_encoding(x, d, pw, t0, t1, 10, 1.0, True, True, 0.0, th.tensor(4.3))
    assert th.equal(out, g["ipe_barf_a4.3"])
    for e, a in zip(g["alpha_epochs"].tolist(), g["alpha_values"].tolist()):
        assert ref_pe.barf_alpha(e, 10, 0.0, 0.5, 2.5) == pytest.approx(a, rel=1e-6)


def test_barf_mask_known_answer():
    # alpha = 2.5 -> [1, 1, 0.5, 0, ...]
    m = ref_pe.barf_mask(th.tensor(2.5), 6, 1)[0]
    assert th.allclose(m, th.tensor([1.0, 1.0, 0.5, 0.0, 0.0, 0.0]), atol=1e-7)


@pytest.mark.parametrize("tag,cfg", [
    ("std_small", dict(n_hidden=2, n_segments=2, delayed_direction=True, delayed_density=False)),
    ("nodelay", dict(n_hidden=1, n_segments=2, delayed_direction=False, delayed_density=True)),
    ("flat", dict(n_hidden=0, n_segments=1, delayed_direction=True, delayed_density=False))])
def test_nerf_model_forward_backward(tag, cfg):
    g = load(f"nerf_model_{tag}")
    sd = {k[3:]: v.clone().requires_grad_(v.dim() > 0) for k, v in g.items() if k.startswith("sd.")}
    pos, d = g["pos"].clone().requires_grad_(), g["dir"].clone().requires_grad_()
    P = ref_pe.barf_encoding(pos, 10, 1.0, True, th.tensor(6.5))
    D = ref_pe.barf_encoding(d, 4, 1.0, True, th.tensor(4.0))
    dens, rgb = ref_mlp.nerf_model_forward(sd, cfg, P, D)
    assert th.equal(dens, g["density"]) and th.equal(rgb, g["rgb"])
    ((dens * g["g_density"]).sum() + (rgb * g["g_rgb"]).sum()).backward()
    assert th.allclose(pos.grad, g["d_pos"], rtol=1e-5, atol=1e-7)
    assert th.allclose(d.grad, g["d_dir"], rtol=1e-5, atol=1e-7)
    for k, v in g.items():
        if k.startswith("grad."):
            assert th.allclose(sd[k[5:]].grad, v, rtol=1e-5, atol=1e-7), k


def test_sampling_bit_exact():
    g = load("sampling_render")
    for name, (B, S, off) in {"equi": (5, 64, -1.0), "strat": (9, 33, -1.0), "strat0": (4, 128, 0.0)}.items():
        jit = g[f"{name}_jitter"] if g[f"{name}_jitter"].numel() else None
        u = g[f"{name}_offset"] if g[f"{name}_offset"].numel() else None
        ts, te = ref_sampling.sample_uniform(2.0, 8.0, B, S, jit, u, off)
        assert th.equal(ts, g[f"{name}_t_start"]) and th.equal(te, g[f"{name}_t_end"])


def test_render_rays_and_gradients():
    g = load("sampling_render")
    sigma, col = g["r_sigma"].clone().requires_grad_(), g["r_color"].clone().requires_grad_()
    rgb, w = ref_render.render_rays(sigma, col, g["r_delta"])
    assert th.equal(rgb, g["r_rgb"]) and th.equal(w, g["r_w"])
    ((rgb * g["r_g_rgb"]).sum() + (w * g["r_g_w"]).sum()).backward()
    assert th.allclose(sigma.grad, g["r_d_sigma"], rtol=1e-6, atol=1e-8)
    assert th.allclose(col.grad, g["r_d_color"], rtol=1e-6, atol=1e-8)


def test_compositing_is_textbook():
    # SURVEY §6 probe: the reference's formula equals the textbook one (up to the 3*(1/3) ulp)
    g = load("sampling_render")
    sigma, col, delta = g["r_sigma"], g["r_color"], g["r_delta"]
    alpha = 1 - th.exp(-sigma * delta)
    T = th.cumprod(th.cat((th.ones(33, 1), 1 - alpha[:, :-1]), 1), 1)
    assert (th.sum((T * alpha).unsqueeze(-1) * col, 1) - g["r_rgb"]).abs().max() < 1e-5


@pytest.mark.parametrize("name,Sf", [("a", 256), ("b", 16), ("c", 300)])
def test_pdf_resampling_bit_exact(name, Sf):
    g = load("sampling_render")
    t0, t1, counts, failed = ref_resample.sample_pdf_weighted(g[f"p{name}_t"].numpy(), g[f"p{name}_w"].numpy(),
                                                              g[f"p{name}_delta"].numpy(), Sf, 2.0, 8.0)
    assert not failed
    assert np.array_equal(t0, g[f"p{name}_t_start"].numpy())
    assert np.array_equal(t1, g[f"p{name}_t_end"].numpy())
    assert (counts.sum(axis=1) == Sf).all() and (counts >= 1).all()


def test_pdf_resampling_known_answer():
    # SURVEY §8c probe: w = [.1,.6,.2,.05], 4 -> 16 gives counts [2, 9, 3, 2]
    n, ok = ref_resample.counts_for_ray(np.array([0.1, 0.6, 0.2, 0.05], dtype=np.float32), 16)
    assert ok and n.tolist() == [2, 9, 3, 2]


def test_render_module_end_to_end():
    g = load("render_module")
    sd_r = {k[4:]: v for k, v in g.items() if k.startswith("rad.")}
    sd_p = {k[5:]: v for k, v in g.items() if k.startswith("prop.")}
    cfg = dict(n_hidden=1, n_segments=1, delayed_direction=True, delayed_density=False)
    pe_cfg = dict(pos_levels=4, dir_levels=2, scale=1.0, identity=False, alpha_pos=th.tensor(4.0), alpha_dir=th.tensor(2.0))
    fine, coarse, _ = ref_step.render(sd_r, cfg, pe_cfg, g["o"], g["d"], 2.0, 8.0, 48, "middle",
                                      {"jitter": g["jitter"], "offset": g["offset"]}, sd_p, 16,
                                      "stratified_uniform", -1.0)
    assert th.allclose(fine, g["rgb_fine"], rtol=0, atol=1e-6)
    assert th.allclose(coarse, g["rgb_coarse"], rtol=0, atol=1e-6)


def test_camera_extrinsics():
    g = load("camera_extrinsics")
    rot, tr = g["rotation"].clone().requires_grad_(), g["translation"].clone().requires_grad_()
    no, nd, R, t = ref_pose.pose_forward(rot, tr, g["idx"], g["o"], g["d"])
    assert th.equal(no, g["new_o"]) and th.equal(nd, g["new_d"]) and th.equal(R, g["R"]) and th.equal(t, g["t"])
    ((no * g["g_o"]).sum() + (nd * g["g_d"]).sum()).backward()
    assert th.allclose(rot.grad, g["d_rotation"], rtol=1e-5, atol=1e-7)
    assert th.allclose(tr.grad, g["d_translation"], rtol=1e-5, atol=1e-7)
    assert th.equal(ref_pose.so3_to_SO3(g["so3"]), g["SO3"])
    # notebook cell 30: orthogonality of so3_to_SO3 below 1e-4
    Rb = g["SO3"]
    assert (th.matmul(Rb.permute(0, 2, 1), Rb) - th.eye(3)).abs().max() < 1e-4


def test_activations():
    g = load("activations")
    assert th.allclose(ref_mlp.gauss_act(g["x"], g["inv_std"]), g["gauss"], rtol=1e-6, atol=1e-8)
    if "sarf" in g:
        assert th.allclose(ref_mlp.sarf_act(g["x"], g["sarf_param"]), g["sarf"], rtol=1e-5, atol=1e-6)
