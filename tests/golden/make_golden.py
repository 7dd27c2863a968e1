"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from
/root/reference with the stub packages under oracle/_stubs) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_import as ri  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.detach().cpu().numpy() if isinstance(v, th.Tensor) else np.asarray(v))
                           for k, v in arrays.items()})
    print("wrote", name)


def main():
    ns = ri.load("barf")
    pe, arch, mi, cam = (ns.positional_encodings, ns.model_interpolation_architecture,
                         ns.model_interpolation, ns.model_camera_extrinsics)
    g = th.Generator().manual_seed(20240)

    # ---- positional encodings ---------------------------------------------------------------
    x = (th.rand((257, 3), generator=g) * 2 - 1) * 2.5
    d = th.nn.functional.normalize(th.randn((257, 3), generator=g), dim=1)
    t0 = th.rand((257, 1), generator=g) * 4 + 2
    t1 = t0 + th.rand((257, 1), generator=g) * 0.2 + 1e-3
    pw = th.full((257, 1), 1 / 555.5)
    out = {"x": x, "dir": d, "t0": t0, "t1": t1, "pw": pw}
    for alpha in (0.0, 2.5, 6.75, 10.0):
        enc = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, True, 1.0)
        enc.alpha = th.tensor(alpha)
        out[f"barf_id_a{alpha}"] = enc(x)
        out[f"mask_a{alpha}"] = enc.compute_mask(enc.alpha)
    enc = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, False, 2 * th.pi)
    enc.alpha = th.tensor(1.5)
    out["barf_noid_l4_2pi_a1.5"] = enc(d)
    out["fourier_l6"] = pe.FourierFeatures(6, 1.0)(x)
    for dv in (False, True):
        for pws in (0.0, 1.5):
            ie = pe.IntegratedFourierFeatures(10, 1.0, True, dv)
            ie.pixel_width_sigma = pws
            out[f"ipe_dv{int(dv)}_pws{pws}"] = ie(x, d, pw, t0, t1)
    ib = pe.IntegratedBarfFourierFeatures(10, 0.0, 1.0, 2.0, True, 1.0, True)
    ib.alpha = th.tensor(4.3)
    ib.pixel_width_sigma = 0.0
    out["ipe_barf_a4.3"] = ib(x, d, pw, t0, t1)
    # alpha schedule (update_alpha)
    enc = pe.BarfPositionalEncoding(10, 0.0, 0.5, 2.5, True, 1.0)
    epochs = np.array([0.0, 0.5, 1.25, 2.0, 2.5, 3.0])
    alphas = []
    for e in epochs:
        enc.update_alpha(float(e))
        alphas.append(float(enc.alpha))
    out["alpha_epochs"], out["alpha_values"] = epochs, np.array(alphas)
    save("positional_encodings", **out)

    # ---- NerfModel (small widths keep the fixture small) ---------------------------------------
    for tag, kw in (("std_small", dict(n_hidden=2, hidden_dim=64, delayed_direction=True, delayed_density=False, n_segments=2)),
                    ("nodelay", dict(n_hidden=1, hidden_dim=64, delayed_direction=False, delayed_density=True, n_segments=2)),
                    ("flat", dict(n_hidden=0, hidden_dim=64, delayed_direction=True, delayed_density=False, n_segments=1))):
        th.manual_seed(7)
        ep = pe.BarfPositionalEncoding(10, 0.0, 1.0, 2.0, True, 1.0)
        ed = pe.BarfPositionalEncoding(4, 0.0, 1.0, 2.0, True, 1.0)
        ep.alpha, ed.alpha = th.tensor(6.5), th.tensor(4.0)
        net = arch.NerfModel(position_encoder=ep, direction_encoder=ed, **kw)
        pos = x[:96].clone().requires_grad_()
        dd = d[:96].clone().requires_grad_()
        dens, rgb = net(pos, dd, pw[:96], t0[:96], t1[:96])
        gs = th.randn(dens.shape, generator=g) * 0.1
        gc = th.randn(rgb.shape, generator=g)
        ((dens * gs).sum() + (rgb * gc).sum()).backward()
        arrays = {"pos": pos, "dir": dd, "density": dens, "rgb": rgb, "g_density": gs, "g_rgb": gc,
                  "d_pos": pos.grad, "d_dir": dd.grad}
        for k, v in net.state_dict().items():
            arrays["sd." + k] = v
        for k, p in net.named_parameters():
            arrays["grad." + k] = p.grad
        save(f"nerf_model_{tag}", **arrays)

    # ---- sampling, compositing, resampling ----------------------------------------------------
    th.manual_seed(3)
    ep = pe.BarfPositionalEncoding(4, 4.0, 1.0, 2.0, False, 1.0)
    ed = pe.BarfPositionalEncoding(2, 2.0, 1.0, 2.0, False, 1.0)
    tiny = arch.NerfModel(1, 64, True, False, 1, ep, ed)
    tiny2 = arch.NerfModel(1, 64, True, False, 1, ep, ed)
    m = mi.NerfInterpolation(2.0, 8.0, tiny, 48, "stratified_uniform", -1.0, "middle", tiny2, 16)
    arrays = {}
    for name, (B, S, strat, off) in {"equi": (5, 64, "equidistant", -1.0), "strat": (9, 33, "stratified_uniform", -1.0),
                                     "strat0": (4, 128, "stratified_uniform", 0.0)}.items():
        th.manual_seed(11)
        ts, te = m._sample_t_stratified_uniform(B, S, strat, off)
        th.manual_seed(11)
        jit = th.rand((B, S)) if strat == "stratified_uniform" else th.zeros(0)
        u = th.rand((B, 1)) if off != 0 else th.zeros(0)
        arrays.update({f"{name}_t_start": ts, f"{name}_t_end": te, f"{name}_jitter": jit, f"{name}_offset": u})
    sigma = th.nn.functional.softplus(th.randn((33, 64), generator=g) * 2).requires_grad_()
    col = th.rand((33, 64, 3), generator=g).requires_grad_()
    delta = th.rand((33, 64), generator=g) * 0.1 + 0.01
    rgb, w = m._render_rays(sigma, col, delta)
    g_rgb, g_w = th.randn((33, 3), generator=g), th.randn((33, 64), generator=g) * 0.1
    ((rgb * g_rgb).sum() + (w * g_w).sum()).backward()
    arrays.update({"r_sigma": sigma, "r_color": col, "r_delta": delta, "r_rgb": rgb, "r_w": w, "r_g_rgb": g_rgb,
                   "r_g_w": g_w, "r_d_sigma": sigma.grad, "r_d_color": col.grad})
    for name, (B, Sc, Sf) in {"a": (24, 64, 256), "b": (7, 4, 16), "c": (5, 100, 300)}.items():
        ts, te = m._sample_t_stratified_uniform(B, Sc, "equidistant", -1.0)
        wts = th.rand((B, Sc), generator=g) ** 4 * (th.rand((B, Sc), generator=g) > 0.3)
        wts[:, 0] += 1e-3
        f0, f1 = m._sample_t_pdf_weighted(ts, wts, te - ts, Sf)
        arrays.update({f"p{name}_t": ts, f"p{name}_w": wts, f"p{name}_delta": te - ts, f"p{name}_t_start": f0,
                       f"p{name}_t_end": f1})
    save("sampling_render", **arrays)

    # ---- the whole render module (proposal + radiance), explicit uniforms ------------------------
    B = 12
    o = th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0
    dr = th.nn.functional.normalize(-o + 0.3 * th.randn((B, 3), generator=g), dim=1)
    th.manual_seed(5)
    fine, coarse = m(o, dr, th.full((B, 1), 1 / 555.0))
    th.manual_seed(5)
    jit, off = th.rand((B, 16)), th.rand((B, 1))
    arrays = {"o": o, "d": dr, "rgb_fine": fine, "rgb_coarse": coarse, "jitter": jit, "offset": off}
    for k, v in tiny.state_dict().items():
        arrays["rad." + k] = v
    for k, v in tiny2.state_dict().items():
        arrays["prop." + k] = v
    save("render_module", **arrays)

    # ---- camera extrinsics -------------------------------------------------------------------------
    ce = cam.CameraExtrinsics(9, 1e-3, 1e-5, 100)
    with th.no_grad():
        ce.rotation.copy_(th.randn((9, 3), generator=g) * 0.4)
        ce.rotation[0] = 0.0
        ce.translation.copy_(th.randn((9, 3), generator=g) * 0.2)
    idx = th.randint(0, 9, (40,), generator=g)
    o = th.randn((40, 3), generator=g)
    dr = th.nn.functional.normalize(th.randn((40, 3), generator=g), dim=1)
    no, nd, R, t = ce(idx, o, dr)
    go, gd = th.randn((40, 3), generator=g), th.randn((40, 3), generator=g)
    ((no * go).sum() + (nd * gd).sum()).backward()
    big = th.randn((200, 3), generator=g) * th.rand((200, 1), generator=g) * 3
    save("camera_extrinsics", rotation=ce.rotation, translation=ce.translation, idx=idx, o=o, d=dr, new_o=no, new_d=nd,
         R=R, t=t, g_o=go, g_d=gd, d_rotation=ce.rotation.grad, d_translation=ce.translation.grad,
         so3=big, SO3=cam.CameraExtrinsics.so3_to_SO3(big))

    # ---- activations of the GARF / SARF / Gabor variants ---------------------------------------------
    gz = th.randn((64, 16), generator=g) * 2
    inv = th.rand(16, generator=g) * 1.5 + 0.5
    arrays = {"x": gz, "inv_std": inv, "gauss": ns.gaussian.GaussActivation.apply(gz, inv ** 2 + 1e-6)}
    sarf = ri.load("sarf").activation
    act = sarf.SarfAct(16, 0.5, 2.0) if hasattr(sarf, "SarfAct") else None
    if act is not None:
        arrays["sarf_param"] = list(act.parameters())[0]
        arrays["sarf"] = act(gz)
    save("activations", **arrays)


if __name__ == "__main__":
    main()
