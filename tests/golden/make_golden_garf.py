"""Generates tests/golden/garf.npz from the UNMODIFIED reference modules (this container only):
Gaussian / SARF / Gabor activations with gradients, and outputs + parameter gradients of the
GARF radiance and proposal networks at a fixed seed.  Run: python tests/golden/make_golden_garf.py"""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_import as ri  # noqa: E402


THIN = 37   # large gradient tensors are stored as every THIN-th element of the flattened tensor


def thin(t):
    return t if t.numel() <= 8192 else t.flatten()[::THIN]


def sums(t):
    """(sum, sum of |.|, sum of squares) in float64: pins a seeded parameter tensor in 24 bytes."""
    d = t.detach().double()
    return th.stack((d.sum(), d.abs().sum(), (d * d).sum()))


def main():
    th.manual_seed(20240)
    g = th.Generator().manual_seed(5)
    out = {}
    # ---- activations (forward + gradients w.r.t. x and the parameters) --------------------------
    x = (th.randn((96, 24), generator=g) * 1.5).requires_grad_()
    x.data[0, :4] = 0.0           # signbit / abs edge case of SARF
    up = th.randn((96, 24), generator=g)
    garf = ri.load("garf")
    inv = (th.rand(24, generator=g) * 1.5 + 0.25).requires_grad_()
    y = garf.gaussian.GaussActivation.apply(x, inv ** 2 + 1e-6)
    gx, gp = th.autograd.grad(y, (x, inv), up)
    out.update(act_x=x, act_up=up, gauss_p=inv, gauss_y=y, gauss_dx=gx, gauss_dp=gp)
    sarf = ri.load("sarf").activation
    act = sarf.SarfAct(24, 0.5, 2.0)
    y = act(x)
    gx, gp = th.autograd.grad(y, (x, act.frequency), up)
    out.update(sarf_p=act.frequency, sarf_y=y, sarf_dx=gx, sarf_dp=gp)
    gab = ri.load("gaborf").gabor
    act = gab.GaborAct(24, 0.25, 1.75)
    y = act(x)
    gx, gp0, gp1 = th.autograd.grad(y, (x, act.inv_standard_deviation, act.spread), up)
    out.update(gabor_p0=act.inv_standard_deviation, gabor_p1=act.spread, gabor_y=y, gabor_dx=gx,
               gabor_dp0=gp0, gabor_dp1=gp1)
    # ---- GARF networks ----------------------------------------------------------------------
    garf = ri.load("garf")
    th.manual_seed(77)
    prop = garf.model_proposal.ProposalNetwork(0.5, 1.5)      # creation order of GarfModel.__init__
    rad = garf.model_radiance.RadianceNetwork(0.5, 1.5)
    pos = th.randn((160, 3), generator=g) * 1.2
    dirs = th.nn.functional.normalize(th.randn((160, 3), generator=g), dim=1)
    rgb, dens = rad(pos, dirs)
    up_rgb, up_d = th.randn(rgb.shape, generator=g), th.randn(dens.shape, generator=g)
    names = [n for n, _ in rad.named_parameters()]
    grads = th.autograd.grad((rgb * up_rgb).sum() + (dens * up_d).sum(), list(rad.parameters()))
    out.update(net_pos=pos, net_dir=dirs, rad_rgb=rgb, rad_density=dens, up_rgb=up_rgb, up_density=up_d)
    for k, v in rad.state_dict().items():
        out["rad.sdsum." + k] = sums(v)
    for n, gr in zip(names, grads):
        out["rad.grad." + n] = thin(gr)
    sp = prop(pos)
    up_p = th.randn(sp.shape, generator=g)
    gradsp = th.autograd.grad((sp * up_p).sum(), list(prop.parameters()))
    out.update(prop_sigma=sp, up_prop=up_p)
    for k, v in prop.state_dict().items():
        out["prop.sdsum." + k] = sums(v)
    for (n, _), gr in zip(prop.named_parameters(), gradsp):
        out["prop.grad." + n] = thin(gr)
    arrays = {k: (v.detach().numpy() if isinstance(v, th.Tensor) else np.asarray(v)) for k, v in out.items()}
    # the fixture stays small: weights as float32, compressed
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "garf.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(arrays), "arrays")


if __name__ == "__main__":
    main()
