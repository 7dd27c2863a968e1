"""The GARF tile programs (host-side compiler nerf_experiments_b200/garf_program.py), executed by the
CPU interpreter tests/garf_sim.py with the kernels' roundings, against the oracle restatement of the
reference networks (oracle/ref_garf.py, pinned by tests/golden/garf.npz) — outputs, parameter gradients
and input gradients. No GPU: this pins op / step order, packing descriptors, stash layouts and the
weight-gradient units; tests/test_gpu_garf.py then holds the CUDA kernels to the same oracle."""
import os
import sys

import pytest
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import garf_sim  # noqa: E402
from oracle import ref_garf  # noqa: E402


def _flat_and_layers(net, names, gauss_names):
    """Flat fp32 parameter vector of `net` and GaussLinear descriptors of the named Linear layers."""
    from nerf_experiments_b200.garf_program import GaussLinear
    from nerf_experiments_b200.mlp_program import Linear
    params = list(net.parameters())
    offs, o = {}, 0
    for p in params:
        offs[id(p)] = o
        o += p.numel()
    flat = th.cat([p.detach().reshape(-1) for p in params])
    mods = dict(net.named_modules())
    layers = []
    for ln, gn in zip(names, gauss_names):
        m = mods[ln]
        g = offs[id(mods[gn].inv_standard_deviation)] if gn else -1
        layers.append(GaussLinear(Linear(offs[id(m.weight)], offs[id(m.bias)], m.out_features, m.in_features), g))
    return flat, layers, offs


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_radiance_program_matches_oracle():
    from nerf_experiments_b200 import garf_program as gp
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    th.manual_seed(77)
    net = RadianceNetwork(0.5, 1.5)
    names = [f"model_density_1.{k}" for k in (0, 2, 4, 6)] + [f"model_density_2.{k}" for k in (0, 2, 4, 6)]
    gn = [f"model_density_1.{k}" for k in (1, 3, 5, 7)] + [f"model_density_2.{k}" for k in (1, 3, 5)] + [None]
    flat, L, offs = _flat_and_layers(net, names + ["model_color.0", "model_color.2"], gn + ["model_color.1", None])
    cg = gp.compile_radiance(L[:8], L[8:])
    N = 200                                         # two tiles, the second one ragged
    g = th.Generator().manual_seed(3)
    pos = th.randn((N, 3), generator=g) * 1.2
    dirs = th.nn.functional.normalize(th.randn((N, 3), generator=g), dim=1)
    up_s, up_c = th.randn(N, generator=g), th.randn((N, 3), generator=g)
    sigma, rgb, grad, dpos, ddir = garf_sim.run_network(cg, flat, pos, dirs, up_s, up_c)

    sd = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items()}
    p_ref, d_ref = pos.clone().requires_grad_(), dirs.clone().requires_grad_()
    r_rgb, r_sigma = ref_garf.radiance_network(sd, p_ref, d_ref)
    ((r_rgb * up_c).sum() + (r_sigma * up_s).sum()).backward()
    assert (rgb - r_rgb.detach()).abs().max() < 1e-2            # north_star: bf16-MLP rgb within 1e-2 abs
    assert (sigma - r_sigma.detach()).abs().max() < 2e-2 * (1 + r_sigma.detach().abs().max())
    for name, p in net.named_parameters():
        o = offs[id(p)]
        got = grad[o: o + p.numel()].view(p.shape)
        assert _rel(got, sd[name].grad) < 4e-2, name
    assert _rel(dpos, p_ref.grad) < 4e-2 and _rel(ddir, d_ref.grad) < 4e-2


def test_proposal_program_matches_oracle():
    from nerf_experiments_b200 import garf_program as gp
    from nerf_experiments_b200.model_garf_proposal import ProposalNetwork
    th.manual_seed(77)
    net = ProposalNetwork(0.5, 1.5)
    flat, L, offs = _flat_and_layers(net, [f"model.{k}" for k in (0, 2, 4, 6)], ["model.1", "model.3", "model.5", None])
    cg = gp.compile_proposal(L)
    N = 130
    g = th.Generator().manual_seed(4)
    pos = th.randn((N, 3), generator=g) * 1.2
    up_s = th.randn(N, generator=g)
    sigma, rgb, grad, dpos, _ = garf_sim.run_network(cg, flat, pos, None, up_s, None)
    assert rgb is None
    sd = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items()}
    p_ref = pos.clone().requires_grad_()
    r_sigma = ref_garf.proposal_network(sd, p_ref)[:, 0]
    (r_sigma * up_s).sum().backward()
    assert (sigma - r_sigma.detach()).abs().max() < 2e-2 * (1 + r_sigma.detach().abs().max())
    for name, p in net.named_parameters():
        o = offs[id(p)]
        assert _rel(grad[o: o + p.numel()].view(p.shape), sd[name].grad) < 4e-2, name
    assert _rel(dpos, p_ref.grad) < 4e-2


def test_program_invariants():
    """Structural rules the kernels rely on: a step that rewrites slabs or reads an accumulator waits for
    the op that used them (wait_lag), every op has one step in front, the last step waits for the last op."""
    from nerf_experiments_b200 import _lib
    from nerf_experiments_b200 import garf_program as gp
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    th.manual_seed(1)
    net = RadianceNetwork(0.5, 1.5)
    names = [f"model_density_1.{k}" for k in (0, 2, 4, 6)] + [f"model_density_2.{k}" for k in (0, 2, 4, 6)]
    gn = [f"model_density_1.{k}" for k in (1, 3, 5, 7)] + [f"model_density_2.{k}" for k in (1, 3, 5)] + [None]
    _, L, _ = _flat_and_layers(net, names + ["model_color.0", "model_color.2"], gn + ["model_color.1", None])
    cg = gp.compile_radiance(L[:8], L[8:])
    for prog in (cg.fwd, cg.bwd):
        assert prog.steps[prog.n_ops].wait_lag == 0
        readers = {}                                   # slab -> index of the last op that reads it
        for k in range(prog.n_ops + 1):
            st = prog.steps[k]
            if st.out_slab >= 0 and not (st.flags & _lib.NG_F_DIRECT):
                n_out = st.n_slabs + (1 if (st.kind == _lib.NG_BSTEP_PLAIN and st.flags & _lib.NG_F_SIGMA) else 0)
                for s in range(st.out_slab, st.out_slab + n_out):
                    assert readers.get(s, -10) <= k - 1 - st.wait_lag, (k, s)
            if k < prog.n_ops:
                op = prog.ops[k]
                for c in range(op.n_chunks):
                    readers[op.a_slab[c]] = k
    assert cg.fwd.n_floats <= _lib.NG_MAX_PROGRAM_FLOATS and cg.bwd.n_floats <= _lib.NG_MAX_PROGRAM_FLOATS
