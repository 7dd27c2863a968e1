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


def _compiled_networks():
    from nerf_experiments_b200.model_garf_proposal import ProposalNetwork
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    th.manual_seed(1)
    out = []
    for net in (RadianceNetwork(0.5, 1.5), ProposalNetwork(0.5, 1.5)):
        f = net.fused_field()
        out.append((net, f, f.compile_fn(f.flat)))
    return out


def test_early_ops_never_write_the_columns_their_step_reads():
    """NgOp.early (the op's MMAs start slab by slab under the epilogue of the step in front of it) is only set
    where the op's accumulator blocks are disjoint from the columns that step still reads, and every Gaussian
    step that publishes slabs got it (the TMEM regions of consecutive layers alternate for that)."""
    from nerf_experiments_b200 import _lib
    for _, _, cg in _compiled_networks():
        for prog in (cg.fwd, cg.bwd):
            for k in range(prog.n_ops):
                st, op = prog.steps[k], prog.ops[k]
                publishes = st.kind in (_lib.NG_STEP_ACT, _lib.NG_BSTEP_ACT) and st.out_slab >= 0 and \
                    not (st.flags & _lib.NG_F_DIRECT) and op.n_chunks > 0
                assert bool(op.early) == publishes, (k, st.kind, op.early)
                if op.early:
                    lo, hi = st.src_col, st.src_col + 64 * st.n_slabs
                    for b in range(op.n_blocks):
                        c0, c1 = op.blocks[b].tmem_col, op.blocks[b].tmem_col + op.blocks[b].n
                        assert c1 <= lo or hi <= c0, (k, (lo, hi), (c0, c1))


def test_weight_gradient_units_cover_every_parameter_once():
    """Every weight element belongs to exactly one unit, every bias (= every dz slab's column sums) to exactly
    one unit, a concatenating layer is ONE unit per 256 output features (x2_slab), no unit reads the z stash,
    and every Gaussian layer is listed for nerfb200_gauss_width_grad."""
    from nerf_experiments_b200 import _lib
    for net, f, cg in _compiled_networks():
        n_params = f.flat.numel
        w_cover, b_cover = th.zeros(n_params, dtype=th.int32), th.zeros(n_params, dtype=th.int32)
        for u in cg.units:
            assert u.mode == _lib.WGRAD_MMA and u.n_z_slabs == 0 and u.z_slab < 0
            assert 2 * ((u.n_dy_slabs + 1) // 2) + u.n_x_slabs <= 9 and u.n_real <= 64 * u.n_x_slabs
            for m in range(u.m_real):
                w_cover[u.dst + m * u.ld: u.dst + m * u.ld + u.n_real] += 1
            if u.bias_dst >= 0:
                b_cover[u.bias_dst: u.bias_dst + u.m_real] += 1
        n_gauss = 0
        for name, p in net.named_parameters():
            o = f.flat.offset_of(p)
            if name.endswith("weight"):
                assert bool((w_cover[o: o + p.numel()] == 1).all()), name
            elif name.endswith("bias"):
                assert bool((b_cover[o: o + p.numel()] == 1).all()), name
            else:
                n_gauss += p.numel()
                assert any(l.g_off == o and l.lin.out_f == p.numel() for l in cg.gauss_layers), name
        assert n_gauss == sum(l.lin.out_f for l in cg.gauss_layers) == cg.gauss_per_sample
    rad_units = _compiled_networks()[0][2].units
    assert sum(1 for u in rad_units if u.x2_slab >= 0) == 3          # 131 -> 512 (two units) and 131 -> 256


def test_weight_gradient_schedule_covers_every_tile_once():
    """schedule_wgrad: the items of a unit tile its range exactly once, whatever the number of items per
    worker; the boustrophedon order only permutes them."""
    from nerf_experiments_b200.mlp_program import schedule_wgrad
    cg = _compiled_networks()[0][2]
    for n_tiles, per_worker in ((6144, 4), (37, 3), (1, 8)):
        items = schedule_wgrad(cg.units, n_tiles, 148, per_worker)
        cover = {}
        for it in items:
            key = (it.dy_slab, it.x_slab, it.x2_slab, it.dst)
            cover.setdefault(key, []).append((it.tile_begin, it.tile_end))
        assert len(cover) == len(cg.units)
        for spans in cover.values():
            spans.sort()
            assert spans[0][0] == 0 and spans[-1][1] == n_tiles
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
