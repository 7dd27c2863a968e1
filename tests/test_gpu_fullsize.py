"""The kernels at BASELINE.json's full sizes, checked through size-independent properties (the CPU
oracle would take minutes there): a second GPU formulation in plain torch ops, linearity,
sortedness, conservation of the sample count, invariance of a sample's result to the batch it
sits in, and agreement of a random subset with the fp32 reference arithmetic."""
import pytest
import torch as th

from oracle import ref_mlp, ref_pe

pytestmark = pytest.mark.gpu


def _ops():
    from nerf_experiments_b200 import ops
    return ops


def test_compositing_at_c2_size(cuda):
    """4096 rays x 128 samples (C2) and 8192 x 256 (C3 fine pass)."""
    ops = _ops()
    for B, S in ((4096, 128), (8192, 256)):
        g = th.Generator(device=cuda).manual_seed(S)
        sigma = th.nn.functional.softplus(th.randn((B, S), device=cuda, generator=g) * 2)
        delta = th.rand((B, S), device=cuda, generator=g) * 0.1 + 0.01
        c1 = th.rand((B, S, 3), device=cuda, generator=g)
        c2 = th.rand((B, S, 3), device=cuda, generator=g)
        rgb1, w, _, _ = ops.composite_fwd(sigma, delta, c1)
        rgb2, w2, _, _ = ops.composite_fwd(sigma, delta, c2)
        # second formulation (torch ops on the same device): the reference's own expression
        b = ((-sigma * delta) * 3) * (1 / 3)
        T = th.cat((th.ones((B, 1), device=cuda), th.exp(th.cumsum(b[:, :-1], dim=1))), dim=1)
        w_ref = T * (1 - th.exp(b))
        assert (w - w_ref).abs().max() < 1e-5 and th.equal(w, w2)
        assert (rgb1 - (w_ref.unsqueeze(-1) * c1).sum(1)).abs().max() < 1e-5
        assert float(w.min()) >= 0.0 and float(w.sum(1).max()) <= 1.0 + 1e-5
        # linear in the colours
        rgb12, _, _, _ = ops.composite_fwd(sigma, delta, 0.25 * c1 + 0.75 * c2)
        assert (rgb12 - (0.25 * rgb1 + 0.75 * rgb2)).abs().max() < 1e-5


def test_allocator_resampling_at_c3_size(cuda):
    """8192 rays, 64 -> 256 samples (C3): every ray gets exactly Sf samples, every coarse bin at
    least its own sample, the fine samples are sorted, start at the coarse ray start and end at far."""
    ops = _ops()
    B, Sc, Sf = 8192, 64, 256
    g = th.Generator(device=cuda).manual_seed(0)
    t0, t1 = ops.sample_uniform(2.0, 8.0, B, Sc, cuda, None, th.rand((B, 1), device=cuda, generator=g), -1.0)
    w = th.rand((B, Sc), device=cuda, generator=g) ** 6
    f0, f1, counts, flag = ops.resample_alloc(t0, w, t1 - t0, Sf, 2.0, 8.0, want_counts=True)
    assert int(flag.item()) == 0
    assert th.equal(counts.sum(1), th.full((B,), Sf, device=cuda, dtype=counts.dtype)) and int(counts.min()) >= 1
    assert bool((f0[:, 1:] >= f0[:, :-1]).all()) and th.equal(f0[:, 0], t0[:, 0])
    assert th.equal(f1[:, :-1], f0[:, 1:]) and bool((f1[:, -1] == 8.0).all())
    # bin i holds counts[i] samples starting at the coarse sample itself
    first = th.cat((th.zeros((B, 1), device=cuda, dtype=th.long), counts.long().cumsum(1)[:, :-1]), 1)
    assert th.equal(th.gather(f0, 1, first), t0)
    # heavier bins never get fewer samples than lighter ones (up to the +-1 of the remainder rule)
    order = w.argsort(1)
    sorted_counts = th.gather(counts.long(), 1, order)
    assert int((sorted_counts[:, 1:] - sorted_counts[:, :-1]).min()) >= -1


def test_inverse_cdf_resampling_at_c4_size(cuda):
    """16384 rays, 64 -> 192 samples (C4): bins non-decreasing and exactly the searchsorted ones,
    edges sorted inside the ray's range."""
    ops = _ops()
    B, Sc, Sf = 16384, 64, 192
    g = th.Generator(device=cuda).manual_seed(1)
    edges = th.linspace(0, 1, Sc + 1, device=cuda).repeat(B, 1)
    w = th.rand((B, Sc), device=cuda, generator=g) ** 4
    cdf = th.cat((th.zeros((B, 1), device=cuda), w.cumsum(1)), 1)
    cdf = cdf / cdf[:, -1:]
    u = th.rand((B,), device=cuda, generator=g)
    out, idx = ops.resample_icdf(edges, cdf, Sf, u, want_idx=True)
    assert bool((out[:, 1:] >= out[:, :-1]).all()) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert bool((idx[:, 1:] >= idx[:, :-1]).all()) and int(idx.min()) >= 0 and int(idx.max()) <= Sc - 1
    # the same bins from torch.searchsorted on the same targets (computed with the kernel's formula)
    k = th.arange(Sf, device=cuda, dtype=th.float32)
    step = (cdf[:, -1:] - cdf[:, :1]) / Sf
    targets = cdf[:, :1] + (k.unsqueeze(0) + u.unsqueeze(1)) * step
    ref = (th.searchsorted(cdf.contiguous(), targets.contiguous(), right=True) - 1).clamp(0, Sc - 1)
    assert float((ref == idx.long()).float().mean()) > 0.9999     # fma contraction in torch's broadcast may move a tie


def test_fused_field_at_c2_size(cuda):
    """524 288 samples through the fused forward: a sample's result does not depend on the launch it
    is part of (the two halves launched separately give the same bits), and a random subset agrees
    with the fp32 reference arithmetic within the north-star 1e-2."""
    import bench
    from nerf_experiments_b200.field_function import field_rays
    model = bench.build_model(20).to(cuda)
    net = model.model_radiance
    B, S = 4096, 128
    g = th.Generator().manual_seed(0)
    o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(cuda)
    d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(cuda)
    ops = _ops()
    t0, t1 = ops.sample_uniform(2.0, 8.0, B, S, cuda, None, th.rand((B, 1), generator=g).to(cuda), -1.0)
    pw = th.full((B, 1), 1 / 555.0, device=cuda)
    with th.no_grad():
        sig, rgb = field_rays(net, o, d, t0, t1, pw, "middle")
        h = B // 2
        sig_a, rgb_a = field_rays(net, o[:h], d[:h], t0[:h], t1[:h], pw[:h], "middle")
        sig_b, rgb_b = field_rays(net, o[h:], d[h:], t0[h:], t1[h:], pw[h:], "middle")
    assert th.equal(sig, th.cat((sig_a, sig_b))) and th.equal(rgb, th.cat((rgb_a, rgb_b)))
    assert bool(th.isfinite(sig).all()) and float(sig.min()) >= 0.0 and float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0
    # random subset against the fp32 reference arithmetic (CPU)
    pick = th.randperm(B, generator=g)[:16]
    oc, dc, tq = o[pick].cpu(), d[pick].cpu(), ((t0 + t1) / 2)[pick].cpu()
    pos = (oc.unsqueeze(1) + tq.unsqueeze(2) * dc.unsqueeze(1)).reshape(-1, 3)
    dirs = dc.unsqueeze(1).expand(-1, S, -1).reshape(-1, 3)
    ep, ed = net.position_encoder, net.direction_encoder
    P = ref_pe.barf_encoding(pos, 10, 1.0, True, ep.alpha.detach().cpu())
    D = ref_pe.barf_encoding(dirs, 4, 1.0, True, ed.alpha.detach().cpu())
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    cfg = dict(n_hidden=4, n_segments=2, delayed_direction=True, delayed_density=False)
    r_sig, r_rgb = ref_mlp.nerf_model_forward(sd, cfg, P, D)
    assert (rgb[pick].cpu().reshape(-1, 3) - r_rgb).abs().max() < 1e-2
    assert (sig[pick].cpu().reshape(-1) - r_sig).abs().max() < 2e-2 * max(1.0, float(r_sig.abs().max()))


def test_garf_fields_at_c4_size(cuda):
    """BASELINE C4 (4096 rays x 64 proposal / 192 radiance samples) through the fused GARF kernels:
    the forward of a sample is bit-identical whether it is launched with the whole batch or with half of
    it; a random subset agrees with the fp32 oracle arithmetic (rgb 1e-2, density 2 %); the parameter
    gradients of the whole batch equal the sum of the gradients of its two halves and are linear in the
    upstream gradient (relative L2 2e-3: bf16 stashes are identical, only the summation order differs)."""
    from oracle import ref_garf
    from nerf_experiments_b200.model_garf_proposal import ProposalNetwork
    from nerf_experiments_b200.model_garf_radiance import RadianceNetwork
    th.manual_seed(77)
    prop, rad = ProposalNetwork(0.5, 1.5).to(cuda), RadianceNetwork(0.5, 1.5).to(cuda)
    g = th.Generator().manual_seed(4)

    def grads(net):
        out = th.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
        for p in net.parameters():
            p.grad = None
        return out

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-20))

    for net, n, has_dir in ((rad, 4096 * 192, True), (prop, 4096 * 64, False)):
        pos = (th.randn((n, 3), generator=g) * 1.5).to(cuda)
        dirs = th.nn.functional.normalize(th.randn((n, 3), generator=g), dim=1).to(cuda)
        up_s = th.randn(n, generator=g).to(cuda) / n
        up_c = th.randn((n, 3), generator=g).to(cuda) / n
        h = n // 2

        def run(sl, scale=1.0, want_grad=True):
            if has_dir:
                rgb, dens = net(pos[sl], dirs[sl])
                loss = (rgb * up_c[sl]).sum() + (dens * up_s[sl]).sum()
            else:
                dens = net(pos[sl])[:, 0]
                rgb = None
                loss = (dens * up_s[sl]).sum()
            if not want_grad:
                return rgb, dens, None
            (loss * scale).backward()
            return (rgb.detach() if has_dir else None), dens.detach(), grads(net)

        rgb, dens, g_all = run(slice(0, n))
        rgb_a, dens_a, g_a = run(slice(0, h))
        rgb_b, dens_b, g_b = run(slice(h, n))
        assert th.equal(dens, th.cat((dens_a, dens_b)))
        if has_dir:
            assert th.equal(rgb, th.cat((rgb_a, rgb_b)))
            assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0
        assert bool(th.isfinite(dens).all()) and float(dens.min()) >= 0.0 and bool(th.isfinite(g_all).all())
        assert rel(g_a + g_b, g_all) < 2e-3
        _, _, g_2 = run(slice(0, n), scale=2.0)
        assert rel(g_2, 2.0 * g_all) < 2e-3
        # a random subset against the fp32 oracle
        pick = th.randperm(n, generator=g)[:2048].to(cuda)
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        if has_dir:
            r_rgb, r_dens = ref_garf.radiance_network(sd, pos[pick].cpu(), dirs[pick].cpu())
            assert (rgb[pick].cpu() - r_rgb).abs().max() < 1e-2
        else:
            r_dens = ref_garf.proposal_network(sd, pos[pick].cpu())[:, 0]
        assert (dens[pick].cpu() - r_dens).abs().max() < 2e-2 * max(1.0, float(r_dens.abs().max()))
