"""Oracle: GARF / SARF / Gabor activations with gradients (a6, a9), the GARF radiance and
proposal networks (a7, a8) and the PropNet sampling / rendering / loss chain (a12) in plain
PyTorch fp32 on the CPU.  Test infrastructure only.

The activations and networks restate reference barf/gaussian.py:8-63, sarf/activation.py:63-65,
gaborf/gabor.py:8-29, garf/model_radiance.py:84-96 and garf/model_proposal.py:55-56 and are
pinned by tests/golden/garf.npz (outputs and gradients of the unmodified reference modules).

PARITY UNPINNED for `garf_forward` / `proposal_loss`: they restate the nerfacc 0.5.x algorithm the
reference calls at garf/model_garf.py:210-230,257 (see oracle/ref_nerfacc.py); the package is
neither in /root/reference nor installed, and the reference holds no fixture at that boundary."""
import numpy as np
import torch as th
import torch.nn.functional as F

from . import ref_nerfacc, ref_render


def gauss_act(x, inv_std):
    return th.exp(-x ** 2 * (inv_std ** 2 + 1e-6))


def sarf_act(x, f):
    xs = (th.signbit(x) * 2 - 1) * (x.abs() + 1e-4)
    return th.cos(f / (xs ** 2 + 1 / f ** 2)) * th.exp(-xs ** 2)


def gabor_act(x, inv_std, spread):
    return th.exp(-(inv_std ** 2 + 1e-6) * x ** 2) * th.cos(spread * x)


def softplus8(x):
    return F.softplus(x, beta=1.0, threshold=8.0)


def _seq(sd, prefix, x, n_linear, last_act):
    """Linear/GaussAct chain `prefix.{0,1,2,...}`; the last Linear is followed by a Gaussian only
    if last_act."""
    for k in range(n_linear):
        x = F.linear(x, sd[f"{prefix}.{2 * k}.weight"], sd[f"{prefix}.{2 * k}.bias"])
        if k < n_linear - 1 or last_act:
            x = gauss_act(x, sd[f"{prefix}.{2 * k + 1}.inv_standard_deviation"])
    return x


def radiance_network(sd, pos, dir):
    """reference garf/model_radiance.py:84-96 -> (rgb (N,3), density (N,))."""
    z1 = _seq(sd, "model_density_1", pos, 4, True)
    z2 = _seq(sd, "model_density_2", th.cat((z1, pos), dim=1), 4, False)
    density = softplus8(z2[:, 128] - 1)
    rgb = th.sigmoid(_seq(sd, "model_color", th.cat((z1[:, :128] + z2[:, :128], dir), dim=1), 2, False))
    return rgb, density


def proposal_network(sd, pos):
    """reference garf/model_proposal.py:55-56 -> (N,1)."""
    return softplus8(_seq(sd, "model", pos, 4, False))


def transmittance_cdf(sigma, t0, t1):
    sd = sigma * (t1 - t0)
    trans = th.exp(-(th.cumsum(sd, dim=1) - sd))
    return trans, 1.0 - th.cat((trans, th.zeros_like(trans[:, :1])), dim=1)


def pdf_outer_loss(t_q, cdf_q, t_k, cdf_k, eps=1e-7):
    ids_right = th.searchsorted(t_k.contiguous(), t_q.contiguous(), right=False).clamp(0, t_k.shape[1] - 1)
    ids_left = (th.searchsorted(t_k.contiguous(), t_q.contiguous(), right=True) - 1).clamp(0, t_k.shape[1] - 1)
    w = cdf_q[:, 1:] - cdf_q[:, :-1]
    w_outer = cdf_k.gather(1, ids_right[:, 1:]) - cdf_k.gather(1, ids_left[:, :-1])
    return th.clip(w - w_outer, min=0) ** 2 / (w + eps)


def garf_forward(sd_prop, sd_rad, o, d, near, far, n_prop, n_rad, u_prop=None, u_rad=None):
    """The chain of reference garf/model_garf.py:194-236 with explicit uniforms.
    Returns rgb, opacity, depth, proposal loss and the sample intervals."""
    B = o.shape[0]
    s_edges = np.tile(np.array([0.0, 1.0], dtype=np.float32), (B, 1))
    s1, _ = ref_nerfacc.importance_sampling(s_edges, s_edges, n_prop, None if u_prop is None else u_prop.numpy())
    t = ref_nerfacc.lindisp_s_to_t(th.from_numpy(s1), near, far)
    t0, t1 = t[:, :-1], t[:, 1:]
    pos = o[:, None] + d[:, None] * (t0 + t1)[..., None] / 2
    sigma_p = proposal_network(sd_prop, pos.reshape(-1, 3)).view(t0.shape)
    _, cdf_p = transmittance_cdf(sigma_p, t0, t1)
    s2, _ = ref_nerfacc.importance_sampling(s1, cdf_p.detach().numpy(), n_rad, None if u_rad is None else u_rad.numpy())
    tr = ref_nerfacc.lindisp_s_to_t(th.from_numpy(s2), near, far)
    r0, r1 = tr[:, :-1], tr[:, 1:]
    pos = o[:, None] + d[:, None] * (r0 + r1)[..., None] / 2
    rgb_s, sigma = radiance_network(sd_rad, pos.reshape(-1, 3), d.repeat_interleave(n_rad, dim=0))
    rgb_s, sigma = rgb_s.view(B, n_rad, 3), sigma.view(B, n_rad)
    rgb, opacity, depth, w, trans, _ = ref_render.render_rays_nerfacc(sigma, rgb_s, r0, r1)
    cdf_q = 1.0 - th.cat((trans.detach(), th.zeros_like(trans[:, :1])), dim=1)
    loss_p = pdf_outer_loss(tr, cdf_q, t, cdf_p).mean()
    return rgb, opacity, depth, loss_p, (r0, r1)
