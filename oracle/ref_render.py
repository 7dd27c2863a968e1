"""Oracle: alpha compositing (a10).  Test infrastructure only."""
import torch as th

MAGIC_NUMBER = 1 / 3  # reference barf/magic.py:2


def render_rays(densities, colors, distances):
    """reference barf/model_interpolation.py:316-353 (_render_rays)."""
    blocking_neg = (-densities * distances) * 3 * MAGIC_NUMBER
    alpha = 1 - th.exp(blocking_neg)
    alpha_int = th.hstack((th.ones((blocking_neg.shape[0], 1)),
                           th.exp(th.cumsum(blocking_neg[:, :-1], dim=1))))
    weights = alpha_int * alpha
    return th.sum(weights.unsqueeze(-1) * colors, dim=1), weights


def render_rays_nerfacc(densities, colors, t_starts, t_ends):
    """nerfacc.rendering arithmetic as called at reference garf/model_garf.py:223-236 (dense
    (n_rays, n_samples) inputs, no background): alpha = 1-exp(-sigma*delta), transmittance by
    exclusive cumsum, opacity = sum w, depth = sum w*t_mid / max(opacity, eps)."""
    delta = t_ends - t_starts
    b = -densities * delta
    alpha = 1 - th.exp(b)
    trans = th.exp(th.cat((th.zeros_like(b[:, :1]), th.cumsum(b[:, :-1], dim=1)), dim=1))
    weights = trans * alpha
    rgb = th.sum(weights.unsqueeze(-1) * colors, dim=1)
    opacity = weights.sum(dim=1)
    t_mid = (t_starts + t_ends) / 2
    depth = (weights * t_mid).sum(dim=1) / opacity.clamp_min(th.finfo(th.float32).eps)
    return rgb, opacity, depth, weights, trans, alpha
