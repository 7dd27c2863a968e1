"""Oracle: positional encodings (a3, a4).  Test infrastructure only."""
import math

import torch as th


def barf_alpha(epoch: float, levels: int, alpha_start: float, start_epoch: float, end_epoch: float) -> float:
    """reference barf/positional_encodings.py:84-103 (update_alpha)."""
    if epoch < start_epoch:
        return alpha_start
    if start_epoch <= epoch < end_epoch:
        return alpha_start + (epoch - start_epoch) * (levels - alpha_start) / (end_epoch - start_epoch)
    return float(levels)


def barf_mask(alpha: th.Tensor, levels: int, space_dimensions: int = 3) -> th.Tensor:
    """reference barf/positional_encodings.py:105-122 (compute_mask)."""
    alpha = th.as_tensor(alpha, dtype=th.float32)
    mask = th.zeros(levels)
    idx_ramp = int(alpha)
    mask[:idx_ramp] = 1.0
    if idx_ramp < levels:
        mask[idx_ramp] = (1 - th.cos((alpha - idx_ramp) * th.pi)) / 2
    return mask.repeat(space_dimensions).view(1, -1)


def fourier_features(x: th.Tensor, levels: int, scale: float) -> th.Tensor:
    """reference barf/positional_encodings.py:43-57 (FourierFeatures.forward)."""
    sc = scale * (2 ** th.arange(levels)).repeat(x.shape[1])
    args = x.repeat_interleave(levels, dim=1) * sc
    return th.hstack((th.cos(args), th.sin(args)))


def barf_encoding(x: th.Tensor, levels: int, scale: float, include_identity: bool, alpha=None) -> th.Tensor:
    """reference barf/positional_encodings.py:124-148 (BarfPositionalEncoding.forward);
    alpha=None means no mask (FourierFeatures, :43-57)."""
    sc = scale * (2 ** th.arange(levels)).repeat(x.shape[1])
    args = x.repeat_interleave(levels, dim=1) * sc
    mask = barf_mask(alpha, levels, x.shape[1]) if alpha is not None else 1.0
    parts = (mask * th.cos(args), mask * th.sin(args))
    if include_identity:
        parts = (x,) + parts
    return th.hstack(parts)


def integrated_encoding(pos, dir, pixel_width, t_start, t_end, levels: int, scale: float,
                        include_identity: bool, distribute_variance: bool,
                        pixel_width_sigma: float, alpha=None) -> th.Tensor:
    """reference barf/positional_encodings.py:170-240 (IntegratedFourierFeatures.forward) and,
    with alpha given, :266-282 (IntegratedBarfFourierFeatures.forward)."""
    space_dim = 3
    t_mu = (t_start + t_end) / 2
    t_delta = (t_end - t_start) / 2
    mu_diff = 2 * t_mu * t_delta ** 2 / (3 * t_mu ** 2 + t_delta ** 2)
    pos_mu = pos + mu_diff * dir
    r_dot = pixel_width * 2 / (12 ** 0.5)
    sigma_t_sq = t_delta ** 2 / 3 - (4 * t_delta ** 4 * (12 * t_mu ** 2 - t_delta ** 2)) / (15 * (3 * t_mu ** 2 + t_delta ** 2) ** 2)
    sigma_r_sq = r_dot ** 2 * (t_mu ** 2 / 4 + 5 * t_delta ** 2 / 12 - 4 * t_delta ** 4 / (15 * (3 * t_mu ** 2 + t_delta ** 2)))
    add_sigma = (pixel_width_sigma * pixel_width * t_mu) ** 2 if pixel_width_sigma > 0.25 else 0.0
    sigma_t_sq = sigma_t_sq + add_sigma
    sigma_r_sq = sigma_r_sq + add_sigma
    sc = 4 ** th.arange(levels).repeat(space_dim)
    if distribute_variance:
        Sigma = (sigma_t_sq + sigma_r_sq * 2) / space_dim * sc
        weight = th.exp(-Sigma / 2)
    else:
        diag = sigma_t_sq * dir ** 2 + sigma_r_sq * (1 - dir ** 2 / th.sum(dir ** 2, dim=1, keepdim=True))
        weight = th.exp(-(diag.repeat_interleave(levels, dim=1) * sc) / 2)
    pe = fourier_features(pos_mu, levels, scale)
    ipe = pe * weight.repeat(1, 2)
    if alpha is not None:
        m = barf_mask(alpha, levels, 3)
        ipe = ipe * m.repeat(1, 2)
    if include_identity:
        ipe = th.cat((pos_mu, ipe), dim=1)
    return ipe


def mip_sigma_factor(step: float, start: float, end: float, sigma_blur0: float, sigma_pw0: float) -> float:
    """Mip-BARF sigma schedule factor, reference barf/model_mip.py:184-225:
    1 before `start`; (0.25/max(s0))**((start-step)/(start-end)) inside; 0 after."""
    if step < start:
        return 1.0
    if step > end:
        return 0.0
    return (0.25 / max(sigma_blur0, sigma_pw0)) ** ((start - step) / (start - end))
