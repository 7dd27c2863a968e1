"""Oracle: one BARF / vanilla training step (forward + loss + backward) on the CPU, composed
from the restated reference functions — the arithmetic of NerfInterpolation.forward +
_step_helper (reference barf/model_interpolation.py:417-526) with CameraExtrinsics.forward in
front (barf/model_camera_calibration.py:322-326).  Test infrastructure and the CPU-baseline leg
of bench.py only."""
import torch as th
import torch.nn.functional as F

from . import ref_mlp, ref_pe, ref_pose, ref_render, ref_resample, ref_sampling


def field(sd, cfg, pe_cfg, o, d, t_start, t_end, strategy: str, emulate_bf16: bool = False):
    """sigma (B,S), rgb (B,S,3) of one network for rays sampled at (t_start, t_end)."""
    B, S = t_start.shape
    pos, dirs = ref_sampling.compute_positions(o, d, t_start, t_end, strategy)
    pos, dirs = pos.reshape(B * S, 3), dirs.reshape(B * S, 3)
    P = ref_pe.barf_encoding(pos, pe_cfg["pos_levels"], pe_cfg["scale"], pe_cfg["identity"], pe_cfg.get("alpha_pos"))
    D = ref_pe.barf_encoding(dirs, pe_cfg["dir_levels"], pe_cfg["scale"], pe_cfg["identity"], pe_cfg.get("alpha_dir"))
    sigma, rgb = ref_mlp.nerf_model_forward(sd, cfg, P, D, emulate_bf16=emulate_bf16)
    return sigma.view(B, S), rgb.view(B, S, 3)


def render(sd_rad, cfg, pe_cfg, o, d, near, far, n_rad, strategy, uniforms, sd_prop=None, n_prop=0,
           sampling="equidistant", offset_size=-1.0, emulate_bf16=False):
    """rgb_fine, rgb_coarse (or None), weights of the last pass.  uniforms: dict with optional
    'jitter' (B,S0) and 'offset' (B,1) for the first (uniform) sampling pass."""
    B = o.shape[0]
    s0 = n_prop if n_prop > 0 else n_rad
    jitter = uniforms.get("jitter") if sampling == "stratified_uniform" else None
    t0, t1 = ref_sampling.sample_uniform(near, far, B, s0, jitter, uniforms.get("offset"), offset_size)
    if n_prop > 0:
        sig, col = field(sd_prop, cfg, pe_cfg, o, d, t0, t1, strategy, emulate_bf16)
        rgb_c, w = ref_render.render_rays(sig, col, t1 - t0)
        f0, f1, _, _ = ref_resample.sample_pdf_weighted(t0.numpy(), w.detach().numpy(), (t1 - t0).numpy(), n_rad,
                                                        near, far, uniforms.get("fallback"))
        t0, t1 = th.from_numpy(f0), th.from_numpy(f1)
    else:
        rgb_c = None
    sig, col = field(sd_rad, cfg, pe_cfg, o, d, t0, t1, strategy, emulate_bf16)
    rgb_f, w = ref_render.render_rays(sig, col, t1 - t0)
    return rgb_f, rgb_c, w


def barf_step(sd_rad, cfg, pe_cfg, rotation, translation, img_idx, o, d, target, near, far, n_rad,
              strategy="middle", uniforms=None, sd_prop=None, n_prop=0, sampling="equidistant",
              offset_size=-1.0, emulate_bf16=False):
    """loss (scalar tensor with graph), rgb_fine.  Call .backward() on the loss for gradients of
    sd_rad / sd_prop / rotation / translation (leaf tensors requiring grad)."""
    uniforms = uniforms or {}
    if rotation is not None:
        o, d, _, _ = ref_pose.pose_forward(rotation, translation, img_idx.long(), o, d)
    rgb_f, rgb_c, _ = render(sd_rad, cfg, pe_cfg, o, d, near, far, n_rad, strategy, uniforms, sd_prop, n_prop,
                             sampling, offset_size, emulate_bf16)
    loss = F.mse_loss(rgb_f, target)
    if rgb_c is not None:
        loss = loss + F.mse_loss(rgb_c, target)
    return loss, rgb_f
