"""Thin torch wrappers over the C ABI (include/nerfb200.h): argument checking, raw pointers and
the current CUDA stream in; tensors out.  The autograd Functions give the ops the gradients the
reference obtains from PyTorch autograd.  CUDA only — there is no CPU path."""
from typing import Optional

import torch as th

from . import _lib
from ._lib import check, lib


def _ptr(t: Optional[th.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return th.cuda.current_stream().cuda_stream


def _f32(t: th.Tensor, name: str, shape=None) -> th.Tensor:
    if not isinstance(t, th.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (nerfb200 has no CPU fallback)")
    if t.dtype != th.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _i32(t: th.Tensor, name: str) -> th.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor")
    return t.to(th.int32).contiguous()


# ---------------------------------------------------------------------------------------------
# a1 uniform sampling
# ---------------------------------------------------------------------------------------------
def sample_uniform(near: float, far: float, batch: int, n_samples: int, device,
                   jitter: Optional[th.Tensor] = None, offset_u: Optional[th.Tensor] = None,
                   offset_size: float = 0.0):
    """t_start, t_end (B,S) — barf/model_interpolation.py:135-180 with explicit uniforms."""
    t_start = th.empty((batch, n_samples), device=device, dtype=th.float32)
    t_end = th.empty_like(t_start)
    if jitter is not None:
        jitter = _f32(jitter, "jitter", (batch, n_samples))
    if offset_u is not None:
        offset_u = _f32(offset_u.reshape(-1), "offset_u", (batch,))
    with th.cuda.device(t_start.device):
        check(lib().nerfb200_sample_uniform(float(near), float(far), batch, n_samples, _ptr(jitter),
                                            _ptr(offset_u), float(offset_size), _ptr(t_start),
                                            _ptr(t_end), _stream()), "sample_uniform")
    return t_start, t_end


# ---------------------------------------------------------------------------------------------
# a10 compositing
# ---------------------------------------------------------------------------------------------
def composite_fwd(sigma, delta, rgb, t_mid=None, flavour=_lib.COMPOSITE_BARF, want_w=True,
                  want_opacity=False, want_depth=False):
    B, S = sigma.shape
    sigma = _f32(sigma, "sigma", (B, S))
    delta = _f32(delta, "delta", (B, S))
    rgb = _f32(rgb, "rgb", (B, S, 3))
    if t_mid is not None:
        t_mid = _f32(t_mid, "t_mid", (B, S))
    out_rgb = th.empty((B, 3), device=sigma.device, dtype=th.float32)
    out_w = th.empty((B, S), device=sigma.device, dtype=th.float32) if want_w else None
    out_o = th.empty((B,), device=sigma.device, dtype=th.float32) if want_opacity else None
    out_d = th.empty((B,), device=sigma.device, dtype=th.float32) if want_depth else None
    with th.cuda.device(sigma.device):
        check(lib().nerfb200_composite_fwd(_ptr(sigma), _ptr(delta), _ptr(rgb), _ptr(t_mid), B, S,
                                           flavour, _ptr(out_rgb), _ptr(out_w), _ptr(out_o),
                                           _ptr(out_d), _stream()), "composite_fwd")
    return out_rgb, out_w, out_o, out_d


def composite_bwd(sigma, delta, rgb, g_rgb, g_w=None, t_mid=None, g_opacity=None, g_depth=None,
                  flavour=_lib.COMPOSITE_BARF):
    B, S = sigma.shape
    sigma = _f32(sigma, "sigma", (B, S))
    delta = _f32(delta, "delta", (B, S))
    rgb = _f32(rgb, "rgb", (B, S, 3))
    g_rgb = _f32(g_rgb, "g_rgb", (B, 3))
    g_w = None if g_w is None else _f32(g_w, "g_w", (B, S))
    t_mid = None if t_mid is None else _f32(t_mid, "t_mid", (B, S))
    g_opacity = None if g_opacity is None else _f32(g_opacity.reshape(-1), "g_opacity", (B,))
    g_depth = None if g_depth is None else _f32(g_depth.reshape(-1), "g_depth", (B,))
    d_sigma = th.empty((B, S), device=sigma.device, dtype=th.float32)
    d_rgb = th.empty((B, S, 3), device=sigma.device, dtype=th.float32)
    with th.cuda.device(sigma.device):
        check(lib().nerfb200_composite_bwd(_ptr(sigma), _ptr(delta), _ptr(rgb), _ptr(t_mid),
                                           _ptr(g_rgb), _ptr(g_w), _ptr(g_opacity), _ptr(g_depth),
                                           B, S, flavour, _ptr(d_sigma), _ptr(d_rgb), _stream()),
              "composite_bwd")
    return d_sigma, d_rgb


class _Composite(th.autograd.Function):
    """rgb, weights = render(sigma, rgb_samples, delta) — barf/model_interpolation.py:316-353."""

    @staticmethod
    def forward(ctx, sigma, rgb, delta):
        out_rgb, out_w, _, _ = composite_fwd(sigma, delta, rgb)
        ctx.save_for_backward(sigma.detach(), rgb.detach(), delta.detach())
        return out_rgb, out_w

    @staticmethod
    def backward(ctx, g_rgb, g_w):
        sigma, rgb, delta = ctx.saved_tensors
        g_rgb = th.zeros_like(rgb[:, 0, :]) if g_rgb is None else g_rgb
        d_sigma, d_rgb = composite_bwd(sigma, delta, rgb, g_rgb.contiguous(),
                                       None if g_w is None else g_w.contiguous())
        return d_sigma, d_rgb, None


def render_rays(densities, colors, distances):
    """Drop-in for NerfInterpolation._render_rays."""
    return _Composite.apply(densities, colors, distances)


# ---------------------------------------------------------------------------------------------
# a11 deterministic resampling
# ---------------------------------------------------------------------------------------------
def resample_alloc(t_coarse, weights, delta_coarse, n_samples: int, near: float, far: float,
                   fallback_u: Optional[th.Tensor] = None, want_counts: bool = False):
    """t_start, t_end (B,Sf) [, counts (B,Sc) int32, fail_flag int32[1]] —
    barf/model_interpolation.py:193-277, including its whole-batch equidistant fallback,
    applied on the device without a host sync."""
    B, Sc = t_coarse.shape
    t_coarse = _f32(t_coarse.detach(), "t_coarse", (B, Sc))
    weights = _f32(weights.detach(), "weights", (B, Sc))
    delta_coarse = _f32(delta_coarse.detach(), "delta_coarse", (B, Sc))
    dev = t_coarse.device
    t_start = th.empty((B, n_samples), device=dev, dtype=th.float32)
    t_end = th.empty_like(t_start)
    counts = th.empty((B, Sc), device=dev, dtype=th.int32) if want_counts else None
    flag = th.zeros((1,), device=dev, dtype=th.int32)
    if fallback_u is not None:
        fallback_u = _f32(fallback_u.reshape(-1), "fallback_u", (B,))
    with th.cuda.device(dev):
        check(lib().nerfb200_resample_alloc(_ptr(t_coarse), _ptr(weights), _ptr(delta_coarse), B, Sc,
                                            n_samples, float(far), _ptr(t_start), _ptr(t_end),
                                            _ptr(counts), _ptr(flag), _stream()), "resample_alloc")
        check(lib().nerfb200_resample_fallback(_ptr(flag), float(near), float(far), B, n_samples,
                                               _ptr(fallback_u), _ptr(t_start), _ptr(t_end),
                                               _stream()), "resample_fallback")
    if want_counts:
        return t_start, t_end, counts, flag
    return t_start, t_end


# ---------------------------------------------------------------------------------------------
# a12 inverse-CDF resampling
# ---------------------------------------------------------------------------------------------
def resample_icdf(edges, cdf, n_out: int, u_ray: Optional[th.Tensor] = None, want_idx: bool = False):
    B, E = edges.shape
    edges = _f32(edges.detach(), "edges", (B, E))
    cdf = _f32(cdf.detach(), "cdf", (B, E))
    if u_ray is not None:
        u_ray = _f32(u_ray.reshape(-1), "u_ray", (B,))
    out = th.empty((B, n_out + 1), device=edges.device, dtype=th.float32)
    idx = th.empty((B, n_out), device=edges.device, dtype=th.int32) if want_idx else None
    with th.cuda.device(edges.device):
        check(lib().nerfb200_resample_icdf(_ptr(edges), _ptr(cdf), _ptr(u_ray), B, E - 1, n_out,
                                           _ptr(out), _ptr(idx), _stream()), "resample_icdf")
    return (out, idx) if want_idx else out


def lindisp_intervals(s_edges: th.Tensor, near: float, far: float, want_mid: bool = False):
    """nerfacc's "lindisp" map t = 1 / (s / far + (1 - s) / near) of normalised edges (B, S+1) and the
    intervals it defines: returns (t_edges (B,S+1), t_start (B,S), t_end (B,S)[, delta, t_mid]) from ONE
    launch (garf/model_garf.py:210-220 through PropNetEstimator.sampling)."""
    B, E = s_edges.shape
    s_edges = _f32(s_edges.detach(), "s_edges", (B, E))
    dev = s_edges.device
    t = th.empty((B, E), device=dev)
    t0, t1 = th.empty((B, E - 1), device=dev), th.empty((B, E - 1), device=dev)
    delta = th.empty_like(t0) if want_mid else None
    mid = th.empty_like(t0) if want_mid else None
    with th.cuda.device(dev):
        check(lib().nerfb200_lindisp_intervals(_ptr(s_edges), float(near), float(far), B, E, _ptr(t), _ptr(t0), _ptr(t1),
                                               _ptr(delta), _ptr(mid), _stream()), "lindisp_intervals")
    return (t, t0, t1, delta, mid) if want_mid else (t, t0, t1)


class _TransCdf(th.autograd.Function):
    """cdf (B, S+1) = 1 - [exp(-exclusive cumsum(sigma delta)), 0] — the proposal level of
    PropNetEstimator.sampling (render_transmittance_from_density + the cdf of its weights); the gradient
    reaches sigma only (the sample positions carry none in the reference)."""

    @staticmethod
    def forward(ctx, sigma, t_start, t_end):
        B, S = sigma.shape
        sigma = _f32(sigma, "sigma", (B, S))
        t_start = _f32(t_start, "t_start", (B, S))
        t_end = _f32(t_end, "t_end", (B, S))
        cdf = th.empty((B, S + 1), device=sigma.device)
        with th.cuda.device(sigma.device):
            check(lib().nerfb200_trans_cdf_fwd(_ptr(sigma), _ptr(t_start), _ptr(t_end), B, S, None, _ptr(cdf),
                                               _stream()), "trans_cdf_fwd")
        ctx.save_for_backward(sigma.detach(), t_start, t_end)
        return cdf

    @staticmethod
    def backward(ctx, g_cdf):
        sigma, t_start, t_end = ctx.saved_tensors
        B, S = sigma.shape
        d_sigma = th.empty_like(sigma)
        with th.cuda.device(sigma.device):
            check(lib().nerfb200_trans_cdf_bwd(_ptr(sigma), _ptr(t_start), _ptr(t_end), _ptr(g_cdf.contiguous()), None,
                                               B, S, _ptr(d_sigma), _stream()), "trans_cdf_bwd")
        return d_sigma, None, None


def transmittance_cdf(sigma: th.Tensor, t_start: th.Tensor, t_end: th.Tensor) -> th.Tensor:
    return _TransCdf.apply(sigma, t_start, t_end)


def transmittance(sigma: th.Tensor, t_start: th.Tensor, t_end: th.Tensor):
    """(trans (B,S), cdf (B,S+1)) without gradient (the radiance level: extras["trans"])."""
    B, S = sigma.shape
    sigma = _f32(sigma.detach(), "sigma", (B, S))
    trans = th.empty((B, S), device=sigma.device)
    cdf = th.empty((B, S + 1), device=sigma.device)
    with th.cuda.device(sigma.device):
        check(lib().nerfb200_trans_cdf_fwd(_ptr(sigma), _ptr(_f32(t_start, "t_start", (B, S))),
                                           _ptr(_f32(t_end, "t_end", (B, S))), B, S, _ptr(trans), _ptr(cdf), _stream()),
              "trans_cdf_fwd")
    return trans, cdf


class _PropLoss(th.autograd.Function):
    """mean( clip(w - w_outer, 0)^2 / (w + eps) ) — PropNetEstimator.compute_loss (garf/model_garf.py:257):
    the key (proposal) histogram must bound the query (radiance) histogram from above. One launch
    computes the loss and its gradient w.r.t. the key cdf (the only input that carries one)."""

    @staticmethod
    def forward(ctx, t_query, cdf_query, t_key, cdf_key, eps):
        B, Eq = t_query.shape
        Ek = t_key.shape[1]
        t_query = _f32(t_query.detach(), "t_query", (B, Eq))
        cdf_query = _f32(cdf_query.detach(), "cdf_query", (B, Eq))
        t_key = _f32(t_key.detach(), "t_key", (B, Ek))
        cdf_key_c = _f32(cdf_key.detach(), "cdf_key", (B, Ek))
        loss = th.zeros((), device=t_query.device)
        d_key = th.empty((B, Ek), device=t_query.device) if cdf_key.requires_grad else None
        with th.cuda.device(t_query.device):
            check(lib().nerfb200_prop_loss(_ptr(t_query), _ptr(cdf_query), _ptr(t_key), _ptr(cdf_key_c), B, Eq - 1, Ek - 1,
                                           float(eps), 1.0 / float(B * (Eq - 1)), _ptr(loss), _ptr(d_key), _stream()),
                  "prop_loss")
        ctx.d_key = d_key
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_key
        ctx.d_key = None
        return None, None, None, (d * g if d is not None else None), None


def proposal_loss(t_query, cdf_query, t_key, cdf_key, eps: float = 1e-7) -> th.Tensor:
    return _PropLoss.apply(t_query, cdf_query, t_key, cdf_key, eps)


# ---------------------------------------------------------------------------------------------
# a13 camera extrinsics
# ---------------------------------------------------------------------------------------------
def so3_to_SO3(so3: th.Tensor) -> th.Tensor:
    flat = _f32(so3.reshape(-1, 3), "so3")
    out = th.empty((flat.shape[0], 3, 3), device=flat.device, dtype=th.float32)
    with th.cuda.device(flat.device):
        check(lib().nerfb200_so3_to_SO3(_ptr(flat), flat.shape[0], _ptr(out), _stream()), "so3_to_SO3")
    return out


class _Pose(th.autograd.Function):
    """(o', d', R, t) = CameraExtrinsics.forward(i, o, d) — barf/model_camera_extrinsics.py:77-85.
    Gradients flow to rotation / translation (and to o, d for completeness); R and t outputs
    are returned detached, as nothing in the reference differentiates through them."""

    @staticmethod
    def forward(ctx, rotation, translation, img_idx, o, d, grad_sink=None):
        ctx.grad_sink = grad_sink
        B = o.shape[0]
        rotation = _f32(rotation, "rotation")
        translation = _f32(translation, "translation")
        idx = _i32(img_idx, "img_idx")
        o = _f32(o, "o", (B, 3))
        d = _f32(d, "d", (B, 3))
        out_o = th.empty_like(o)
        out_d = th.empty_like(d)
        out_R = th.empty((B, 3, 3), device=o.device, dtype=th.float32)
        out_t = th.empty_like(o)
        with th.cuda.device(o.device):
            check(lib().nerfb200_pose_fwd(_ptr(rotation), _ptr(translation), _ptr(idx), _ptr(o), _ptr(d),
                                          B, rotation.shape[0], _ptr(out_o), _ptr(out_d), _ptr(out_R),
                                          _ptr(out_t), _stream()), "pose_fwd")
        ctx.save_for_backward(rotation.detach(), idx, d.detach(), out_R)
        ctx.mark_non_differentiable(out_R, out_t)
        return out_o, out_d, out_R, out_t

    @staticmethod
    def backward(ctx, g_o, g_d, _gR, _gt):
        rotation, idx, d, R = ctx.saved_tensors
        B = d.shape[0]
        g_o = th.zeros_like(d) if g_o is None else g_o.contiguous()
        g_d = th.zeros_like(d) if g_d is None else g_d.contiguous()
        sink = getattr(ctx, "grad_sink", None)
        if sink is not None:      # engine mode: accumulate straight into the flat gradient buffer
            d_rot, d_tr = sink
        else:
            d_rot = th.zeros_like(rotation)
            d_tr = th.zeros_like(rotation)
        with th.cuda.device(d.device):
            check(lib().nerfb200_pose_bwd(_ptr(rotation), _ptr(idx), _ptr(d), _ptr(g_o), _ptr(g_d), B,
                                          rotation.shape[0], _ptr(d_rot), _ptr(d_tr), _stream()),
                  "pose_bwd")
        g_in_d = th.matmul(R.transpose(1, 2), g_d.unsqueeze(-1)).squeeze(-1) if ctx.needs_input_grad[4] else None
        if sink is not None:
            d_rot = d_tr = None
        return d_rot, d_tr, None, (g_o if ctx.needs_input_grad[3] else None), g_in_d, None


def pose_forward(rotation, translation, img_idx, o, d, grad_sink=None):
    """grad_sink: optional (d_rotation, d_translation) views the backward accumulates into."""
    return _Pose.apply(rotation, translation, img_idx, o, d, grad_sink)


# ---------------------------------------------------------------------------------------------
# a6 / a9 learnable activations (GARF Gaussian, SARF, Gabor)
# ---------------------------------------------------------------------------------------------
class _Activation(th.autograd.Function):
    """y = act(x; p0[, p1]) over (N, F) activations with per-feature parameters; the backward
    kernel returns dx and the row-reduced parameter gradients (what autograd's sum-to-shape does
    in the reference, barf/gaussian.py:21-34, gaborf/gabor.py:19-29)."""

    @staticmethod
    def forward(ctx, kind, x, p0, p1):
        shape = x.shape
        F = shape[-1]
        x2 = _f32(x.reshape(-1, F), "x")
        p0c = _f32(p0.reshape(-1), "p0", (F,))
        p1c = None if p1 is None else _f32(p1.reshape(-1), "p1", (F,))
        y = th.empty_like(x2)
        with th.cuda.device(x2.device):
            check(lib().nerfb200_act_fwd(kind, _ptr(x2), _ptr(p0c), _ptr(p1c), x2.shape[0], F, _ptr(y), 0,
                                         _stream()), "act_fwd")
        ctx.kind = kind
        ctx.shape = shape
        ctx.save_for_backward(x2, p0c, p1c)
        return y.view(shape)

    @staticmethod
    def backward(ctx, g):
        x2, p0c, p1c = ctx.saved_tensors
        F = x2.shape[1]
        g2 = _f32(g.reshape(-1, F), "g")
        dx = th.empty_like(x2)
        dp0 = th.zeros_like(p0c)
        dp1 = None if p1c is None else th.zeros_like(p1c)
        with th.cuda.device(x2.device):
            check(lib().nerfb200_act_bwd(ctx.kind, _ptr(x2), _ptr(p0c), _ptr(p1c), _ptr(g2), x2.shape[0], F,
                                         _ptr(dx), _ptr(dp0), _ptr(dp1), None, 0, _stream()), "act_bwd")
        return None, dx.view(ctx.shape), dp0, dp1


def activation(kind: int, x: th.Tensor, p0: th.Tensor, p1: Optional[th.Tensor] = None) -> th.Tensor:
    return _Activation.apply(kind, x, p0, p1)


class _tf32_matmul:
    """GEMMs inside run on the tensor cores with TF32 operands and fp32 accumulation — the precision
    the reference trains at (th.set_float32_matmul_precision("high"), barf/run_barf.py:101;
    garf/main.py:93 goes further down to fp16 autocast)."""

    def __init__(self, enabled: bool):
        self.enabled = enabled

    def __enter__(self):
        self.prev = th.backends.cuda.matmul.allow_tf32
        if self.enabled:
            th.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        th.backends.cuda.matmul.allow_tf32 = self.prev


class _LinearActivation(th.autograd.Function):
    """y = act(x W^T + b; p0[, p1]) (kind < 0: no activation).  The GEMMs are library calls
    (cuBLAS through torch); the activation, its input gradient, the parameter gradients AND the
    bias gradient come from one pass of the activation kernels.  precision:
      "bf16"  bf16 operands, fp32 accumulation and fp32 pre-activations (the arithmetic of the fused
              NeRF kernels; the reference's garf/main.py:93 trains in fp16 autocast). The activation
              kernels then emit y / dz directly as bf16, the operand type of the next GEMM.
      "tf32"  fp32 tensors, TF32 tensor-core GEMMs (barf/run_barf.py:101).
      "fp32"  fp32 SIMT GEMMs (tight parity tests).
    Round-1 shape of the GARF networks (their fused tile kernel is round 2)."""

    @staticmethod
    def forward(ctx, kind, precision, last, x, weight, bias, p0, p1):
        bf16 = precision == "bf16"
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_cuda:
            raise RuntimeError("x: expected a CUDA tensor (nerfb200 has no CPU fallback)")
        if bf16:
            xg = x2 if x2.dtype == th.bfloat16 else x2.to(th.bfloat16)
            wg = weight.to(th.bfloat16)
            z = th.addmm(bias, xg, wg.t(), out_dtype=th.float32)
        else:
            xg, wg = _f32(x2, "x"), weight
            with _tf32_matmul(precision == "tf32"):
                z = th.addmm(bias, xg, weight.t())
        F = z.shape[1]
        out_bf16 = bf16 and not last         # the last layer's output leaves the network in fp32
        if kind >= 0:
            p0c = _f32(p0.reshape(-1), "p0", (F,))
            p1c = None if p1 is None else _f32(p1.reshape(-1), "p1", (F,))
            y = th.empty(z.shape, device=z.device, dtype=th.bfloat16 if out_bf16 else th.float32)
            with th.cuda.device(z.device):
                check(lib().nerfb200_act_fwd(kind, _ptr(z), _ptr(p0c), _ptr(p1c), z.shape[0], F, _ptr(y),
                                             int(out_bf16), _stream()), "act_fwd")
            ctx.save_for_backward(xg, wg, z, p0c, p1c)
        else:
            y = z.to(th.bfloat16) if out_bf16 else z
            ctx.save_for_backward(xg, wg, None, None, None)
        ctx.kind, ctx.precision, ctx.in_shape, ctx.in_dtype = kind, precision, x.shape, x.dtype
        return y.view(*x.shape[:-1], F)

    @staticmethod
    def backward(ctx, g):
        xg, wg, z, p0c, p1c = ctx.saved_tensors
        bf16 = ctx.precision == "bf16"
        F = wg.shape[0]
        g2 = _f32(g.reshape(-1, F).float(), "g")
        dp0 = dp1 = None
        if ctx.kind >= 0:
            dz = th.empty(z.shape, device=z.device, dtype=th.bfloat16 if bf16 else th.float32)
            dp0 = th.zeros_like(p0c)
            dp1 = None if p1c is None else th.zeros_like(p1c)
            db = th.zeros(F, device=z.device, dtype=th.float32)
            with th.cuda.device(z.device):
                check(lib().nerfb200_act_bwd(ctx.kind, _ptr(z), _ptr(p0c), _ptr(p1c), _ptr(g2), z.shape[0], F,
                                             _ptr(dz), _ptr(dp0), _ptr(dp1), _ptr(db), int(bf16), _stream()), "act_bwd")
        else:
            db = g2.sum(0) if ctx.needs_input_grad[5] else None
            dz = g2.to(th.bfloat16) if bf16 else g2
        dx = dw = None
        if bf16:
            if ctx.needs_input_grad[3]:
                dx = th.mm(dz, wg, out_dtype=th.float32).view(ctx.in_shape)
            if ctx.needs_input_grad[4]:
                dw = th.mm(dz.t(), xg, out_dtype=th.float32)
        else:
            with _tf32_matmul(ctx.precision == "tf32"):
                dx = (dz @ wg).view(ctx.in_shape) if ctx.needs_input_grad[3] else None
                dw = dz.t() @ xg if ctx.needs_input_grad[4] else None
        return None, None, None, dx, dw, db, dp0, dp1


def linear_activation(x, weight, bias, kind: int = -1, p0=None, p1=None, precision: str = "tf32", last: bool = False):
    return _LinearActivation.apply(kind, precision, last, x, weight, bias, p0, p1)


# ---------------------------------------------------------------------------------------------
# pose alignment (Kabsch with outlier rejection) and pose error — SURVEY.md §8f rank 4
# ---------------------------------------------------------------------------------------------
def kabsch(point_cloud_from: th.Tensor, point_cloud_to: th.Tensor, remove_outliers: bool = True,
           want_error: bool = False):
    """R (3,3), t (1,3), c (1,) with  to ~= c * R from + t  — the convention of
    CameraCalibrationModel.kabsch_algorithm (barf/model_camera_calibration.py:69-156); one launch,
    no host synchronisation. With want_error also the mean alignment error (compute_pose_error)."""
    if point_cloud_from.shape != point_cloud_to.shape or point_cloud_from.dim() != 2 or point_cloud_from.shape[1] != 3:
        raise ValueError("point_cloud_from and point_cloud_to must both be of shape (N, 3)")
    a = _f32(point_cloud_from.detach(), "point_cloud_from")
    b = _f32(point_cloud_to.detach(), "point_cloud_to")
    dev = a.device
    R = th.empty((3, 3), device=dev)
    t = th.empty((1, 3), device=dev)
    c = th.empty((1,), device=dev)
    err = th.empty((1,), device=dev) if want_error else None
    with th.cuda.device(dev):
        check(lib().nerfb200_kabsch(_ptr(a), _ptr(b), a.shape[0], int(bool(remove_outliers)), _ptr(R), _ptr(t),
                                    _ptr(c), _ptr(err), _stream()), "kabsch")
    return (R, t, c, err[0]) if want_error else (R, t, c)
