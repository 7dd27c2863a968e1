"""autograd glue of the fused field kernels: (rays | samples, parameters) -> (sigma, rgb)."""
from typing import Optional

import torch as th

from .fused_mlp import make_inputs


def _prep(t: Optional[th.Tensor], n: int, width: int, device):
    if t is None:
        return None
    if not isinstance(t, th.Tensor):
        t = th.full((n, width), float(t), device=device, dtype=th.float32)
    t = t.to(device=device, dtype=th.float32)
    if t.numel() == n * width:
        return t.reshape(n, width).contiguous()
    return t.reshape(-1, width).expand(n, width).contiguous()


class _FieldFunction(th.autograd.Function):
    """sigma, rgb = field(inputs; params).  `mode` = "samples" (pos/dir per sample, the
    NerfModel.forward signature) or "rays" (o, d per ray + t bins per sample)."""

    @staticmethod
    def forward(ctx, model, mode, S, t_mode, training, a, b, t_start, t_end, pixel_width, *params):
        dev = a.device
        if not a.is_cuda:
            raise RuntimeError("the fused field runs on CUDA only (nerfb200 has no CPU fallback)")
        field = model.fused_field()
        field.prepare(dev)
        if mode == "samples":
            n = a.shape[0]
            inputs = make_inputs(n, 1, 0, pos=a, dir=b, t_start=t_start, t_end=t_end,
                                 pixel_width=pixel_width, pixel_width_per_sample=True)
        else:
            n = a.shape[0] * S
            inputs = make_inputs(n, S, t_mode, ray_o=a, ray_d=b, t_start=t_start, t_end=t_end,
                                 pixel_width=pixel_width, pixel_width_per_sample=False)
        sigma, rgb, stash, masks = field.forward(inputs, n, training, (a, b, t_start, t_end, pixel_width))
        ctx.model, ctx.mode, ctx.S, ctx.t_mode, ctx.n = model, mode, S, t_mode, n
        ctx.stash, ctx.masks = stash, masks
        ctx.save_for_backward(a, b, t_start, t_end, pixel_width, sigma, rgb)
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        from .field_backward import field_backward
        return field_backward(ctx, g_sigma, g_rgb)


def _tracking(tensors) -> bool:
    """Does autograd record this call?  (Decided outside Function.forward, where grad mode is
    always off; under no_grad the kernel then skips the activation stash.)"""
    return th.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def field_samples(model, pos, dir, pixel_width=None, t_start=None, t_end=None):
    n = pos.shape[0]
    dev = pos.device
    pos = _prep(pos, n, 3, dev)
    dir = _prep(dir, n, 3, dev)
    pw = _prep(pixel_width, n, 1, dev)
    t0 = _prep(t_start, n, 1, dev)
    t1 = _prep(t_end, n, 1, dev)
    params = model.fused_field().own_params
    return _FieldFunction.apply(model, "samples", 1, 0, _tracking([pos, dir, *params]), pos, dir, t0, t1, pw, *params)


def field_rays(model, ray_o, ray_d, t_start, t_end, pixel_width, integration_strategy: str):
    """sigma (B,S), rgb (B,S,3) for rays sampled at the bins (t_start, t_end)."""
    B, S = t_start.shape
    dev = ray_o.device
    t_mode = {"left": 0, "middle": 1}[integration_strategy]
    pw = None if pixel_width is None else _prep(pixel_width, B, 1, dev)
    params = model.fused_field().own_params
    sigma, rgb = _FieldFunction.apply(model, "rays", S, t_mode, _tracking([ray_o, ray_d, *params]),
                                      ray_o.contiguous().float(), ray_d.contiguous().float(),
                                      t_start.contiguous(), t_end.contiguous(), pw, *params)
    return sigma.view(B, S), rgb.view(B, S, 3)
