"""Backward of the fused field Function (see field_function.py)."""
import torch as th

from .fused_mlp import make_inputs


def field_backward(ctx, g_sigma, g_rgb):
    a, b, t_start, t_end, pixel_width, sigma, rgb = ctx.saved_tensors
    model = ctx.model
    field = model.fused_field()
    if ctx.stash is None:
        raise RuntimeError("fused field: backward requested but the forward ran without gradient tracking")
    n = ctx.n
    if ctx.mode == "samples":
        inputs = make_inputs(n, 1, 0, pos=a, dir=b, t_start=t_start, t_end=t_end, pixel_width=pixel_width,
                             pixel_width_per_sample=True)
        n_rays = n
    else:
        inputs = make_inputs(n, ctx.S, ctx.t_mode, ray_o=a, ray_d=b, t_start=t_start, t_end=t_end,
                             pixel_width=pixel_width, pixel_width_per_sample=False)
        n_rays = a.shape[0]
    want_inputs = ctx.needs_input_grad[5] or ctx.needs_input_grad[6]
    g_sigma = None if g_sigma is None else g_sigma.contiguous().float()
    g_rgb = None if g_rgb is None else g_rgb.contiguous().float()
    flat_grad, d_a, d_b = field.backward(inputs, n, sigma, rgb, g_sigma, g_rgb, ctx.stash, ctx.masks,
                                         want_inputs, n_rays)
    ctx.stash = ctx.masks = None
    field.flat.last_grad = flat_grad
    if field.flat.grad_sink is not None:
        param_grads = tuple(None for _ in ctx.needs_input_grad[10:])     # engine mode
    else:
        offs = [field.flat.offset_of(p) for p in field.own_params]
        param_grads = tuple(flat_grad[o:o + p.numel()].view(p.shape) if need else None
                            for p, o, need in zip(field.own_params, offs, ctx.needs_input_grad[10:]))
    return (None, None, None, None, None,
            d_a if ctx.needs_input_grad[5] else None,
            d_b if ctx.needs_input_grad[6] else None,
            None, None, None) + param_grads
