"""Host side of the fused GARF field kernels (csrc/garf_fwd.cu, garf_bwd.cu, mlp_wgrad.cu): compiles a
RadianceNetwork / ProposalNetwork into tile programs (garf_program.py), keeps the packed bf16 weight
images and fp32 values fresh, and exposes the launch as an autograd Function over rays or samples.
There is no library GEMM and no CPU path behind it."""
import ctypes as C
from typing import Optional

import torch as th

from . import _lib
from ._lib import NbGaussLayer, NbPackBias, NbPackChunk, NbWgradItem, check, lib
from .fused_mlp import FlatParams, make_inputs
from .mlp_program import schedule_wgrad, to_device_array


def _ptr(t):
    return None if t is None else t.data_ptr()


class FusedGarfField:
    """Compiled tile programs + device buffers of one GARF network."""

    def __init__(self, compile_fn, flat: FlatParams, own_params):
        self.compile_fn = compile_fn          # callable(FlatParams) -> garf_program.CompiledGarf
        self.flat = flat
        self.own_params = list(own_params)
        self.compiled = None
        self.device = None
        self._packed_sig = None
        self.timers = None                    # optional per-kernel CUDA events (bench.py)

    def prepare(self, device):
        flat = self.flat.ensure(device)
        if self.compiled is None or self.device != device:
            cg = self.compile_fn(self.flat)
            self.compiled = cg
            self.device = device
            self.wpack = th.zeros(max(cg.wpack_units * 1024, 1024), device=device, dtype=th.uint8)
            self.floats_fwd = th.zeros(max(cg.fwd.n_floats, 1), device=device, dtype=th.float32)
            self.floats_bwd = th.zeros(max(cg.bwd.n_floats, 1), device=device, dtype=th.float32)
            self.chunks_dev = to_device_array(cg.pack_chunks, NbPackChunk, device)
            self.fdesc_fwd = to_device_array(cg.fwd_floats, NbPackBias, device)
            self.fdesc_bwd = to_device_array(cg.bwd_floats, NbPackBias, device)
            self._wgrad_items = {}
            self._packed_sig = None
            gl = [NbGaussLayer(w_off=l.lin.w_off, b_off=l.lin.b_off, g_off=l.g_off, in_f=l.lin.in_f, out_f=l.lin.out_f)
                  for l in cg.gauss_layers]
            self.gauss_dev = to_device_array(gl, NbGaussLayer, device) if gl else None
            self.n_gauss_layers, self.n_gauss_features = len(gl), sum(l.out_f for l in gl)
        sig = self.flat.signature()
        if sig != self._packed_sig:
            cg = self.compiled
            stream = th.cuda.current_stream().cuda_stream
            with th.cuda.device(device):
                check(lib().nerfb200_mlp_pack(_ptr(flat), _ptr(self.chunks_dev), len(cg.pack_chunks), _ptr(self.wpack),
                                              _ptr(self.fdesc_fwd), len(cg.fwd_floats), _ptr(self.floats_fwd), stream),
                      "mlp_pack")
                check(lib().nerfb200_mlp_pack(_ptr(flat), None, 0, None, _ptr(self.fdesc_bwd), len(cg.bwd_floats),
                                              _ptr(self.floats_bwd), stream), "mlp_pack")
            self._packed_sig = sig
        return self.compiled

    def _timed(self, name, fn):
        if self.timers is None:
            return fn()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.timers.setdefault(name, []).append((e0, e1))

    def forward(self, inputs, n: int, training: bool):
        dev, cg = self.device, self.compiled
        sigma = th.empty((n,), device=dev, dtype=th.float32)
        rgb = th.empty((n, 3), device=dev, dtype=th.float32) if cg.has_rgb else None
        y_stash = z_stash = None
        if training:
            yb, zb = C.c_longlong(), C.c_longlong()
            check(lib().nerfb200_garf_workspace_bytes(C.byref(cg.fwd), n, C.byref(yb), C.byref(zb)), "garf_workspace_bytes")
            y_stash = th.empty(max(yb.value, 1), device=dev, dtype=th.uint8)
            z_stash = th.empty(max(zb.value, 1), device=dev, dtype=th.uint8)
        with th.cuda.device(dev):
            self._timed("garf_fwd_train" if training else "garf_fwd", lambda: check(lib().nerfb200_garf_fwd(
                C.byref(cg.fwd), _ptr(self.wpack), _ptr(self.floats_fwd), _ptr(self.flat.flat), C.byref(inputs),
                _ptr(sigma), _ptr(rgb), _ptr(y_stash), _ptr(z_stash), th.cuda.current_stream().cuda_stream), "garf_fwd"))
        return sigma, rgb, y_stash, z_stash

    def _items(self, n_tiles: int):
        if n_tiles not in self._wgrad_items:
            n_sm = th.cuda.get_device_properties(self.device).multi_processor_count
            # items of ~20 MB of streamed slabs each (mlp_program.schedule_wgrad: measured), at least 3 per SM
            slabs = sum(u.n_dy_slabs + u.n_x_slabs + u.n_z_slabs for u in self.compiled.units) * n_tiles
            per_worker = min(max(round(slabs * _lib.NB_SLAB_BYTES / (n_sm * 20e6)), 3), 8)
            items = schedule_wgrad(self.compiled.units, n_tiles, n_sm, per_worker)
            self._wgrad_items[n_tiles] = (to_device_array(items, NbWgradItem, self.device), len(items))
        return self._wgrad_items[n_tiles]

    def backward(self, inputs, n: int, sigma, rgb, g_sigma, g_rgb, y_stash, z_stash, want_input_grads: bool, n_rays: int):
        dev, cg = self.device, self.compiled
        n_tiles = (n + _lib.NB_TILE_ROWS - 1) // _lib.NB_TILE_ROWS
        dy_stash = th.empty(n_tiles * cg.bwd.y_slabs_per_tile * _lib.NB_SLAB_BYTES, device=dev, dtype=th.uint8)
        flat_grad = self.flat.grad_sink if self.flat.grad_sink is not None else \
            th.zeros(self.flat.numel, device=dev, dtype=th.float32)
        d_a = d_b = None
        samples_mode = bool(inputs.pos)
        if want_input_grads:
            rows = n if samples_mode else n_rays
            d_a = th.zeros((rows, 3), device=dev, dtype=th.float32)
            d_b = th.zeros((rows, 3), device=dev, dtype=th.float32)
        stream = th.cuda.current_stream().cuda_stream
        with th.cuda.device(dev):
            self._timed("garf_bwd", lambda: check(lib().nerfb200_garf_bwd(
                C.byref(cg.bwd), _ptr(self.wpack), _ptr(self.floats_bwd), _ptr(self.flat.flat), C.byref(inputs),
                _ptr(sigma), _ptr(rgb), _ptr(g_sigma), _ptr(g_rgb), _ptr(z_stash), _ptr(dy_stash),
                int(bool(want_input_grads)),
                None if samples_mode else _ptr(d_a), None if samples_mode else _ptr(d_b),
                _ptr(d_a) if samples_mode else None, _ptr(d_b) if samples_mode else None, stream), "garf_bwd"))
            items_dev, n_items = self._items(n_tiles)
            # Gaussian widths: d s = (W . dW + b db) s / (s^2 + 1e-6) from THIS call's dW, db — whatever the
            # gradient buffer already holds is subtracted before the weight-gradient kernel and added back after
            width = lambda sign: check(lib().nerfb200_gauss_width_grad(
                _ptr(self.gauss_dev), self.n_gauss_layers, self.n_gauss_features, _ptr(self.flat.flat), _ptr(flat_grad),
                sign, stream), "gauss_width_grad")
            if self.flat.grad_sink is not None:
                width(-1.0)
            self._timed("garf_wgrad", lambda: check(lib().nerfb200_mlp_wgrad(
                _ptr(items_dev), n_items, _ptr(y_stash), cg.fwd.y_slabs_per_tile, _ptr(dy_stash),
                cg.bwd.y_slabs_per_tile, _ptr(z_stash), cg.fwd.z_slabs_per_tile, _ptr(self.flat.flat),
                _ptr(flat_grad), stream), "mlp_wgrad"))
            width(1.0)
        return flat_grad, d_a, d_b


class _GarfFunction(th.autograd.Function):
    """sigma[, rgb] = network(inputs; params); mode "samples" (pos / dir per sample, the reference's
    forward signature) or "rays" (o, d per ray + sample bins: x = o + (t0 + t1) / 2 d in registers)."""

    @staticmethod
    def forward(ctx, net, mode, S, training, a, b, t_start, t_end, *params):
        if not a.is_cuda:
            raise RuntimeError("the fused GARF field runs on CUDA only (nerfb200 has no CPU fallback)")
        field = net.fused_field()
        field.prepare(a.device)
        if mode == "samples":
            n = a.shape[0]
            inputs = make_inputs(n, 1, 0, pos=a, dir=b if b is not None else a, pixel_width_per_sample=True)
        else:
            n = a.shape[0] * S
            inputs = make_inputs(n, S, 1, ray_o=a, ray_d=b, t_start=t_start, t_end=t_end)
        sigma, rgb, y_stash, z_stash = field.forward(inputs, n, training)
        ctx.net, ctx.mode, ctx.S, ctx.n = net, mode, S, n
        ctx.y_stash, ctx.z_stash = y_stash, z_stash
        ctx.save_for_backward(a, b, t_start, t_end, sigma, rgb)
        if rgb is None:
            return sigma
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb=None):
        a, b, t_start, t_end, sigma, rgb = ctx.saved_tensors
        field = ctx.net.fused_field()
        if ctx.y_stash is None:
            raise RuntimeError("fused GARF field: backward requested but the forward ran without gradient tracking")
        n = ctx.n
        if ctx.mode == "samples":
            inputs = make_inputs(n, 1, 0, pos=a, dir=b if b is not None else a, pixel_width_per_sample=True)
            n_rays = n
        else:
            inputs = make_inputs(n, ctx.S, 1, ray_o=a, ray_d=b, t_start=t_start, t_end=t_end)
            n_rays = a.shape[0]
        want = ctx.needs_input_grad[4] or (b is not None and ctx.needs_input_grad[5])
        g_sigma = None if g_sigma is None else g_sigma.contiguous().float()
        g_rgb = None if g_rgb is None else g_rgb.contiguous().float()
        flat_grad, d_a, d_b = field.backward(inputs, n, sigma, rgb, g_sigma, g_rgb, ctx.y_stash, ctx.z_stash, want, n_rays)
        ctx.y_stash = ctx.z_stash = None
        field.flat.last_grad = flat_grad
        if field.flat.grad_sink is not None:
            param_grads = tuple(None for _ in ctx.needs_input_grad[8:])
        else:
            offs = [field.flat.offset_of(p) for p in field.own_params]
            param_grads = tuple(flat_grad[o:o + p.numel()].view(p.shape) if need else None
                                for p, o, need in zip(field.own_params, offs, ctx.needs_input_grad[8:]))
        return (None, None, None, None, d_a if ctx.needs_input_grad[4] else None,
                d_b if (b is not None and ctx.needs_input_grad[5]) else None, None, None) + param_grads


def _tracking(tensors) -> bool:
    return th.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def garf_samples(net, pos: th.Tensor, dir: Optional[th.Tensor]):
    """(sigma[, rgb]) for per-sample positions / directions."""
    pos = pos.contiguous().float()
    dir = None if dir is None else dir.contiguous().float()
    params = net.fused_field().own_params
    return _GarfFunction.apply(net, "samples", 1, _tracking([pos, dir, *params]), pos, dir, None, None, *params)


def garf_rays(net, ray_o, ray_d, t_start, t_end):
    """(sigma (B,S)[, rgb (B,S,3)]) for rays sampled at the bins (t_start, t_end), mid-point rule."""
    B, S = t_start.shape
    params = net.fused_field().own_params
    out = _GarfFunction.apply(net, "rays", S, _tracking([ray_o, ray_d, *params]), ray_o.contiguous().float(),
                              ray_d.contiguous().float(), t_start.contiguous(), t_end.contiguous(), *params)
    if isinstance(out, tuple):
        return out[0].view(B, S), out[1].view(B, S, 3)
    return out.view(B, S)
