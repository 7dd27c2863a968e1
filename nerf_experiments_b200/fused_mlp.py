"""Host side of the fused field (positional encoding + radiance MLP) kernels: keeps the fp32
master parameters in one flat buffer (the nn.Parameters become views of it, so state-dict keys
and optimizers are unaffected), refreshes the packed bf16 weight images when the parameters
change, and exposes the launch as an autograd Function."""
import ctypes as C
from typing import List, Optional

import torch as th
import torch.nn as nn

from . import _lib
from ._lib import NbMlpInputs, NbPackBias, NbPackChunk, check, lib
from ._lib import NbWgradItem
from .mlp_program import (SLAB_PE_DIR, SLAB_PE_POS, CompiledMlp, LayerSpec, Linear, compile_backward,
                          compile_forward, schedule_wgrad, to_device_array)


def _ptr(t):
    return None if t is None else t.data_ptr()


class FlatParams:
    """All parameters of a set of modules as views of one flat fp32 CUDA buffer."""

    def __init__(self, params: List[nn.Parameter]):
        self.params = list(params)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()
        self.numel = off
        self.flat: Optional[th.Tensor] = None
        self.flat_grad: Optional[th.Tensor] = None
        self.version = 0   # bumped by whoever writes `flat` outside torch's version counters
        # engine mode: when set (a flat fp32 tensor of `numel` floats), the backward kernels
        # accumulate parameter gradients straight into it and autograd sees no parameter grads
        self.grad_sink: Optional[th.Tensor] = None
        self.last_grad: Optional[th.Tensor] = None

    def ensure(self, device) -> th.Tensor:
        ok = (self.flat is not None and self.flat.device == device and
              all(p.data.data_ptr() == self.flat.data_ptr() + 4 * o for p, o in zip(self.params, self.offsets)))
        if not ok:
            flat = th.empty(self.numel, device=device, dtype=th.float32)
            for p, o in zip(self.params, self.offsets):
                flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = flat[o:o + p.numel()].view(p.shape)
            self.flat = flat
            self.flat_grad = None
            self.version += 1
        return self.flat

    def signature(self):
        return (self.version, tuple(p._version for p in self.params))

    def offset_of(self, p: nn.Parameter) -> int:
        for q, o in zip(self.params, self.offsets):
            if q is p:
                return o
        raise KeyError("parameter not in the flat buffer")

    def grad_buffer(self) -> th.Tensor:
        if self.flat_grad is None or self.flat_grad.device != self.flat.device:
            self.flat_grad = th.zeros(self.numel, device=self.flat.device, dtype=th.float32)
        return self.flat_grad

    def grad_views(self, flat_grad: th.Tensor):
        return [flat_grad[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]


class FusedField:
    """Compiled tile programs + device buffers of one network (one per NerfModel instance)."""

    def __init__(self, layers_fn, flat: FlatParams, pe_pos, pe_dir, sigma_bias: float = 0.0,
                 own_params=None):
        self.layers_fn = layers_fn        # callable(FlatParams) -> List[LayerSpec]
        self.flat = flat
        # parameters of THIS network (the flat buffer may also hold other networks / the poses)
        self.own_params = list(own_params) if own_params is not None else list(flat.params)
        self.pe_pos = pe_pos
        self.pe_dir = pe_dir
        self.sigma_bias = sigma_bias
        self.compiled: Optional[CompiledMlp] = None
        self.device = None
        self._packed_sig = None
        # optional per-kernel timing (bench.py): name -> list of (start, end) CUDA events
        # recorded on the launching stream; None = off
        self.timers = None
        # Experimental: two 128-sample tiles in flight per CTA (csrc/mlp_fwd.cu, mlp_fwd2) when the
        # program qualifies. Opt-in: measured on B200 it ties the one-tile kernel without the stash
        # (1.00 ms) and loses with it (1.40 vs 1.08 ms) — see DESIGN.md section 4.
        import os
        self.use_two_tile = os.environ.get("NERFB200_TWO_TILE", "0") == "1"

    # -- compilation / packing ------------------------------------------------------------
    def prepare(self, device):
        flat = self.flat.ensure(device)
        if self.compiled is None or self.device != device:
            self.compiled = compile_forward(self.layers_fn(self.flat))
            cm = self.compiled
            self.device = device
            # backward programs: [False] weights only, [True] also gradients w.r.t. the inputs.
            # Their transposed weight images follow the forward images in one packed buffer.
            self.bwd = {}
            all_chunks = list(cm.pack_chunks)
            units = cm.wpack_bytes // 1024
            for want in (False, True):
                cb = compile_backward(cm, want, self._encoders())
                shift = units - cm.wpack_bytes // 1024
                for ch in cb.pack_chunks:
                    ch.dst_off += shift
                for i in range(cb.program.n_ops):
                    op = cb.program.ops[i]
                    for c in range(op.n_chunks):
                        op.w_off[c] += shift
                units += cb.wpack_units
                all_chunks += cb.pack_chunks
                self.bwd[want] = cb
            # the forward weights once more as per-K-step images for the two-tile forward kernel
            self.k16_units = -1
            if cm.two_tile_ok:
                self.k16_units = units
                for ch in cm.pack_chunks_k16:
                    ch.dst_off += units
                all_chunks += cm.pack_chunks_k16
                units += cm.wpack_bytes // 1024
            self.n_pack_chunks = len(all_chunks)
            # zero-initialised: image rows no pack descriptor covers must read as zero weights
            self.wpack = th.zeros(max(units * 1024, 1024), device=device, dtype=th.uint8)
            self.bias = th.zeros(max(cm.bias_floats, 1), device=device, dtype=th.float32)
            self.chunks_dev = to_device_array(all_chunks, NbPackChunk, device)
            self.biases_dev = to_device_array(cm.pack_biases, NbPackBias, device)
            self._wgrad_items = {}
            self._packed_sig = None
        sig = self.flat.signature()
        if sig != self._packed_sig:
            cm = self.compiled
            with th.cuda.device(device):
                check(lib().nerfb200_mlp_pack(_ptr(flat), _ptr(self.chunks_dev), self.n_pack_chunks,
                                              _ptr(self.wpack), _ptr(self.biases_dev), len(cm.pack_biases),
                                              _ptr(self.bias), th.cuda.current_stream().cuda_stream),
                      "mlp_pack")
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

    def macs_per_sample(self):
        """Algorithmic (unpadded) multiply-accumulates per sample: forward, data-gradient (with /
        without the encoding-gradient blocks) and weight-gradient passes."""
        fwd = sum(L.lin.out_f * L.lin.in_f for L in self.compiled.layers)
        enc = sum(L.lin.out_f * s.width for L in self.compiled.layers for s in L.sources if s.kind != "act")
        return {"fwd": fwd, "bwd_inputs": fwd, "bwd": fwd - enc, "wgrad": fwd}

    def _encoders(self):
        out = {}
        from .positional_encodings import IdentityPositionalEncoding
        for kind, enc in (("pos", self.pe_pos), ("dir", self.pe_dir)):
            identity_only = isinstance(enc, IdentityPositionalEncoding)
            has_id = 1 if (identity_only or getattr(enc, "include_identity", False)) else 0
            out[kind] = (0 if identity_only else int(enc.levels), has_id)
        return out

    def pe_cfgs(self):
        cp = self.pe_pos.describe()
        cd = self.pe_dir.describe()
        cp.slab, cp.stash_slab, cp.encode_before_op = SLAB_PE_POS, 0, 0
        cd.slab, cd.stash_slab = self.compiled.dir_slab, 1
        cd.encode_before_op = self.compiled.dir_encode_before_op
        return cp, cd

    # -- launches -------------------------------------------------------------------------
    def forward(self, inputs: NbMlpInputs, n: int, training: bool, keep):
        """Runs the fused forward; returns sigma (n,), rgb (n,3) and the (stash, masks) pair the
        backward kernels need (None, None when not training).  `keep` holds the tensors whose
        pointers sit in `inputs` alive."""
        dev = self.device
        cm = self.compiled
        sigma = th.empty((n,), device=dev, dtype=th.float32)
        rgb = th.empty((n, 3), device=dev, dtype=th.float32)
        stash = masks = None
        if training:
            n_tiles = (n + _lib.NB_TILE_ROWS - 1) // _lib.NB_TILE_ROWS
            stash = th.empty(n_tiles * cm.stash_slabs_per_tile * _lib.NB_SLAB_BYTES, device=dev, dtype=th.uint8)
            masks = th.empty(max(n_tiles * cm.mask_words_per_tile * _lib.NB_TILE_ROWS, 1), device=dev, dtype=th.int32)
        cp, cd = self.pe_cfgs()
        if self.k16_units >= 0 and self.use_two_tile:
            with th.cuda.device(dev):
                self._timed("mlp_fwd_train" if training else "mlp_fwd", lambda: check(lib().nerfb200_mlp_fwd2(
                    C.byref(cm.program), self.wpack.data_ptr() + self.k16_units * 1024, _ptr(self.bias), C.byref(inputs),
                    C.byref(cp), C.byref(cd), _ptr(self.pe_pos.alpha_tensor()), _ptr(self.pe_dir.alpha_tensor()),
                    float(self.sigma_bias), _ptr(sigma), _ptr(rgb), _ptr(stash), _ptr(masks), cm.bias_floats,
                    cm.density_w_off, th.cuda.current_stream().cuda_stream), "mlp_fwd2"))
            return sigma, rgb, stash, masks
        with th.cuda.device(dev):
            self._timed("mlp_fwd_train" if training else "mlp_fwd", lambda: check(lib().nerfb200_mlp_fwd(
                C.byref(cm.program), _ptr(self.wpack), _ptr(self.bias), C.byref(inputs), C.byref(cp), C.byref(cd),
                _ptr(self.pe_pos.alpha_tensor()), _ptr(self.pe_dir.alpha_tensor()), float(self.sigma_bias),
                _ptr(sigma), _ptr(rgb), _ptr(stash), _ptr(masks), cm.bias_floats,
                th.cuda.current_stream().cuda_stream), "mlp_fwd"))
        return sigma, rgb, stash, masks


    def _items(self, n_tiles: int):
        if n_tiles not in self._wgrad_items:
            n_sm = th.cuda.get_device_properties(self.device).multi_processor_count
            items = schedule_wgrad(self.bwd[False].units, n_tiles, n_sm)
            self._wgrad_items[n_tiles] = (to_device_array(items, NbWgradItem, self.device), len(items))
        return self._wgrad_items[n_tiles]

    def backward(self, inputs: NbMlpInputs, n: int, sigma, rgb, g_sigma, g_rgb, stash, masks,
                 want_input_grads: bool, n_rays: int):
        """Backward of `forward`: returns the flat fp32 parameter-gradient buffer and the input
        gradients ((d_o, d_d) per ray in rays mode, (d_pos, d_dir) per sample otherwise)."""
        dev = self.device
        cm = self.compiled
        cb = self.bwd[bool(want_input_grads)]
        n_tiles = (n + _lib.NB_TILE_ROWS - 1) // _lib.NB_TILE_ROWS
        dy_stash = th.empty(n_tiles * cb.dy_slabs_per_tile * _lib.NB_SLAB_BYTES, device=dev, dtype=th.uint8)
        if self.flat.grad_sink is not None:
            flat_grad = self.flat.grad_sink
        else:
            flat_grad = th.zeros(self.flat.numel, device=dev, dtype=th.float32)
        d_a = d_b = None
        samples_mode = bool(inputs.pos)
        if want_input_grads:
            if samples_mode:
                d_a = th.zeros((n, 3), device=dev, dtype=th.float32)   # the kernel accumulates (+=)
                d_b = th.zeros((n, 3), device=dev, dtype=th.float32)
            else:
                d_a = th.zeros((n_rays, 3), device=dev, dtype=th.float32)
                d_b = th.zeros((n_rays, 3), device=dev, dtype=th.float32)
        cp, cd = self.pe_cfgs()
        stream = th.cuda.current_stream().cuda_stream
        with th.cuda.device(dev):
            self._timed("mlp_bwd_inputs" if want_input_grads else "mlp_bwd", lambda: check(lib().nerfb200_mlp_bwd(
                C.byref(cb.program), _ptr(self.wpack), C.byref(inputs), C.byref(cp), C.byref(cd),
                _ptr(self.pe_pos.alpha_tensor()), _ptr(self.pe_dir.alpha_tensor()), _ptr(sigma), _ptr(rgb),
                _ptr(g_sigma), _ptr(g_rgb), _ptr(masks), cm.mask_words_per_tile, _ptr(dy_stash),
                cb.head_sigma_col3, cb.pos_grad_cols if want_input_grads else 0,
                cb.dir_grad_cols if want_input_grads else 0,
                None if samples_mode else _ptr(d_a), None if samples_mode else _ptr(d_b),
                _ptr(d_a) if samples_mode else None, _ptr(d_b) if samples_mode else None, stream), "mlp_bwd"))
            items_dev, n_items = self._items(n_tiles)
            self._timed("mlp_wgrad", lambda: check(lib().nerfb200_mlp_wgrad(
                _ptr(items_dev), n_items, _ptr(stash), cm.stash_slabs_per_tile, _ptr(dy_stash),
                cb.dy_slabs_per_tile, None, 0, _ptr(self.flat.flat), _ptr(flat_grad), stream), "mlp_wgrad"))
        return flat_grad, d_a, d_b


def make_inputs(n: int, S: int, t_mode: int, ray_o=None, ray_d=None, t_start=None, t_end=None,
                pixel_width=None, pos=None, dir=None, pixel_width_per_sample=False) -> NbMlpInputs:
    return NbMlpInputs(N=n, S=S, t_mode=t_mode, ray_o=_ptr(ray_o), ray_d=_ptr(ray_d),
                       t_start=_ptr(t_start), t_end=_ptr(t_end), pixel_width=_ptr(pixel_width),
                       pos=_ptr(pos), dir=_ptr(dir), pixel_width_per_sample=int(pixel_width_per_sample),
                       reserved=0)
