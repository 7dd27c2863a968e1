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
from .mlp_program import (SLAB_PE_DIR, SLAB_PE_POS, CompiledMlp, LayerSpec, Linear, compile_forward,
                          to_device_array)


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

    def __init__(self, layers_fn, flat: FlatParams, pe_pos, pe_dir, sigma_bias: float = 0.0):
        self.layers_fn = layers_fn        # callable(FlatParams) -> List[LayerSpec]
        self.flat = flat
        self.pe_pos = pe_pos
        self.pe_dir = pe_dir
        self.sigma_bias = sigma_bias
        self.compiled: Optional[CompiledMlp] = None
        self.device = None
        self._packed_sig = None

    # -- compilation / packing ------------------------------------------------------------
    def prepare(self, device):
        flat = self.flat.ensure(device)
        if self.compiled is None or self.device != device:
            self.compiled = compile_forward(self.layers_fn(self.flat))
            cm = self.compiled
            self.device = device
            self.wpack = th.empty(max(cm.wpack_bytes, 1024), device=device, dtype=th.uint8)
            self.bias = th.zeros(max(cm.bias_floats, 1), device=device, dtype=th.float32)
            self.chunks_dev = to_device_array(cm.pack_chunks, NbPackChunk, device)
            self.biases_dev = to_device_array(cm.pack_biases, NbPackBias, device)
            self._packed_sig = None
        sig = self.flat.signature()
        if sig != self._packed_sig:
            cm = self.compiled
            with th.cuda.device(device):
                check(lib().nerfb200_mlp_pack(_ptr(flat), _ptr(self.chunks_dev), len(cm.pack_chunks),
                                              _ptr(self.wpack), _ptr(self.biases_dev), len(cm.pack_biases),
                                              _ptr(self.bias), th.cuda.current_stream().cuda_stream),
                      "mlp_pack")
            self._packed_sig = sig
        return self.compiled

    def pe_cfgs(self):
        cp = self.pe_pos.describe()
        cd = self.pe_dir.describe()
        cp.slab, cp.stash_slab = SLAB_PE_POS, 0
        cd.slab, cd.stash_slab = SLAB_PE_DIR, 1
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
        with th.cuda.device(dev):
            check(lib().nerfb200_mlp_fwd(C.byref(cm.program), _ptr(self.wpack), _ptr(self.bias),
                                         C.byref(inputs), C.byref(cp), C.byref(cd),
                                         _ptr(self.pe_pos.alpha_tensor()), _ptr(self.pe_dir.alpha_tensor()),
                                         float(self.sigma_bias), _ptr(sigma), _ptr(rgb), _ptr(stash),
                                         _ptr(masks), th.cuda.current_stream().cuda_stream), "mlp_fwd")
        return sigma, rgb, stash, masks


def make_inputs(n: int, S: int, t_mode: int, ray_o=None, ray_d=None, t_start=None, t_end=None,
                pixel_width=None, pos=None, dir=None, pixel_width_per_sample=False) -> NbMlpInputs:
    return NbMlpInputs(N=n, S=S, t_mode=t_mode, ray_o=_ptr(ray_o), ray_d=_ptr(ray_d),
                       t_start=_ptr(t_start), t_end=_ptr(t_end), pixel_width=_ptr(pixel_width),
                       pos=_ptr(pos), dir=_ptr(dir), pixel_width_per_sample=int(pixel_width_per_sample),
                       reserved=0)
