"""Positional-encoding modules with the reference's class names, constructor arguments,
attributes and buffers (reference barf/positional_encodings.py), evaluated by CUDA kernels.

Inside a fused network the encoders are not called at all: `describe()` hands their
configuration to the fused MLP kernel, which evaluates the encoding in registers.  Calling
`forward` on its own runs the stand-alone kernel (nerfb200_pe_fwd / nerfb200_pe_bwd).
"""
import math
from typing import Optional

import torch as th
import torch.nn as nn

from . import _lib
from ._lib import NbPeCfg, check, lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _col(t, n, name):
    """(n,1)/(n,) tensor or python scalar -> contiguous (n,) float32 CUDA tensor."""
    return t


class _PeFunction(th.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, alpha, out_dim, x, dir, pixel_width, t_start, t_end):
        if not x.is_cuda:
            raise RuntimeError("positional encodings run on CUDA only (nerfb200 has no CPU fallback)")
        n = x.shape[0]
        x = x.contiguous().float()

        def prep(t, width):
            if t is None:
                return None
            if not isinstance(t, th.Tensor):
                t = th.full((n, width), float(t), device=x.device)
            return t.to(x.device).float().expand(n, width).contiguous() if t.dim() == 2 else t.to(x.device).float().reshape(-1, 1).expand(n, width).contiguous()

        dir_c = prep(dir, 3) if dir is not None else None
        pw = prep(pixel_width, 1)
        t0 = prep(t_start, 1)
        t1 = prep(t_end, 1)
        out = th.empty((n, out_dim), device=x.device, dtype=th.float32)
        with th.cuda.device(x.device):
            check(lib().nerfb200_pe_fwd(cfg, _ptr(alpha), _ptr(x), _ptr(dir_c), _ptr(pw), _ptr(t0), _ptr(t1),
                                        n, _ptr(out), th.cuda.current_stream().cuda_stream), "pe_fwd")
        ctx.cfg, ctx.alpha = cfg, alpha
        ctx.save_for_backward(x, dir_c, pw, t0, t1)
        return out

    @staticmethod
    def backward(ctx, g):
        x, dir_c, pw, t0, t1 = ctx.saved_tensors
        n = x.shape[0]
        d_pos = th.empty_like(x)
        d_dir = th.empty_like(x) if dir_c is not None and ctx.needs_input_grad[4] else None
        with th.cuda.device(x.device):
            check(lib().nerfb200_pe_bwd(ctx.cfg, _ptr(ctx.alpha), _ptr(x), _ptr(dir_c), _ptr(pw), _ptr(t0),
                                        _ptr(t1), _ptr(g.contiguous()), n, _ptr(d_pos), _ptr(d_dir),
                                        th.cuda.current_stream().cuda_stream), "pe_bwd")
        return None, None, None, d_pos, d_dir, None, None, None


class PositionalEncoding(nn.Module):
    """Base class (barf/positional_encodings.py:7-14)."""

    def __init__(self):
        super().__init__()
        self.output_dim = None
        self.space_dimensions = None

    def describe(self) -> NbPeCfg:
        raise NotImplementedError()

    def alpha_tensor(self) -> Optional[th.Tensor]:
        return None

    def forward(self, x, dir=None, pixel_width=None, t_start=None, t_end=None):
        return _PeFunction.apply(self.describe(), self.alpha_tensor(), self.output_dim, x, dir,
                                 pixel_width, t_start, t_end)


class IdentityPositionalEncoding(PositionalEncoding):
    """barf/positional_encodings.py:17-25."""

    def __init__(self, space_dimensions: int = 3):
        super().__init__()
        self.output_dim = space_dimensions
        self.space_dimensions = space_dimensions
        self.levels = 0

    def describe(self):
        return NbPeCfg(kind=_lib.PE_IDENTITY, levels=0, include_identity=1, use_mask=0,
                       distribute_variance=0, scale=1.0, pixel_width_sigma=0.0, slab=-1, stash_slab=-1)

    def forward(self, x, dir=None, pixel_width=None, t_start=None, t_end=None):
        assert x.shape[1] == self.space_dimensions
        return x


class FourierFeatures(PositionalEncoding):
    """barf/positional_encodings.py:28-57."""

    def __init__(self, levels: int, scale: float = 2 * math.pi, space_dimensions: int = 3):
        super().__init__()
        self.levels = levels
        self.scale = scale
        self.space_dimensions = space_dimensions
        self.output_dim = levels * 2 * space_dimensions

    def describe(self):
        return NbPeCfg(kind=_lib.PE_FOURIER, levels=self.levels, include_identity=0, use_mask=0,
                       distribute_variance=0, scale=float(self.scale), pixel_width_sigma=0.0,
                       slab=-1, stash_slab=-1)


class BarfPositionalEncoding(PositionalEncoding):
    """barf/positional_encodings.py:61-148: Fourier features under the BARF coarse-to-fine mask.
    `alpha` stays a registered buffer (it is part of the reference's checkpoints); the kernels
    read it on the device, so no host sync is needed to build the mask."""

    def __init__(self, levels: int, alpha_start: float, alpha_increase_start_epoch: float,
                 alpha_increase_end_epoch: float, include_identity: bool = True,
                 scale: float = 2 * math.pi, space_dimensions: int = 3):
        super().__init__()
        self.levels = levels
        self.alpha_start = alpha_start
        self.alpha_increase_start_epoch = alpha_increase_start_epoch
        self.alpha_increase_end_epoch = alpha_increase_end_epoch
        self.include_identity = include_identity
        self.scale = scale
        self.space_dimensions = space_dimensions
        self.output_dim = (levels * 2 + include_identity) * space_dimensions
        self.register_buffer("alpha", th.tensor(float(alpha_start)))

    def update_alpha(self, epoch: float) -> None:
        lo, hi = self.alpha_increase_start_epoch, self.alpha_increase_end_epoch
        if epoch < lo:
            value = self.alpha_start
        elif epoch < hi:
            value = self.alpha_start + (epoch - lo) * (self.levels - self.alpha_start) / (hi - lo)
        else:
            value = float(self.levels)
        # in place: keeps the device pointer the kernels (and CUDA graphs) hold
        self.alpha.fill_(float(value))
        self.alpha_value = float(value)     # host copy: schedules that depend on alpha need no device sync

    def compute_mask(self, alpha: th.Tensor) -> th.Tensor:
        k = th.arange(self.levels, device=alpha.device, dtype=th.float32)
        ramp = th.floor(alpha.float())
        part = (1 - th.cos((alpha.float() - ramp) * th.pi)) / 2
        mask = th.where(k < ramp, th.ones_like(k), th.where(k == ramp, part, th.zeros_like(k)))
        return mask.repeat(self.space_dimensions).view(1, -1)

    def alpha_tensor(self):
        return self.alpha

    def describe(self):
        return NbPeCfg(kind=_lib.PE_FOURIER, levels=self.levels, include_identity=int(self.include_identity),
                       use_mask=1, distribute_variance=0, scale=float(self.scale),
                       pixel_width_sigma=0.0, slab=-1, stash_slab=-1)


class IntegratedFourierFeatures(PositionalEncoding):
    """barf/positional_encodings.py:151-240: Mip-NeRF integrated encoding of conical frustums."""

    def __init__(self, levels: int, scale: float = 2 * math.pi, include_identity=True,
                 distribute_variance: Optional[bool] = False):
        super().__init__()
        self.levels = levels
        self.space_dimensions = 3
        self.scale = scale
        self.include_identity = include_identity
        self.output_dim = (levels * 2 + include_identity) * self.space_dimensions
        self.distribute_variance = distribute_variance
        self.pixel_width_sigma = None

    def describe(self):
        if self.pixel_width_sigma is None:
            raise TypeError("IntegratedFourierFeatures.pixel_width_sigma must be set before use "
                            "(the reference compares it with 0.25)")
        return NbPeCfg(kind=_lib.PE_INTEGRATED, levels=self.levels, include_identity=int(bool(self.include_identity)),
                       use_mask=0, distribute_variance=int(bool(self.distribute_variance)),
                       scale=float(self.scale), pixel_width_sigma=float(self.pixel_width_sigma),
                       slab=-1, stash_slab=-1)

    def forward(self, pos, dir, pixel_width, t_start, t_end, include_identity=None, diagnose=False):
        if pos.shape[1] != 3:
            raise ValueError(f"Only 3D supported - was {pos.shape[1]}D")
        if diagnose:
            raise NotImplementedError("diagnose=True is a plotting aid of the reference and is not provided")
        cfg = self.describe()
        out_dim = self.output_dim
        if include_identity is not None and bool(include_identity) != bool(self.include_identity):
            cfg.include_identity = int(bool(include_identity))
            out_dim = (self.levels * 2 + int(bool(include_identity))) * 3
        return _PeFunction.apply(cfg, self.alpha_tensor(), out_dim, pos, dir, pixel_width, t_start, t_end)


class IntegratedBarfFourierFeatures(BarfPositionalEncoding):
    """barf/positional_encodings.py:242-282: integrated encoding under the BARF mask."""

    def __init__(self, levels: int, alpha_start: float, alpha_increase_start_epoch: float,
                 alpha_increase_end_epoch: float, include_identity: bool = True,
                 scale: float = 2 * math.pi, distribute_variance=True):
        super().__init__(levels=levels, alpha_start=alpha_start,
                         alpha_increase_start_epoch=alpha_increase_start_epoch,
                         alpha_increase_end_epoch=alpha_increase_end_epoch,
                         include_identity=include_identity, scale=scale, space_dimensions=3)
        self.distribute_variance = distribute_variance
        self.pixel_width_sigma = None

    def describe(self):
        if self.pixel_width_sigma is None:
            raise TypeError("IntegratedBarfFourierFeatures.pixel_width_sigma must be set before use")
        return NbPeCfg(kind=_lib.PE_INTEGRATED, levels=self.levels, include_identity=int(self.include_identity),
                       use_mask=1, distribute_variance=int(bool(self.distribute_variance)),
                       scale=float(self.scale), pixel_width_sigma=float(self.pixel_width_sigma),
                       slab=-1, stash_slab=-1)

    def forward(self, pos, dir, pixel_width, t_start, t_end):
        return _PeFunction.apply(self.describe(), self.alpha, self.output_dim, pos, dir, pixel_width,
                                 t_start, t_end)
