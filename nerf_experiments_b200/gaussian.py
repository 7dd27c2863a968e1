"""Learnable activations of the GARF / SARF / Gabor networks: same module surface (constructor
arguments, parameter names, seeded initialisation) as the reference's barf/gaussian.py:37-63
(== garf/gaussian.py), sarf/activation.py:40-66 and gaborf/gabor.py:32-64; forward and backward
run in the CUDA kernels of csrc/activations.cu (no autograd graph of elementwise ops, no
Inductor)."""
import torch as th
import torch.nn as nn

from . import _lib, ops


class GaussAct(nn.Module):
    """y = exp(-x^2 (inv_std^2 + 1e-6)) — barf/gaussian.py:37-63."""

    def __init__(self, features_in: int, inv_standard_deviation_init_min: float = 0.,
                 inv_standard_deviation_init_max: float = 1.):
        super().__init__()
        # NOTE (reference): a negative inverse standard deviation is allowed, only its square is used
        self.inv_standard_deviation = nn.Parameter(
            th.rand(features_in) * (inv_standard_deviation_init_max - inv_standard_deviation_init_min)
            + inv_standard_deviation_init_min)

    def forward(self, x: th.Tensor):
        return ops.activation(_lib.ACT_GAUSS, x, self.inv_standard_deviation)


class SarfAct(nn.Module):
    """y = cos(f / (x'^2 + 1/f^2)) exp(-x'^2), x' = (signbit(x)*2-1)(|x|+1e-4) —
    sarf/activation.py:40-66 (the expression the reference evaluates; its custom Function is
    commented out at :66)."""

    def __init__(self, features_in: int, frequency_init_min: float, frequency_init_max: float):
        super().__init__()
        self.frequency = nn.Parameter(
            th.rand(features_in) * (frequency_init_max - frequency_init_min) + frequency_init_min)

    def forward(self, x: th.Tensor):
        return ops.activation(_lib.ACT_SARF, x, self.frequency)


class GaborAct(nn.Module):
    """y = exp(-(inv_std^2 + 1e-6) x^2) cos(spread x) — gaborf/gabor.py:32-64."""

    def __init__(self, features_in: int, inv_standard_deviation_init_min: float = 0.,
                 inv_standard_deviation_init_max: float = 1.):
        super().__init__()
        self.inv_standard_deviation = nn.Parameter(
            th.rand(features_in) * (inv_standard_deviation_init_max - inv_standard_deviation_init_min)
            + inv_standard_deviation_init_min)
        self.spread = nn.Parameter(th.rand(features_in) * 2 * th.pi)

    def forward(self, x: th.Tensor):
        return ops.activation(_lib.ACT_GABOR, x, self.inv_standard_deviation, self.spread)
