"""GARF radiance network — module surface of reference garf/model_radiance.py:9-96
(== barf/model_garf_radiance.py:9-113): same constructor, sub-module names (state-dict keys
`model_density_1.{0..7}`, `model_density_2.{0..6}`, `model_color.{0..3}`), seeded initialisation,
`parameters_linear()` / `parameters_gaussian()` and `forward(pos, dir) -> (rgb, density)`.

Round-1 status: the Gaussian activations (forward, input / parameter / bias gradients) run in the
CUDA kernels of csrc/activations.cu; the Linear layers are plain library GEMMs (cuBLAS through
torch) on the tensor cores with TF32 operands — the precision class the reference trains at
(`matmul_precision = "fp32"` restores fp32 GEMMs, `"bf16"` runs bf16 operands with bf16
activations between the layers). The layers are up to 1024 wide, which
does not fit the 256-column tile program of the fused kernel: fusing this network
is the round-2 item. th.compile of the reference is dropped (no tracing compiler)."""
from typing import Iterator

import torch as th
import torch.nn as nn

from . import _lib, ops
from .gaussian import GaussAct


class _GaussNetBase(nn.Module):
    def __init__(self, gaussian_init_min: float, gaussian_init_max: float):
        super().__init__()
        self.gaussian_init_min = gaussian_init_min
        self.gaussian_init_max = gaussian_init_max
        self._parameters_linear: list = []
        self._parameters_gaussian: list = []
        # GEMM arithmetic: "tf32" (default: fp32 tensors, TF32 tensor-core GEMMs), "bf16" (bf16 operands,
        # fp32 accumulation and pre-activations, bf16 activations between the layers: measured SLOWER
        # with cuBLAS, whose bf16-in / fp32-out GEMMs fall back to pre-Blackwell kernels) or "fp32"
        self.matmul_precision = "tf32"

    def _run(self, seq: nn.Sequential, x: th.Tensor) -> th.Tensor:
        """seq(x) with every Linear (+ GaussAct) pair as one fused-gradient op; fp32 in, fp32 out."""
        mods = list(seq)
        n_lin = sum(isinstance(m, nn.Linear) for m in mods)
        i = seen = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear) and x.is_cuda:
                seen += 1
                last = seen == n_lin
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if isinstance(nxt, GaussAct):
                    x = ops.linear_activation(x, m.weight, m.bias, _lib.ACT_GAUSS, nxt.inv_standard_deviation,
                                              None, self.matmul_precision, last)
                    i += 2
                    continue
                x = ops.linear_activation(x, m.weight, m.bias, -1, None, None, self.matmul_precision, last)
            else:
                x = m(x.float() if x.dtype == th.bfloat16 else x)
            i += 1
        return x.float() if x.dtype == th.bfloat16 else x

    def _create_linear(self, features_in: int, features_out: int) -> nn.Linear:
        linear = nn.Linear(features_in, features_out)
        self._parameters_linear.append(linear.weight)
        self._parameters_linear.append(linear.bias)
        return linear

    def _create_gaussian(self, features_in) -> GaussAct:
        act = GaussAct(features_in, self.gaussian_init_min, self.gaussian_init_max)
        self._parameters_gaussian.append(act.inv_standard_deviation)
        return act

    def parameters_linear(self) -> Iterator[nn.Parameter]:
        return iter(self._parameters_linear)

    def parameters_gaussian(self) -> Iterator[nn.Parameter]:
        return iter(self._parameters_gaussian)


class RadianceNetwork(_GaussNetBase):
    def __init__(self, gaussian_init_min: float, gaussian_init_max: float):
        super().__init__(gaussian_init_min, gaussian_init_max)
        # creation order = the reference's (it fixes the seeded initial values)
        self.model_density_1 = nn.Sequential(
            self._create_linear(3, 1024), self._create_gaussian(1024),
            self._create_linear(1024, 256), self._create_gaussian(256),
            self._create_linear(256, 128), self._create_gaussian(128),
            self._create_linear(128, 128), self._create_gaussian(128))
        self.model_density_2 = nn.Sequential(
            self._create_linear(128 + 3, 512), self._create_gaussian(512),
            self._create_linear(512, 256), self._create_gaussian(256),
            self._create_linear(256, 128), self._create_gaussian(128),
            self._create_linear(128, 128 + 1))
        self.softplus = nn.Softplus(threshold=8)
        self.model_color = nn.Sequential(
            self._create_linear(128 + 3, 256), self._create_gaussian(256),
            self._create_linear(256, 3), nn.Sigmoid())

    def forward(self, pos: th.Tensor, dir: th.Tensor):
        z1 = self._run(self.model_density_1, pos)
        z2 = self._run(self.model_density_2, th.cat((z1, pos), dim=1))
        density = self.softplus(z2[:, 128] - 1)
        rgb = self._run(self.model_color, th.cat((z1[:, :128] + z2[:, :128], dir), dim=1))
        return rgb, density
