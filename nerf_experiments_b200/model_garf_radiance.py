"""GARF radiance network — module surface of reference garf/model_radiance.py:9-96
(== barf/model_garf_radiance.py:9-113): same constructor, sub-module names (state-dict keys
`model_density_1.{0..7}`, `model_density_2.{0..6}`, `model_color.{0..3}`), seeded initialisation,
`parameters_linear()` / `parameters_gaussian()` and `forward(pos, dir) -> (rgb, density)`.

The whole network (11 Linear layers, 8 Gaussian activations, the residual and the two raw-xyz /
direction concatenations) runs as ONE fused kernel per pass on a 128-sample tile that stays on the
SM — csrc/garf_fwd.cu, garf_bwd.cu, mlp_wgrad.cu; tile program from garf_program.py. No library GEMM
and no CPU path: a CPU tensor raises. th.compile of the reference is dropped (no tracing compiler)."""
from typing import Iterator

import torch as th
import torch.nn as nn

from .fused_garf import FusedGarfField, garf_rays, garf_samples
from .fused_mlp import FlatParams
from .garf_program import GaussLinear, compile_proposal, compile_radiance
from .gaussian import GaussAct
from .mlp_program import Linear


class _GaussNetBase(nn.Module):
    def __init__(self, gaussian_init_min: float, gaussian_init_max: float):
        super().__init__()
        self.gaussian_init_min = gaussian_init_min
        self.gaussian_init_max = gaussian_init_max
        self._parameters_linear: list = []
        self._parameters_gaussian: list = []
        self._flat = None
        self._field = None

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

    # ---- fused field -----------------------------------------------------------------------------
    def _gauss_linears(self, fp: FlatParams, seqs):
        """[GaussLinear] of every Linear of the given Sequentials, in order (a GaussAct behind a Linear is
        its activation)."""
        out = []
        for seq in seqs:
            mods = list(seq)
            for i, m in enumerate(mods):
                if isinstance(m, nn.Linear):
                    nxt = mods[i + 1] if i + 1 < len(mods) else None
                    g = fp.offset_of(nxt.inv_standard_deviation) if isinstance(nxt, GaussAct) else -1
                    out.append(GaussLinear(Linear(fp.offset_of(m.weight), fp.offset_of(m.bias), m.out_features,
                                                  m.in_features), g))
        return out

    def _compile(self, fp: FlatParams):
        raise NotImplementedError

    def fused_field(self, flat: FlatParams = None) -> FusedGarfField:
        """The compiled fused field of this network; `flat` lets a caller (engine.TrainEngine) place the
        parameters of several networks in one shared flat buffer."""
        if self._field is None or (flat is not None and flat is not self._flat):
            self._flat = flat if flat is not None else FlatParams(list(self.parameters()))
            self._field = FusedGarfField(self._compile, self._flat, own_params=list(self.parameters()))
        return self._field


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

    def _compile(self, fp: FlatParams):
        layers = self._gauss_linears(fp, (self.model_density_1, self.model_density_2, self.model_color))
        return compile_radiance(layers[:8], layers[8:])

    def forward(self, pos: th.Tensor, dir: th.Tensor):
        """(rgb (N,3), density (N,)) for per-sample positions / directions (garf/model_radiance.py:84-96)."""
        density, rgb = garf_samples(self, pos, dir)
        return rgb, density

    def forward_rays(self, ray_origs: th.Tensor, ray_dirs: th.Tensor, t_starts: th.Tensor, t_ends: th.Tensor):
        """(rgb (B,S,3), density (B,S)) at the mid-points of the bins: GarfModel._get_positions +
        repeat_interleave + forward (garf/model_garf.py:105,163-188) without materialising positions."""
        density, rgb = garf_rays(self, ray_origs, ray_dirs, t_starts, t_ends)
        return rgb, density
