"""GARF proposal network — module surface of reference garf/model_proposal.py:9-56
(== barf/model_garf_proposal.py): `model.{0..7}` state-dict keys, `forward(pos) -> (N, 1)`
density, as one fused kernel per pass (see model_garf_radiance.py)."""
import torch as th
import torch.nn as nn

from .fused_garf import garf_rays, garf_samples
from .garf_program import compile_proposal
from .model_garf_radiance import _GaussNetBase


class ProposalNetwork(_GaussNetBase):
    def __init__(self, gaussian_init_min: float, gaussian_init_max: float):
        super().__init__(gaussian_init_min, gaussian_init_max)
        self.model = nn.Sequential(
            self._create_linear(3, 512), self._create_gaussian(512),
            self._create_linear(512, 256), self._create_gaussian(256),
            self._create_linear(256, 128), self._create_gaussian(128),
            self._create_linear(128, 1), nn.Softplus(threshold=8))

    def _compile(self, fp):
        return compile_proposal(self._gauss_linears(fp, (self.model,)))

    def forward(self, pos: th.Tensor) -> th.Tensor:
        """(N,1) density for per-sample positions (garf/model_proposal.py:55-56)."""
        return garf_samples(self, pos, None)[:, None]

    def forward_rays(self, ray_origs: th.Tensor, ray_dirs: th.Tensor, t_starts: th.Tensor, t_ends: th.Tensor):
        """(B,S) density at the mid-points of the bins (the closure of garf/model_garf.py:127-141)."""
        return garf_rays(self, ray_origs, ray_dirs, t_starts, t_ends)
