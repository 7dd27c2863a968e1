"""GARF proposal network — module surface of reference garf/model_proposal.py:9-56
(== barf/model_garf_proposal.py): `model.{0..7}` state-dict keys, `forward(pos) -> (N, 1)`
density. See model_garf_radiance.py for the round-1 status of the Linear layers."""
import torch as th
import torch.nn as nn

from .model_garf_radiance import _GaussNetBase


class ProposalNetwork(_GaussNetBase):
    def __init__(self, gaussian_init_min: float, gaussian_init_max: float):
        super().__init__(gaussian_init_min, gaussian_init_max)
        self.model = nn.Sequential(
            self._create_linear(3, 512), self._create_gaussian(512),
            self._create_linear(512, 256), self._create_gaussian(256),
            self._create_linear(256, 128), self._create_gaussian(128),
            self._create_linear(128, 1), nn.Softplus(threshold=8))

    def forward(self, pos: th.Tensor) -> th.Tensor:
        return self._run(self.model, pos)
