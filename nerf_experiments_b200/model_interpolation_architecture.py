"""NerfBaseModel / NerfModel with the reference's constructor, attributes, parameter-group
registry and state-dict keys (reference barf/model_interpolation_architecture.py:11-161); the
forward pass is the fused sm_100a kernel instead of 12 cuBLAS GEMMs + elementwise launches."""
from typing import Iterator, List

import torch as th
import torch.nn as nn

from .fused_mlp import FlatParams, FusedField, make_inputs
from .mlp_program import Linear, nerf_model_layers
from .positional_encodings import PositionalEncoding


class NerfBaseModel(nn.Module):
    """Parameter-group registry (barf/model_interpolation_architecture.py:11-29)."""

    def __init__(self):
        super().__init__()
        self.param_groups: List[dict] = []

    def _add_param_group(self, parameters: Iterator, learning_rate_start: float,
                         learning_rate_stop: float, learning_rate_decay_end: float,
                         weight_decay: float = 0.0):
        self.param_groups.append({
            "parameters": parameters,
            "learning_rate_start": learning_rate_start,
            "learning_rate_stop": learning_rate_stop,
            "learning_rate_decay_end": learning_rate_decay_end,
            "weight_decay": weight_decay,
        })


class NerfModel(NerfBaseModel):
    def __init__(self, n_hidden: int, hidden_dim: int, delayed_direction: bool, delayed_density: bool,
                 n_segments: int, position_encoder: PositionalEncoding,
                 direction_encoder: PositionalEncoding, learning_rate_start: float = 5e-4,
                 learning_rate_stop: float = 5e-5, learning_rate_decay_end: float = 0):
        super().__init__()
        if n_segments == 0:
            raise NotImplementedError("n_segments must be greater than 0")
        self.n_hidden = n_hidden
        self.hidden_dim = hidden_dim
        self.delayed_direction = delayed_direction
        self.delayed_density = delayed_density
        self.n_segments = n_segments
        self.position_encoder = position_encoder
        self.direction_encoder = direction_encoder

        p_dim, d_dim = position_encoder.output_dim, direction_encoder.output_dim
        self.model_segments = nn.ModuleList()
        for i in range(n_segments):
            d_in = p_dim + (0 if delayed_direction else d_dim) + (hidden_dim if i > 0 else 0)
            d_out = hidden_dim + (1 if (not delayed_density and i == n_segments - 1) else 0)
            self.model_segments.append(self._build_segment(d_in, hidden_dim, d_out))
        self.model_color = nn.Sequential(
            nn.Linear(hidden_dim + (d_dim if delayed_direction else 0), hidden_dim // 2),
            nn.ReLU(inplace=True),
            nn.Linear(hidden_dim // 2, 3 + (1 if delayed_density else 0)),
        )
        self.relu = nn.ReLU(inplace=True)
        self.softplus = nn.Softplus(threshold=8)
        self.sigmoid = nn.Sigmoid()
        self._add_param_group(self.parameters(), learning_rate_start, learning_rate_stop, learning_rate_decay_end)

        self._flat = None
        self._field = None

    def _build_segment(self, d_in: int, d_hidden: int, d_out: int) -> nn.Module:
        if self.n_hidden == 0:
            return nn.Linear(d_in, d_out)
        # construction order of the reference (first, last, then the intermediate Linears),
        # so that seeded default initialisation draws the same numbers per parameter
        first = nn.Linear(d_in, d_hidden)
        last = nn.Linear(d_hidden, d_out)
        mids: List[nn.Module] = []
        for _ in range(self.n_hidden - 1):
            mids += [nn.ReLU(True), nn.Linear(d_hidden, d_hidden)]
        return nn.Sequential(first, *mids, nn.ReLU(True), last)

    # reference spelling kept for callers
    def contruct_model_density(self, input_dim: int, hidden_dim: int, output_dim: int) -> nn.Module:
        return self._build_segment(input_dim, hidden_dim, output_dim)

    # ---- fused field -------------------------------------------------------------------------
    def linears(self) -> dict:
        out = {}
        for i, seg in enumerate(self.model_segments):
            if isinstance(seg, nn.Linear):
                out[f"model_segments.{i}"] = seg
            else:
                for k, m in enumerate(seg):
                    if isinstance(m, nn.Linear):
                        out[f"model_segments.{i}.{k}"] = m
        out["model_color.0"] = self.model_color[0]
        out["model_color.2"] = self.model_color[2]
        return out

    def fused_field(self, flat: FlatParams = None) -> FusedField:
        """The compiled fused field of this network; `flat` lets a caller (the render module)
        place the parameters of several networks in one shared flat buffer."""
        if self._field is None or (flat is not None and flat is not self._flat):
            self._flat = flat if flat is not None else FlatParams(list(self.parameters()))

            def layers_fn(fp: FlatParams):
                lins = {name: Linear(fp.offset_of(m.weight), fp.offset_of(m.bias), m.out_features, m.in_features)
                        for name, m in self.linears().items()}
                return nerf_model_layers(lins, self.n_hidden, self.hidden_dim, self.n_segments,
                                         self.delayed_direction, self.delayed_density,
                                         self.position_encoder.output_dim, self.direction_encoder.output_dim)

            self._field = FusedField(layers_fn, self._flat, self.position_encoder, self.direction_encoder,
                                     own_params=list(self.parameters()))
        return self._field

    def forward(self, pos: th.Tensor, dir: th.Tensor, pixel_width: th.Tensor = None,
                t_start: th.Tensor = None, t_end: th.Tensor = None):
        """(density (N,), rgb (N,3)) for per-sample positions / directions (:96-141)."""
        from .field_function import field_samples
        return field_samples(self, pos, dir, pixel_width, t_start, t_end)

    def list_segments(self):
        for i, segment in enumerate(self.model_segments):
            print(f"Segment {i}: {segment}")
        print(f"Final layer: {self.model_color}")
