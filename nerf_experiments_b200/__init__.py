"""nerf_experiments_b200 — B200-native (sm_100a) NeRF train/render hot path behind the module
surface of sarphiv/nerf-experiments.  The compute lives in libnerfb200.so
(hand-written CUDA, C ABI in include/nerfb200.h); this package is the thin host-side mirror of
the reference's nn.Module interface."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
