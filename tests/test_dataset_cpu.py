"""Ray generation / batch assembly oracle against the unmodified reference's outputs
(tests/golden/dataset.npz) and the host-side blur schedule of the GPU batcher."""
import os

import numpy as np
import pytest
import torch as th

from oracle import ref_dataset

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset.npz")


def _g():
    z = np.load(G)
    return {k: (th.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def test_oracle_matches_reference_ray_generation():
    g = _g()
    H, W, focal = int(g["H"]), int(g["W"]), float(g["focal"])
    grid = ref_dataset.directions_meshgrid(H, W, focal)
    assert th.equal(grid, g["grid"])
    assert th.equal(ref_dataset.directions_meshgrid(4, 2, 4.0), g["grid_4x2_f4"])   # notebook cell 8
    o, d = ref_dataset.meshgrid_to_world(grid, g["c2w"])
    assert th.equal(o, g["o_raw"]) and th.equal(d, g["d_raw"])
    # unit norm, -z forward for the identity pose
    assert th.allclose(grid.norm(dim=1), th.ones(H * W), atol=1e-6) and (grid[:, 2] < 0).all()


def test_oracle_matches_reference_blur_interpolation():
    g = _g()
    sig = [float(s) for s in g["sigmas"]]
    colors = g["images"].view(-1, len(sig), 3)
    for s in (0.1, 0.25, 1.0, 3.0, 5.5, 8.0):
        assert th.equal(ref_dataset.blurred_pixel_colors(colors, sig, s), g[f"blur_{s}"]), s


def test_batcher_blur_schedule_matches_reference_indices():
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher
    b = GpuRayBatcher.__new__(GpuRayBatcher)
    b.gaussian_blur_sigmas = [8.0, 4.0, 2.0, 0.0]
    b.n_sigmas = 4
    assert b.blur_levels(None) == (-1, -1, 0.0)
    assert b.blur_levels(0.2) == (3, 3, 1.0) and b.blur_levels(9.0) == (0, 0, 1.0)
    lo, hi, c = b.blur_levels(3.0)
    assert (lo, hi) == (1, 2) and c == pytest.approx((3.0 - 2.0) / (4.0 - 2.0 + 1e-8))
    lo, hi, c = b.blur_levels(1.0)
    assert (lo, hi) == (2, 3) and c == pytest.approx(0.5, rel=1e-6)
