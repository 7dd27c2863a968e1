"""Generates tests/golden/dataset.npz from the UNMODIFIED reference (this container only):
ray directions / origins from ImagePoseDataset's static methods and the blur interpolation of
ImagePoseDataModule.get_blurred_pixel_colors.  Run: python tests/golden/make_golden_dataset.py"""
import os
import sys
import types

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
sys.path.insert(0, "/root/reference/barf")


def main():
    import data_module
    import dataset
    DS = dataset.ImagePoseDataset
    g = th.Generator().manual_seed(9)
    H, W, focal, N = 6, 5, 7.5, 3
    grid = DS._get_directions_meshgrid(H, W, focal)
    # random rigid poses
    A = th.randn((N, 3, 3), generator=g)
    Q, _ = th.linalg.qr(A)
    c2w = th.eye(4).repeat(N, 1, 1)
    c2w[:, :3, :3] = Q
    c2w[:, :3, 3] = th.randn((N, 3), generator=g) * 3
    c2w_noisy = c2w.clone()
    c2w_noisy[:, :3, 3] += 0.1 * th.randn((N, 3), generator=g)
    o_r, d_r = DS._meshgrid_to_world(grid, c2w)
    o_n, d_n = DS._meshgrid_to_world(grid, c2w_noisy)
    sigmas = [8.0, 4.0, 2.0, 0.0]
    images = th.rand((N, H, W, len(sigmas), 3), generator=g)
    out = dict(H=H, W=W, focal=focal, grid=grid, c2w=c2w, c2w_noisy=c2w_noisy, o_raw=o_r, d_raw=d_r, o_noisy=o_n,
               d_noisy=d_n, images=images, sigmas=np.array(sigmas),
               grid_4x2_f4=DS._get_directions_meshgrid(4, 2, 4.0))      # the notebook's cell-8 example
    fake = types.SimpleNamespace(gaussian_blur_sigmas=sigmas)
    colors = images.view(N * H * W, len(sigmas), 3)
    for s in (0.1, 0.25, 1.0, 3.0, 5.5, 8.0):
        batch = (None, None, None, None, colors, None, None)
        out[f"blur_{s}"] = data_module.ImagePoseDataModule.get_blurred_pixel_colors(fake, batch, s)[4]
    arrays = {k: (v.detach().numpy() if isinstance(v, th.Tensor) else np.asarray(v)) for k, v in out.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dataset.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
