"""Oracle: ray generation and batch assembly of the reference's dataset
in plain PyTorch on the CPU.  Test infrastructure only.  Pinned by tests/golden/dataset.npz
(outputs of the unmodified reference's ImagePoseDataset static methods and
ImagePoseDataModule.get_blurred_pixel_colors)."""
import torch as th


def directions_meshgrid(height: int, width: int, focal: float) -> th.Tensor:
    """reference barf/dataset.py:407-451 -> (H*W, 3) unit directions, pixel centres, -z forward."""
    y, x = th.meshgrid(-th.linspace(-(height - 1) / 2, (height - 1) / 2, height) / focal,
                       th.linspace(-(width - 1) / 2, (width - 1) / 2, width) / focal, indexing="ij")
    d = th.stack((x, y, -th.ones_like(x)), dim=-1)
    d = d / th.norm(d, p=2, dim=-1, keepdim=True)
    return d.view(-1, 3)


def meshgrid_to_world(meshgrid: th.Tensor, c2w: th.Tensor):
    """reference barf/dataset.py:454-482 -> origins, directions (N, H*W, 3)."""
    return (c2w[:, :3, 3].unsqueeze(1).repeat(1, meshgrid.shape[0], 1),
            th.matmul(c2w[:, :3, :3].unsqueeze(1), meshgrid.unsqueeze(0).unsqueeze(-1)).squeeze(-1))


def get_items(index: th.Tensor, images, c2w_raw, c2w_noisy, focal: float, id_map=None):
    """Batched reference ImagePoseDataset.__getitem__ (barf/dataset.py:613-637).
    images: (N, H, W, n_sigmas, 3)."""
    N, H, W, n_sig, _ = images.shape
    grid = directions_meshgrid(H, W, focal)
    o_r, d_r = meshgrid_to_world(grid, c2w_raw)
    o_n, d_n = meshgrid_to_world(grid, c2w_noisy)
    img = index // (H * W)
    pix = index % (H * W)
    ids = img if id_map is None else th.as_tensor(id_map)[img]
    return (o_r[img, pix], o_n[img, pix], d_r[img, pix], d_n[img, pix],
            images.view(N, H * W, n_sig, 3)[img, pix], ids, th.full((index.numel(),), 1.0 / focal))


def blurred_pixel_colors(colors: th.Tensor, sigmas, sigma: float) -> th.Tensor:
    """reference barf/data_module.py:326-358 -> (B, 2, 3) = [blurred, original]."""
    if sigma <= 0.25:
        return th.stack([colors[:, -1], colors[:, -1]], dim=1)
    if sigma >= max(sigmas):
        return th.stack([colors[:, 0], colors[:, -1]], dim=1)
    index_low = index_high = 0
    for index_high, s in enumerate(sigmas):
        if s < sigma:
            break
        index_low = index_high
    coef = (sigma - sigmas[index_high]) / (sigmas[index_low] - sigmas[index_high] + 1e-8)
    return th.stack([colors[:, index_low] * coef + colors[:, index_high] * (1 - coef), colors[:, -1]], dim=1)
