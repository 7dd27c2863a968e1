"""GPU ray batcher (through the C ABI) against the reference's dataset outputs
(tests/golden/dataset.npz) and the oracle; full-image render helper."""
import os

import numpy as np
import pytest
import torch as th

from oracle import ref_dataset

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset.npz")


def _g():
    z = np.load(G)
    return {k: (th.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def _batcher(cuda, g, id_map=None):
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher
    return GpuRayBatcher(g["images"], g["c2w"], float(g["focal"]), g["c2w_noisy"],
                         [float(s) for s in g["sigmas"]], id_map, cuda)


def test_batch_matches_reference_dataset(cuda):
    g = _g()
    H, W = int(g["H"]), int(g["W"])
    N = g["images"].shape[0]
    b = _batcher(cuda, g, id_map=[7, 3, 11])
    assert len(b) == N * H * W and b.pixel_width == pytest.approx(1 / float(g["focal"]))
    idx = th.randperm(len(b), generator=th.Generator().manual_seed(0))
    o_r, o_n, d_r, d_n, c, ids, pw = b.batch(idx.to(cuda))
    img, pix = idx // (H * W), idx % (H * W)
    # fp32 within tolerance: the 3x3 rotation is a different summation order than ATen's bmm
    assert th.allclose(d_r.cpu(), g["d_raw"][img, pix], atol=1e-6)
    assert th.allclose(d_n.cpu(), g["d_noisy"][img, pix], atol=1e-6)
    assert th.equal(o_r.cpu(), g["o_raw"][img, pix]) and th.equal(o_n.cpu(), g["o_noisy"][img, pix])
    assert th.equal(c.cpu(), g["images"].view(N, H * W, -1, 3)[img, pix])
    assert th.equal(ids.cpu(), th.tensor([7, 3, 11])[img]) and ids.dtype == th.int64
    assert th.equal(pw.cpu(), th.full((len(b),), 1 / float(g["focal"])))
    # single item = the reference's __getitem__
    item = b[17]
    ref = ref_dataset.get_items(th.tensor([17]), g["images"], g["c2w"], g["c2w_noisy"], float(g["focal"]), [7, 3, 11])
    for a, r in zip(item, ref):
        assert th.allclose(a.cpu().float(), r[0].float(), atol=1e-6)


@pytest.mark.parametrize("sigma", [0.1, 0.25, 1.0, 3.0, 5.5, 8.0])
def test_fused_blur_interpolation(cuda, sigma):
    g = _g()
    b = _batcher(cuda, g)
    idx = th.arange(len(b), device=cuda)
    c = b.batch(idx, sigma=sigma)[4]
    assert c.shape == (len(b), 2, 3)
    assert th.allclose(c.cpu(), g[f"blur_{sigma}"], atol=1e-7)


def test_large_batch_and_image_rays(cuda):
    """400x400 images, one million random rays: directions stay unit length and agree with the
    oracle's stored-direction gather; image_rays() returns the row-major rays of one view."""
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher
    gen = th.Generator().manual_seed(2)
    N, H, W, focal = 4, 400, 400, 555.5
    A = th.randn((N, 3, 3), generator=gen)
    Q, _ = th.linalg.qr(A)
    c2w = th.eye(4).repeat(N, 1, 1)
    c2w[:, :3, :3] = Q
    c2w[:, :3, 3] = th.randn((N, 3), generator=gen) * 4
    images = th.rand((N, H, W, 1, 3), generator=gen)
    b = GpuRayBatcher(images, c2w, focal, None, [0.0], None, cuda)
    idx = th.randint(0, len(b), (1 << 20,), generator=gen)
    o_r, o_n, d_r, d_n, c, ids, pw = b.batch(idx.to(cuda))
    assert th.allclose(d_r.norm(dim=1), th.ones(1 << 20, device=cuda), atol=1e-5)
    ref = ref_dataset.get_items(idx[:5000], images, c2w, c2w, focal)
    assert th.allclose(d_r[:5000].cpu(), ref[2], atol=1e-6) and th.equal(c[:5000].cpu(), ref[4])
    o, d = b.image_rays(2)
    grid = ref_dataset.directions_meshgrid(H, W, focal)
    _, d_ref = ref_dataset.meshgrid_to_world(grid, c2w[2:3])
    assert th.allclose(d.cpu(), d_ref[0], atol=1e-6) and th.equal(o.cpu(), c2w[2, :3, 3].expand(H * W, 3))


def test_render_image_helper(cuda):
    from nerf_experiments_b200 import model_interpolation as mi
    from nerf_experiments_b200 import model_interpolation_architecture as arch
    from nerf_experiments_b200 import positional_encodings as pe
    from nerf_experiments_b200.ray_batcher import GpuRayBatcher, render_image
    th.manual_seed(0)
    ep = pe.BarfPositionalEncoding(10, 10.0, 0.0, 1.0, True, 1.0)
    ed = pe.BarfPositionalEncoding(4, 4.0, 0.0, 1.0, True, 1.0)
    net = arch.NerfModel(2, 64, True, False, 2, ep, ed)
    model = mi.NerfInterpolation(2.0, 8.0, net, 32, "equidistant", 0.0, "middle").to(cuda).eval()
    c2w = th.eye(4).repeat(1, 1, 1)
    c2w[0, 2, 3] = 4.0
    b = GpuRayBatcher(th.zeros((1, 24, 20, 1, 3)), c2w, 30.0, None, [0.0], None, cuda)
    o, d = b.image_rays(0)
    img = render_image(model, o, d, 24, 20, b.pixel_width, chunk=100)
    assert img.shape == (24, 20, 3) and th.isfinite(img).all() and (img >= 0).all() and (img <= 1).all()
    whole = render_image(model, o, d, 24, 20, b.pixel_width, chunk=10 ** 6)
    assert th.allclose(img, whole, atol=1e-6)       # chunking does not change the image
