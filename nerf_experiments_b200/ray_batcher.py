"""GPU-resident ray batcher and full-image renderer.

`GpuRayBatcher` keeps what the reference's `ImagePoseDataset` keeps (images with their blur
pyramid, raw and noisy camera-to-world matrices, focal length) on the device and produces the
7-tuple the reference's `_step_helper` consumes (barf/model_interpolation.py:490-500) for any
tensor of flat ray indices in one kernel launch — instead of one Python `__getitem__` per ray,
a collate and a host-to-device copy (barf/dataset.py:613-637, barf/data_module.py:202-209).
The blur-pyramid interpolation of `ImagePoseDataModule.get_blurred_pixel_colors`
(barf/data_module.py:276-369) is fused into the same launch.

`render_image` is the chunked full-image render of `Log2dImageReconstruction`
(barf/image_logger.py:155-214) on top of it."""
from typing import Optional, Sequence

import torch as th

from ._lib import check, lib


class GpuRayBatcher:
    def __init__(self, images: th.Tensor, camera_to_worlds: th.Tensor, focal_length: float,
                 camera_to_worlds_noisy: Optional[th.Tensor] = None,
                 gaussian_blur_sigmas: Sequence[float] = (0.0,),
                 index_to_index: Optional[Sequence[int]] = None, device="cuda"):
        """images: (N, H, W, n_sigmas, 3) fp32 in [0,1] (blur levels ordered as the reference's
        `gaussian_blur_sigmas`: decreasing sigma, the last level is the unblurred image);
        camera_to_worlds: (N, 4, 4)."""
        dev = th.device(device)
        if images.dim() != 5 or images.shape[-1] != 3:
            raise ValueError(f"images must be (N, H, W, n_sigmas, 3), got {tuple(images.shape)}")
        self.n_images, self.image_height, self.image_width, self.n_sigmas, _ = images.shape
        if len(gaussian_blur_sigmas) != self.n_sigmas:
            raise ValueError("one image level per gaussian_blur_sigma is required")
        self.images = images.to(dev, th.float32).contiguous()
        self.camera_to_worlds = camera_to_worlds.to(dev, th.float32).contiguous()
        noisy = camera_to_worlds if camera_to_worlds_noisy is None else camera_to_worlds_noisy
        self.camera_to_worlds_noisy = noisy.to(dev, th.float32).contiguous()
        self.camera_origins = self.camera_to_worlds[:, :3, 3]
        self.camera_origins_noisy = self.camera_to_worlds_noisy[:, :3, 3]
        self.focal_length = float(focal_length)
        self.pixel_width = 1.0 / self.focal_length            # dataset.py:99
        self.image_batch_size = self.image_height * self.image_width
        self.gaussian_blur_sigmas = list(gaussian_blur_sigmas)
        self.index_to_index = list(range(self.n_images)) if index_to_index is None else list(index_to_index)
        self._id_map = th.tensor(self.index_to_index, dtype=th.int32, device=dev)
        self.device = dev

    def __len__(self) -> int:
        return self.n_images * self.image_batch_size

    # -- blur schedule (host logic of data_module.py:326-358) ---------------------------------------
    def blur_levels(self, sigma: Optional[float]):
        """(index_low, index_high, coefficient) of get_blurred_pixel_colors, or (-1, -1, 0) for the
        raw pyramid."""
        if sigma is None:
            return -1, -1, 0.0
        sig = self.gaussian_blur_sigmas
        if sigma <= 0.25:
            return self.n_sigmas - 1, self.n_sigmas - 1, 1.0
        if sigma >= max(sig):
            return 0, 0, 1.0
        index_low = index_high = 0
        for index_high, s in enumerate(sig):
            if s < sigma:
                break
            index_low = index_high
        coef = (sigma - sig[index_high]) / (sig[index_low] - sig[index_high] + 1e-8)
        return index_low, index_high, float(coef)

    # -- batches ---------------------------------------------------------------------------------
    def batch(self, ray_index: th.Tensor, sigma: Optional[float] = None):
        """(o_raw, o_noisy, d_raw, d_noisy, colors, img_idx, pixel_width) for flat ray indices (B,).
        colors: (B, n_sigmas, 3) if sigma is None, else (B, 2, 3) = [blurred, original]."""
        if not ray_index.is_cuda:
            raise RuntimeError("ray_index: expected a CUDA tensor (nerfb200 has no CPU fallback)")
        idx = ray_index.to(th.int64).contiguous()
        B = idx.numel()
        dev = self.device
        lo, hi, coef = self.blur_levels(sigma)
        o_r = th.empty((B, 3), device=dev)
        o_n = th.empty((B, 3), device=dev)
        d_r = th.empty((B, 3), device=dev)
        d_n = th.empty((B, 3), device=dev)
        colors = th.empty((B, self.n_sigmas if lo < 0 else 2, 3), device=dev)
        img_idx = th.empty((B,), device=dev, dtype=th.int64)
        pw = th.empty((B,), device=dev)
        with th.cuda.device(dev):
            check(lib().nerfb200_ray_batch(
                idx.data_ptr(), B, self.camera_to_worlds.data_ptr(), self.camera_to_worlds_noisy.data_ptr(),
                self.images.data_ptr(), self._id_map.data_ptr(), self.n_images, self.image_height,
                self.image_width, self.n_sigmas, self.focal_length, self.pixel_width, lo, hi, coef,
                o_r.data_ptr(), o_n.data_ptr(), d_r.data_ptr(), d_n.data_ptr(), colors.data_ptr(),
                img_idx.data_ptr(), pw.data_ptr(), th.cuda.current_stream().cuda_stream), "ray_batch")
        return o_r, o_n, d_r, d_n, colors, img_idx, pw

    # -- datamodule / dataset surface the calibration models read ----------------------------------
    @property
    def dataset_train(self):
        return self

    def get_blurred_pixel_colors(self, batch, sigma: float):
        """ImagePoseDataModule.get_blurred_pixel_colors on an assembled batch whose colours are the
        raw pyramid (B, n_sigmas, 3): returns the batch with colours (B, 2, 3) = [blurred, original].
        (`batch(idx, sigma)` fuses this into the gather; this form serves callers that already hold
        a batch.)"""
        o_r, o_p, d_r, d_p, colors, img_idx, pw = batch
        if colors.shape[1] == 2 and self.n_sigmas != 2:
            return batch                                    # already interpolated by batch(idx, sigma)
        lo, hi, coef = self.blur_levels(sigma)
        blurred = colors[:, lo] * coef + colors[:, hi] * (1 - coef) if lo != hi else colors[:, lo]
        return o_r, o_p, d_r, d_p, th.stack([blurred, colors[:, -1]], dim=1), img_idx, pw

    def __getitem__(self, index: int):
        """One ray, the reference's DatasetOutput (dataset.py:613-637)."""
        o_r, o_n, d_r, d_n, c, i, pw = self.batch(th.tensor([index], device=self.device))
        return o_r[0], o_n[0], d_r[0], d_n[0], c[0], i[0], pw[0]

    def epoch_permutation(self, generator: Optional[th.Generator] = None) -> th.Tensor:
        """A shuffled epoch of ray indices on the device (what DataLoader(shuffle=True) draws)."""
        return th.randperm(len(self), device=self.device, generator=generator)

    def image_rays(self, image: int, noisy: bool = False):
        """All H*W rays of one image, row-major (origins, directions)."""
        first = image * self.image_batch_size
        idx = th.arange(first, first + self.image_batch_size, device=self.device)
        o_r, o_n, d_r, d_n, _, _, _ = self.batch(idx)
        return (o_n, d_n) if noisy else (o_r, d_r)


@th.no_grad()
def render_image(model, origins: th.Tensor, directions: th.Tensor, height: int, width: int,
                 pixel_width: float, chunk: int = 16384) -> th.Tensor:
    """(H, W, 3) image: `model.forward(o, d, pixel_width)[0]` over the rays in chunks
    (reference barf/image_logger.py:179-196)."""
    out = []
    for s in range(0, origins.shape[0], chunk):
        o, d = origins[s:s + chunk], directions[s:s + chunk]
        pw = th.full((o.shape[0], 1), pixel_width, device=o.device)
        out.append(model.forward(o, d, pw)[0])
    return th.cat(out).view(height, width, 3).clamp(0, 1)


@th.no_grad()
def render_image_sharded(model, origins: th.Tensor, directions: th.Tensor, height: int, width: int,
                         pixel_width: float, chunk: int = 16384, group=None, dst: int = 0):
    """`render_image` with the image rows split over the ranks of a process group (every rank holds
    the model; rank `dst` receives the (H, W, 3) image, the others None) — config 5 of BASELINE.json:
    the 800 x 800 render sharded over 1 / 2 / 4 / 8 GPUs."""
    from .parallel import render_rows_sharded

    def rows(r0: int, r1: int) -> th.Tensor:
        o, d = origins[r0 * width: r1 * width], directions[r0 * width: r1 * width]
        return render_image(model, o, d, r1 - r0, width, pixel_width, chunk)

    return render_rows_sharded(rows, height, width, origins.device, group, dst)
