"""Procedural scene for benchmarks and end-to-end tests: analytic coloured SDF primitives
rendered to posed images (no dataset is available offline).  Camera model and ray convention
follow the reference's Blender-format loader: pixel centres, camera looks along -z, unit-norm
directions, focal = W / 2 / tan(camera_angle_x / 2), pixel_width = 1 / focal
(reference barf/dataset.py:99,303,443-451,479-482; barf/visualise_mip_barf_pe_mask.py:59)."""
import math
from dataclasses import dataclass

import torch as th

CAMERA_ANGLE_X = 0.6911112070083618


def _sdf_and_color(p: th.Tensor):
    """p (N,3) -> signed distance (N,), albedo (N,3).  Two spheres, a box and a torus inside the
    radius-1.5 ball."""
    d_s1 = (p - p.new_tensor([0.45, 0.0, 0.15])).norm(dim=1) - 0.55
    d_s2 = (p - p.new_tensor([-0.6, 0.35, -0.2])).norm(dim=1) - 0.35
    q = (p - p.new_tensor([-0.1, -0.55, 0.0])).abs() - p.new_tensor([0.35, 0.25, 0.45])
    d_box = q.clamp_min(0).norm(dim=1) + q.max(dim=1).values.clamp_max(0)
    pt = p - p.new_tensor([0.0, 0.1, -0.65])
    d_tor = th.stack((pt[:, [0, 1]].norm(dim=1) - 0.5, pt[:, 2]), dim=1).norm(dim=1) - 0.14
    ds = th.stack((d_s1, d_s2, d_box, d_tor), dim=1)
    colors = p.new_tensor([[0.9, 0.2, 0.2], [0.2, 0.7, 0.3], [0.2, 0.3, 0.9], [0.9, 0.8, 0.2]])
    d, idx = ds.min(dim=1)
    return d, colors[idx]


def _normal(p: th.Tensor, eps: float = 1e-3):
    e = th.eye(3, device=p.device) * eps
    g = th.stack([_sdf_and_color(p + e[i])[0] - _sdf_and_color(p - e[i])[0] for i in range(3)], dim=1)
    return th.nn.functional.normalize(g, dim=1)


@th.no_grad()
def shade_rays(o: th.Tensor, d: th.Tensor, near: float = 2.0, far: float = 8.0, steps: int = 96):
    """Sphere-traces rays (N,3)/(N,3) and returns rgb (N,3) with a white background."""
    t = th.full((o.shape[0],), near, device=o.device)
    hit = th.zeros_like(t, dtype=th.bool)
    for _ in range(steps):
        dist, _ = _sdf_and_color(o + t[:, None] * d)
        hit |= dist < 1e-3
        t = th.where(hit | (t > far), t, t + dist.clamp_min(1e-3))
    p = o + t[:, None] * d
    _, albedo = _sdf_and_color(p)
    n = _normal(p)
    light = th.nn.functional.normalize(o.new_tensor([0.5, 0.8, 0.6]), dim=0)
    diff = (n @ light).clamp_min(0.0)[:, None]
    rgb = albedo * (0.35 + 0.65 * diff)
    return th.where((hit & (t <= far))[:, None], rgb, th.ones_like(rgb))


def look_at_poses(n: int, radius: float, generator: th.Generator):
    """n camera-to-world matrices (n,4,4) on a sphere of `radius`, looking at the origin."""
    u = th.rand(n, generator=generator)
    v = th.rand(n, generator=generator)
    theta = 2 * math.pi * u
    phi = th.acos(1 - 1.2 * v)          # upper part of the sphere, like the Blender scenes
    c = th.stack((th.sin(phi) * th.cos(theta), th.sin(phi) * th.sin(theta), th.cos(phi)), dim=1) * radius
    z = th.nn.functional.normalize(c, dim=1)            # camera looks along -z => z axis points away
    up = th.tensor([0.0, 0.0, 1.0]).expand_as(z)
    x = th.nn.functional.normalize(th.cross(up, z, dim=1), dim=1)
    y = th.cross(z, x, dim=1)
    c2w = th.eye(4).repeat(n, 1, 1)
    c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = x, y, z, c
    return c2w


def camera_rays(c2w: th.Tensor, height: int, width: int, focal: float):
    """(H*W,3) origins and unit directions of one camera (pixel centres, -z forward)."""
    j, i = th.meshgrid(th.arange(height, device=c2w.device), th.arange(width, device=c2w.device), indexing="ij")
    dirs = th.stack(((i + 0.5 - width / 2) / focal, -(j + 0.5 - height / 2) / focal,
                     -th.ones_like(i, dtype=th.float32)), dim=-1).reshape(-1, 3).float()
    dirs = th.nn.functional.normalize(dirs, dim=1)
    d = dirs @ c2w[:3, :3].T
    o = c2w[:3, 3].expand_as(d)
    return o.contiguous(), d.contiguous()


@dataclass
class SyntheticScene:
    origins: th.Tensor        # (n_img*H*W, 3) ray origins with the (noisy) initial poses
    directions: th.Tensor     # (n_img*H*W, 3)
    origins_true: th.Tensor
    directions_true: th.Tensor
    colors: th.Tensor         # (n_img*H*W, 3)
    image_index: th.Tensor    # (n_img*H*W,) int32
    pixel_width: float
    n_images: int
    height: int
    width: int
    batcher: object = None    # GpuRayBatcher over the images + blur pyramid (make_scene(blur_sigmas=...))

    @property
    def n_rays(self):
        return self.colors.shape[0]

    def batch(self, idx: th.Tensor):
        pw = th.full((idx.shape[0], 1), self.pixel_width, device=idx.device)
        return self.origins[idx], self.directions[idx], self.colors[idx], self.image_index[idx], pw


def so3_exp(w: th.Tensor) -> th.Tensor:
    K = th.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    return th.matrix_exp(K)


@th.no_grad()
def gaussian_blur_pyramid(images: th.Tensor, sigmas) -> th.Tensor:
    """(N, H, W, 3) -> (N, H, W, n_sigmas, 3): one separable Gaussian blur per sigma, sigmas <= 0.25
    kept as the original image (what ImagePoseDataset.gaussian_blur builds with PIL,
    reference barf/dataset.py:251-270; a data-generation helper, not a parity-critical path)."""
    x = images.permute(0, 3, 1, 2).contiguous()
    levels = []
    for sigma in sigmas:
        if sigma <= 0.25:
            levels.append(x)
            continue
        r = max(int(math.ceil(3 * sigma)), 1)
        k = th.exp(-0.5 * (th.arange(-r, r + 1, device=x.device, dtype=th.float32) / sigma) ** 2)
        k = (k / k.sum()).view(1, 1, -1)
        c = x.shape[1]
        y = th.nn.functional.pad(x, (r, r, 0, 0), mode="replicate")
        y = th.nn.functional.conv2d(y, k.view(1, 1, 1, -1).expand(c, 1, 1, -1), groups=c)
        y = th.nn.functional.pad(y, (0, 0, r, r), mode="replicate")
        y = th.nn.functional.conv2d(y, k.view(1, 1, -1, 1).expand(c, 1, -1, 1), groups=c)
        levels.append(y)
    return th.stack(levels, dim=1).permute(0, 3, 4, 1, 2).contiguous()


def make_scene(n_images: int, height: int, width: int, device, seed: int = 134534,
               rotation_noise: float = 0.0, translation_noise: float = 0.0, radius: float = 4.0,
               blur_sigmas=None):
    """Renders the scene from n_images seeded poses; the stored rays use poses perturbed by
    N(0, noise^2) in so(3) and translation (the BARF setting, reference barf/run_barf.py:27-30).
    With `blur_sigmas` (decreasing, last = 0: the reference's `gaussian_blur_sigmas`) the scene also
    carries a `ray_batcher.GpuRayBatcher` over the images and their blur pyramid (`scene.batcher`)."""
    g = th.Generator().manual_seed(seed)
    c2w = look_at_poses(n_images, radius, g)
    focal = width / 2 / math.tan(CAMERA_ANGLE_X / 2)
    rot_n = th.randn((n_images, 3), generator=g) * rotation_noise
    tr_n = th.randn((n_images, 3), generator=g) * translation_noise
    c2w_noisy = c2w.clone()
    c2w_noisy[:, :3, :3] = so3_exp(rot_n) @ c2w[:, :3, :3]
    c2w_noisy[:, :3, 3] = c2w[:, :3, 3] + tr_n
    O, D, On, Dn, Cs, Is = [], [], [], [], [], []
    for k in range(n_images):
        o, d = camera_rays(c2w[k].to(device), height, width, focal)
        on, dn = camera_rays(c2w_noisy[k].to(device), height, width, focal)
        O.append(o); D.append(d); On.append(on); Dn.append(dn)
        Cs.append(shade_rays(o, d))
        Is.append(th.full((o.shape[0],), k, device=device, dtype=th.int32))
    sc = SyntheticScene(origins=th.cat(On), directions=th.cat(Dn), origins_true=th.cat(O),
                        directions_true=th.cat(D), colors=th.cat(Cs), image_index=th.cat(Is),
                        pixel_width=1.0 / focal, n_images=n_images, height=height, width=width)
    sc.c2w, sc.c2w_noisy, sc.focal = c2w, c2w_noisy, focal
    if blur_sigmas is not None:
        from .ray_batcher import GpuRayBatcher
        images = sc.colors.view(n_images, height, width, 3)
        sc.batcher = GpuRayBatcher(gaussian_blur_pyramid(images, list(blur_sigmas)), c2w, focal, c2w_noisy,
                                   list(blur_sigmas), device=device)
    return sc
