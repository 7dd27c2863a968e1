"""GARF model — module surface of reference garf/model_garf.py:18-428 (== barf/model_garf.py):
`forward(ray_origs, ray_dirs) -> (rgb, opacity, depth, extras)`, `_forward_loss`,
`training_step` / `validation_step` with manual optimisation of the two Adam optimisers and
their ExponentialLR schedules.

The reference delegates sampling, compositing and the proposal loss to the third-party package
nerfacc (PropNetEstimator.sampling / rendering / compute_loss, garf/model_garf.py:210-230,257),
which is not part of the reference tree. Here the same chain runs on this repo's kernels:
inverse-CDF resampling (`ops.resample_icdf`), nerfacc-flavour compositing with opacity and depth
(`composite_fwd/bwd`, flavour NERFACC) and the Gaussian activations; the glue (cdf from
transmittance, searchsorted bounds of the proposal loss) is a handful of torch ops. nerfacc's
internal Philox jitter cannot be reproduced, so the per-ray uniforms are an explicit, optional
input (`u_rays`); see oracle/ref_nerfacc.py and oracle/ref_garf.py (parity unpinned)."""
from math import log2
from typing import Dict, Literal, Optional, Tuple

import torch as th
import torch.nn as nn

from . import _lib, ops
from ._lightning_compat import LightningModule
from .model_garf_proposal import ProposalNetwork
from .model_garf_radiance import RadianceNetwork

InnerModelBatchInput = Tuple[th.Tensor, th.Tensor, th.Tensor]


class _CompositeNerfacc(th.autograd.Function):
    """nerfacc.rendering arithmetic on dense (B, S) samples: rgb, weights, opacity, depth."""

    @staticmethod
    def forward(ctx, sigma, rgb, t_start, t_end):
        delta = (t_end - t_start).contiguous()
        t_mid = ((t_start + t_end) * 0.5).contiguous()
        out_rgb, w, opacity, depth = ops.composite_fwd(sigma, delta, rgb, t_mid, _lib.COMPOSITE_NERFACC,
                                                       want_w=True, want_opacity=True, want_depth=True)
        ctx.save_for_backward(sigma.detach(), rgb.detach(), delta, t_mid)
        return out_rgb, w, opacity, depth

    @staticmethod
    def backward(ctx, g_rgb, g_w, g_opacity, g_depth):
        sigma, rgb, delta, t_mid = ctx.saved_tensors
        g_rgb = th.zeros_like(rgb[:, 0, :]) if g_rgb is None else g_rgb.contiguous()
        d_sigma, d_rgb = ops.composite_bwd(sigma, delta, rgb, g_rgb,
                                           None if g_w is None else g_w.contiguous(), t_mid,
                                           None if g_opacity is None else g_opacity.contiguous(),
                                           None if g_depth is None else g_depth.contiguous(),
                                           _lib.COMPOSITE_NERFACC)
        return d_sigma, d_rgb, None, None


def transmittance_cdf(sigma: th.Tensor, t_start: th.Tensor, t_end: th.Tensor):
    """trans = exp(-exclusive cumsum(sigma * delta)); cdf = 1 - [trans, 0]  (B, S+1)."""
    sd = sigma * (t_end - t_start)
    trans = th.exp(-(th.cumsum(sd, dim=1) - sd))
    cdf = 1.0 - th.cat((trans, th.zeros_like(trans[:, :1])), dim=1)
    return trans, cdf


def pdf_outer_loss(t_query: th.Tensor, cdf_query: th.Tensor, t_key: th.Tensor, cdf_key: th.Tensor,
                   eps: float = 1e-7) -> th.Tensor:
    """nerfacc's proposal loss (Mip-NeRF 360 eq. 13): the key (proposal) histogram must bound the
    query (radiance) histogram from above; only the excess is penalised."""
    ids_right = th.searchsorted(t_key.contiguous(), t_query.contiguous(), right=False).clamp(0, t_key.shape[1] - 1)
    ids_left = (th.searchsorted(t_key.contiguous(), t_query.contiguous(), right=True) - 1).clamp(0, t_key.shape[1] - 1)
    w = cdf_query[:, 1:] - cdf_query[:, :-1]
    w_outer = cdf_key.gather(1, ids_right[:, 1:]) - cdf_key.gather(1, ids_left[:, :-1])
    return th.clip(w - w_outer, min=0) ** 2 / (w + eps)


class GarfModel(LightningModule):
    def __init__(self, near_plane: float, far_plane: float, proposal_samples_per_ray: int,
                 radiance_samples_per_ray: int, gaussian_init_min: float, gaussian_init_max: float,
                 gaussian_learning_rate_factor: float, proposal_learning_rate_start: float,
                 proposal_learning_rate_stop: float, proposal_learning_rate_decay_end: int,
                 proposal_weight_decay: float, radiance_learning_rate_start: float,
                 radiance_learning_rate_stop: float, radiance_learning_rate_decay_end: int,
                 radiance_weight_decay: float):
        super().__init__()
        self.save_hyperparameters()
        self.near_plane = near_plane
        self.far_plane = far_plane
        self.proposal_samples_per_ray = proposal_samples_per_ray
        self.radiance_samples_per_ray = radiance_samples_per_ray
        self.gaussian_init_min = gaussian_init_min
        self.gaussian_init_max = gaussian_init_max
        self.gaussian_learning_rate_factor = gaussian_learning_rate_factor
        self.proposal_learning_rate_start = proposal_learning_rate_start
        self.proposal_learning_rate_stop = proposal_learning_rate_stop
        self.proposal_learning_rate_decay_end = proposal_learning_rate_decay_end
        self.proposal_weight_decay = proposal_weight_decay
        self.radiance_learning_rate_start = radiance_learning_rate_start
        self.radiance_learning_rate_stop = radiance_learning_rate_stop
        self.radiance_learning_rate_decay_end = radiance_learning_rate_decay_end
        self.radiance_weight_decay = radiance_weight_decay
        # creation order = the reference's (proposal first): it fixes the seeded initial values
        self.proposal_network = ProposalNetwork(gaussian_init_min=gaussian_init_min,
                                                gaussian_init_max=gaussian_init_max)
        self.radiance_network = RadianceNetwork(gaussian_init_min=gaussian_init_min,
                                                gaussian_init_max=gaussian_init_max)
        self.automatic_optimization = False
        self._prop_cache = None     # (t edges, cdf) of the proposal level of the last forward

    # -- sampling ----------------------------------------------------------------------------
    def _get_positions(self, ray_origs, ray_dirs, t_starts, t_ends):
        return ray_origs[:, None] + ray_dirs[:, None] * ((t_starts + t_ends))[..., None] / 2

    def _s_to_t(self, s: th.Tensor) -> th.Tensor:
        """nerfacc "lindisp" spacing."""
        return 1.0 / (s / self.far_plane + (1.0 - s) / self.near_plane)

    def _sampling(self, ray_origs, ray_dirs, u_rays):
        """PropNetEstimator.sampling with one proposal level (garf/model_garf.py:210-220)."""
        B = ray_origs.shape[0]
        dev = ray_origs.device
        stratified = self.training
        if stratified and u_rays is None:
            u_rays = (th.rand(B, device=dev), th.rand(B, device=dev))
        u_prop, u_rad = u_rays if stratified else (None, None)
        s_edges = th.tensor([0.0, 1.0], device=dev).expand(B, 2).contiguous()
        cdf = s_edges.clone()
        # proposal level
        s_edges = ops.resample_icdf(s_edges, cdf, self.proposal_samples_per_ray, u_prop)
        t = self._s_to_t(s_edges)
        t0, t1 = t[:, :-1].contiguous(), t[:, 1:].contiguous()
        # the closure of garf/model_garf.py:127-141 (positions, network) as one fused launch on the rays
        sigma = self.proposal_network.forward_rays(ray_origs, ray_dirs, t0, t1)
        _, cdf = transmittance_cdf(sigma, t0, t1)
        self._prop_cache = (t, cdf)
        # radiance level
        s_edges = ops.resample_icdf(s_edges, cdf.detach(), self.radiance_samples_per_ray, u_rad)
        t = self._s_to_t(s_edges)
        return t[:, :-1].contiguous(), t[:, 1:].contiguous()

    # -- forward -----------------------------------------------------------------------------
    def forward(self, ray_origs: th.Tensor, ray_dirs: th.Tensor, u_rays=None):
        t_starts, t_ends = self._sampling(ray_origs, ray_dirs, u_rays)
        rgb_s, sigma = self.radiance_network.forward_rays(ray_origs, ray_dirs, t_starts, t_ends)
        rgb, weights, opacity, depth = _CompositeNerfacc.apply(sigma, rgb_s, t_starts, t_ends)
        with th.no_grad():
            trans, _ = transmittance_cdf(sigma, t_starts, t_ends)
        extras = {"weights": weights, "trans": trans, "t_starts": t_starts, "t_ends": t_ends,
                  "rgbs": rgb_s, "sigmas": sigma}
        return rgb, opacity[:, None], depth[:, None], extras

    def compute_proposal_loss(self, extras: Dict) -> th.Tensor:
        """PropNetEstimator.compute_loss(extras["trans"]) (garf/model_garf.py:257)."""
        trans = extras["trans"].detach()
        cdf_q = 1.0 - th.cat((trans, th.zeros_like(trans[:, :1])), dim=1)
        t_q = th.cat((extras["t_starts"], extras["t_ends"][:, -1:]), dim=1)
        t_k, cdf_k = self._prop_cache
        return pdf_outer_loss(t_q, cdf_q, t_k, cdf_k).mean()

    def _forward_loss(self, batch: InnerModelBatchInput, u_rays=None):
        ray_origs, ray_dirs, ray_colors = batch
        ray_colors_pred, _, _, extras = self(ray_origs, ray_dirs, u_rays)
        proposal_loss = self.compute_proposal_loss(extras)
        radiance_loss = nn.functional.mse_loss(ray_colors_pred, ray_colors)
        return ray_colors_pred, (proposal_loss, radiance_loss)

    def _get_logging_losses(self, stage: Literal["train", "val", "test"], batch_idx: int,
                            proposal_loss: th.Tensor, radiance_loss: th.Tensor, *args, **kwargs):
        psnr = -10 * th.log10(radiance_loss)
        return {f"{stage}_proposal_loss": proposal_loss, f"{stage}_radiance_loss": radiance_loss,
                f"{stage}_psnr": psnr}

    # -- steps -------------------------------------------------------------------------------
    def _opt_and_sched(self):
        if getattr(self, "trainer", None) is not None:      # real Lightning
            return self.optimizers(use_pl_optimizer=False), self.lr_schedulers()
        if not hasattr(self, "_proposal_optimizer"):
            self.configure_optimizers()
        return ([self._proposal_optimizer, self._radiance_optimizer],
                [self._proposal_learning_rate_scheduler, self._radiance_learning_rate_scheduler])

    def training_step(self, batch: InnerModelBatchInput, batch_idx: int, u_rays=None):
        _, (proposal_loss, radiance_loss) = self._forward_loss(batch, u_rays)
        optimizers, schedulers = self._opt_and_sched()
        for optimizer in optimizers:
            optimizer.zero_grad()
        (radiance_loss + proposal_loss).backward()
        for optimizer, scheduler in zip(optimizers, schedulers):
            optimizer.step()
            scheduler.step()
        self.log_dict(self._get_logging_losses("train", batch_idx, proposal_loss, radiance_loss))
        return radiance_loss + proposal_loss

    def validation_step(self, batch: InnerModelBatchInput, batch_idx: int):
        _, (proposal_loss, radiance_loss) = self._forward_loss(batch)
        self.log_dict(self._get_logging_losses("val", batch_idx, proposal_loss, radiance_loss))
        return radiance_loss + proposal_loss

    def _calculate_decay_factor(self, learning_rate_start: float, learning_rate_stop: float,
                                learning_rate_decay_end: int) -> float:
        return 2 ** (log2(learning_rate_stop / learning_rate_start) / learning_rate_decay_end)

    def _adam(self, net, lr: float, weight_decay: float):
        g_lr = self.gaussian_learning_rate_factor * lr
        return th.optim.Adam([
            {"params": net.parameters_linear(), "lr": lr, "initial_lr": lr, "weight_decay": weight_decay},
            {"params": net.parameters_gaussian(), "lr": g_lr, "initial_lr": g_lr, "weight_decay": weight_decay}])

    def configure_optimizers(self):
        self._proposal_optimizer = self._adam(self.proposal_network, self.proposal_learning_rate_start,
                                              self.proposal_weight_decay)
        self._proposal_learning_rate_scheduler = th.optim.lr_scheduler.ExponentialLR(
            self._proposal_optimizer,
            gamma=self._calculate_decay_factor(self.proposal_learning_rate_start, self.proposal_learning_rate_stop,
                                               self.proposal_learning_rate_decay_end),
            last_epoch=-self.proposal_learning_rate_decay_end - 1)
        self._radiance_optimizer = self._adam(self.radiance_network, self.radiance_learning_rate_start,
                                              self.radiance_weight_decay)
        self._radiance_learning_rate_scheduler = th.optim.lr_scheduler.ExponentialLR(
            self._radiance_optimizer,
            gamma=self._calculate_decay_factor(self.radiance_learning_rate_start, self.radiance_learning_rate_stop,
                                               self.radiance_learning_rate_decay_end),
            last_epoch=-self.radiance_learning_rate_decay_end - 1)
        return ([self._proposal_optimizer, self._radiance_optimizer],
                [self._proposal_learning_rate_scheduler, self._radiance_learning_rate_scheduler])
