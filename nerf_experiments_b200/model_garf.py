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
    def forward(ctx, sigma, rgb, delta, t_mid):
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
    """trans = exp(-exclusive cumsum(sigma * delta)); cdf = 1 - [trans, 0]  (B, S+1) — one kernel
    (`ops.transmittance` without, `ops.transmittance_cdf` with gradient to sigma)."""
    if th.is_grad_enabled() and sigma.requires_grad:
        cdf = ops.transmittance_cdf(sigma, t_start, t_end)
        return 1.0 - cdf[:, :-1], cdf
    return ops.transmittance(sigma, t_start, t_end)


def pdf_outer_loss(t_query: th.Tensor, cdf_query: th.Tensor, t_key: th.Tensor, cdf_key: th.Tensor,
                   eps: float = 1e-7) -> th.Tensor:
    """nerfacc's proposal loss (Mip-NeRF 360 eq. 13), mean over the query bins: the key (proposal)
    histogram must bound the query (radiance) histogram from above; only the excess is penalised.
    One kernel for the loss and its gradient w.r.t. the key cdf (`ops.proposal_loss`)."""
    return ops.proposal_loss(t_query, cdf_query, t_key, cdf_key, eps)


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

    # -- sampling ----------------------------------------------------------------------------
    def _get_positions(self, ray_origs, ray_dirs, t_starts, t_ends):
        """(B,S,3) sample positions. Kept for callers of the reference surface (garf/ray_logger.py:161-189);
        the hot path never materialises them (the fused kernels form o + (t0 + t1) / 2 d in registers)."""
        return ray_origs[:, None] + ray_dirs[:, None] * ((t_starts + t_ends))[..., None] / 2

    def _s_to_t(self, s: th.Tensor) -> th.Tensor:
        """nerfacc "lindisp" spacing."""
        return 1.0 / (s / self.far_plane + (1.0 - s) / self.near_plane)

    def _unit_edges(self, B: int, dev):
        cache = getattr(self, "_edge_cache", None)
        if cache is None or cache.shape[0] != B or cache.device != dev:
            cache = th.tensor([0.0, 1.0], device=dev).repeat(B, 1).contiguous()
            self._edge_cache = cache
        return cache

    def _sampling(self, ray_origs, ray_dirs, u_rays):
        """PropNetEstimator.sampling with one proposal level (garf/model_garf.py:210-220): every stage is
        one launch — inverse-CDF sampling, lindisp intervals, fused proposal network, transmittance ->
        cdf. Returns the radiance intervals (t_start, t_end, delta, t_mid), their edges and the proposal
        level's (edges, cdf) — handed on through `extras`, never kept on the module: a tensor with an
        autograd graph that outlives its step would keep that step's gradient accumulators alive."""
        B = ray_origs.shape[0]
        dev = ray_origs.device
        stratified = self.training
        if stratified and u_rays is None:
            u_rays = (th.rand(B, device=dev), th.rand(B, device=dev))
        u_prop, u_rad = u_rays if stratified else (None, None)
        s_edges = self._unit_edges(B, dev)
        # proposal level
        s_edges = ops.resample_icdf(s_edges, s_edges, self.proposal_samples_per_ray, u_prop)
        t, t0, t1 = ops.lindisp_intervals(s_edges, self.near_plane, self.far_plane)
        sigma = self.proposal_network.forward_rays(ray_origs, ray_dirs, t0, t1)
        _, cdf = transmittance_cdf(sigma, t0, t1)
        prop = (t, cdf)
        # radiance level
        s_edges = ops.resample_icdf(s_edges, cdf.detach(), self.radiance_samples_per_ray, u_rad)
        t, t0, t1, delta, t_mid = ops.lindisp_intervals(s_edges, self.near_plane, self.far_plane, want_mid=True)
        return t0, t1, delta, t_mid, t, prop

    # -- forward -----------------------------------------------------------------------------
    def forward(self, ray_origs: th.Tensor, ray_dirs: th.Tensor, u_rays=None):
        t_starts, t_ends, delta, t_mid, t_edges, prop = self._sampling(ray_origs, ray_dirs, u_rays)
        rgb_s, sigma = self.radiance_network.forward_rays(ray_origs, ray_dirs, t_starts, t_ends)
        rgb, weights, opacity, depth = _CompositeNerfacc.apply(sigma, rgb_s, delta, t_mid)
        trans, cdf_q = ops.transmittance(sigma, t_starts, t_ends)
        extras = {"weights": weights, "trans": trans, "t_starts": t_starts, "t_ends": t_ends,
                  "rgbs": rgb_s, "sigmas": sigma, "cdf": cdf_q, "t_edges": t_edges,
                  "prop_t_edges": prop[0], "prop_cdf": prop[1]}
        return rgb, opacity[:, None], depth[:, None], extras

    def compute_proposal_loss(self, extras: Dict) -> th.Tensor:
        """PropNetEstimator.compute_loss(extras["trans"]) (garf/model_garf.py:257)."""
        cdf_q = extras.get("cdf")
        if cdf_q is None:
            trans = extras["trans"].detach()
            cdf_q = 1.0 - th.cat((trans, th.zeros_like(trans[:, :1])), dim=1)
        t_q = extras.get("t_edges")
        if t_q is None:
            t_q = th.cat((extras["t_starts"], extras["t_ends"][:, -1:]), dim=1)
        return pdf_outer_loss(t_q, cdf_q, extras["prop_t_edges"], extras["prop_cdf"])

    def _forward_loss(self, batch: InnerModelBatchInput, u_rays=None):
        ray_origs, ray_dirs, ray_colors = batch
        ray_colors_pred, _, _, extras = self(ray_origs, ray_dirs, u_rays)
        proposal_loss = self.compute_proposal_loss(extras)
        radiance_loss = nn.functional.mse_loss(ray_colors_pred, ray_colors)
        return ray_colors_pred, (proposal_loss, radiance_loss)

    # -- engine surface (engine.TrainEngine): one flat buffer, fused Adam + ExponentialLR -----------------
    def fused_networks(self):
        return [self.proposal_network, self.radiance_network]

    @property
    def param_groups(self):
        """The four groups of the reference's two Adam optimisers (garf/model_garf.py:365-428) as ONE
        list for the fused optimiser: arithmetic per parameter is identical (same betas / eps)."""
        if getattr(self, "_param_groups", None) is None:
            groups = []
            for net, lr0, lr1, n, wd in (
                    (self.proposal_network, self.proposal_learning_rate_start, self.proposal_learning_rate_stop,
                     self.proposal_learning_rate_decay_end, self.proposal_weight_decay),
                    (self.radiance_network, self.radiance_learning_rate_start, self.radiance_learning_rate_stop,
                     self.radiance_learning_rate_decay_end, self.radiance_weight_decay)):
                gamma = self._calculate_decay_factor(lr0, lr1, n)
                for params, factor in ((net.parameters_linear(), 1.0), (net.parameters_gaussian(), self.gaussian_learning_rate_factor)):
                    groups.append({"parameters": list(params), "learning_rate_start": factor * lr0,
                                   "learning_rate_stop": factor * lr1, "learning_rate_decay_end": n,
                                   "weight_decay": wd, "schedule": "exponential", "gamma": gamma})
            self._param_groups = groups
        return self._param_groups

    def training_loss(self, ray_origs, ray_dirs, ray_colors, u_prop=None, u_rad=None):
        """Device part of training_step (no host synchronisation; capturable): the summed loss the
        reference back-propagates (garf/model_garf.py:311) and the values it logs."""
        u = (u_prop, u_rad) if u_prop is not None else None
        _, (proposal_loss, radiance_loss) = self._forward_loss((ray_origs, ray_dirs, ray_colors), u)
        logs = {"loss_fine": radiance_loss.detach(), "train_proposal_loss": proposal_loss.detach(),
                "train_psnr": -10 * th.log10(radiance_loss.detach())}
        return radiance_loss + proposal_loss, logs

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


def garf_engine(model: GarfModel, device, process_group=None):
    """engine.TrainEngine for a GarfModel: both networks in one flat parameter buffer, the whole step
    (sampling, both fused networks, compositing, proposal loss, backward, all-reduce, fused Adam with the
    four ExponentialLR groups) without host synchronisation. torch.optim.Adam's default eps (1e-8), as
    the reference constructs its optimisers (garf/model_garf.py:367-409)."""
    from .engine import TrainEngine
    return TrainEngine(model, device, process_group=process_group, eps=1e-8, loss_fn=model.training_loss)
