"""Samplers of the hot path with the reference's signatures.

SpacedSampler / PowerSampler   <- nerfstudio/model_components/ray_samplers.py:55-132,838-852
PDFSampler                     <- nerfstudio/model_components/ray_samplers.py:255-376
ProposalNetworkSampler         <- nerfstudio/model_components/ray_samplers.py:569-666

Random jitter is drawn with torch.rand on the ray device, in the reference's order and shapes, and handed to the
kernels; everything downstream of the draw runs in the warp-per-ray CUDA kernels.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import functional as F
from .rays import RayBundle, RaySamples, ray_data_of


def _make_ray_samples(ray_bundle: RayBundle, rays: F.RayData, sbins: Tensor, ebins: Tensor, spacing, fn) -> RaySamples:
    rs = ray_bundle.get_ray_samples(
        bin_starts=ebins[..., :-1, None],
        bin_ends=ebins[..., 1:, None],
        spacing_starts=sbins[..., :-1, None],
        spacing_ends=sbins[..., 1:, None],
        spacing_to_euclidean_fn=fn,
    )
    rs.ray_data, rs.euclidean_bins, rs.spacing_bins, rs.spacing = rays, ebins, sbins, spacing
    return rs


class Sampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    def forward(self, *args, **kwargs):
        return self.generate_ray_samples(*args, **kwargs)


class PowerSampler(Sampler):
    """ZipNeRF power-transform spacing: spacing_fn(x) = power_fn(x*scaling, lambda_)."""

    def __init__(self, num_samples: Optional[int] = None, lambda_=-1.5, scaling=2.0, train_stratified=True,
                 single_jitter=False) -> None:
        super().__init__(num_samples=num_samples)
        if lambda_ in (0, 1) or abs(lambda_) > 1e10:
            raise NotImplementedError("lambda_ must be finite and not 0 or 1")
        self.lambda_ = float(lambda_)
        self.scaling = float(scaling)
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter

    def _closure(self, nears: Tensor, fars: Tensor) -> Callable:
        lam, scaling = self.lambda_, self.scaling
        lam_1 = abs(lam - 1)

        def spacing_fn(x):
            return (lam_1 / lam) * ((x * scaling / lam_1 + 1) ** lam - 1)

        def spacing_to_euclidean_fn(x):  # utils/math.py:561-580 / ray_samplers.py:117-120, for foreign callers
            # (evaluated on demand: the kernels carry (lambda, scaling) and the per-ray near / far themselves, so on the
            #  path this closure is never called and its ~12 element-wise launches per step are not spent)
            s_near, s_far = spacing_fn(nears), spacing_fn(fars)
            v = x * s_far + (1 - x) * s_near
            return (((v * lam / lam_1 + 1).clamp_min(1e-10) ** (1 / lam) - 1) * lam_1) / scaling

        return spacing_to_euclidean_fn

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None
        assert ray_bundle.nears is not None
        assert ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        rays = ray_data_of(ray_bundle)
        jitter = None
        if self.train_stratified and self.training:
            shape = (rays.num_rays, 1) if self.single_jitter else (rays.num_rays, num_samples + 1)
            jitter = torch.rand(shape, dtype=torch.float32, device=rays.origins.device)
        sbins, ebins = F.spaced_bins(rays, num_samples, jitter, self.lambda_, self.scaling)
        return _make_ray_samples(ray_bundle, rays, sbins, ebins, (self.lambda_, self.scaling),
                                 self._closure(ray_bundle.nears, ray_bundle.fars))


class PDFSampler(Sampler):
    """Inverse-CDF importance sampling, one warp per ray."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__(num_samples=num_samples)
        if include_original:
            raise NotImplementedError("include_original=True is not used on the NeuRadar path")
        self.train_stratified = train_stratified
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.single_jitter = single_jitter

    def generate_ray_samples(
        self,
        ray_bundle: Optional[RayBundle] = None,
        ray_samples: Optional[RaySamples] = None,
        weights: Optional[Tensor] = None,
        num_samples: Optional[int] = None,
        eps: float = 1e-5,
    ) -> RaySamples:
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        assert weights is not None, "weights must be provided"
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        assert ray_samples.spacing_starts is not None and ray_samples.spacing_ends is not None, \
            "ray_sample spacing_starts and spacing_ends must be provided"
        if getattr(ray_samples, "spacing", None) is None:
            raise NotImplementedError("PDFSampler needs ray samples produced by PowerSampler/PDFSampler of this package")
        rays = getattr(ray_samples, "ray_data", None)
        if rays is None or rays.nears is None or rays.fars is None:
            rays = ray_data_of(ray_bundle)
        sbins_in = getattr(ray_samples, "spacing_bins", None)
        if sbins_in is None:
            sbins_in = torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)
        jitter = None
        if self.train_stratified and self.training:
            if not self.single_jitter:
                raise NotImplementedError("PDFSampler on the NeuRadar path uses single_jitter=True")
            jitter = torch.rand((rays.num_rays, 1), device=rays.origins.device)
        lam, scaling = ray_samples.spacing
        sbins, ebins = F.pdf_sample(rays, weights[..., 0], sbins_in, num_samples, jitter, lam, scaling,
                                    self.histogram_padding, eps)
        return _make_ray_samples(ray_bundle, rays, sbins, ebins, ray_samples.spacing, ray_samples.spacing_to_euclidean_fn)


def _data_parallel() -> bool:
    """Under data parallelism the backward stays serial: the all-reduce of the main table's gradient (dist.GradArena,
    `early`) then hides behind the proposal backward, which is worth more than overlapping the two backward branches."""
    import torch.distributed as dist

    if os.environ.get("NRB_DP_SERIAL_BACKWARD", "1") == "0":  # experiment switch: overlap the branches under DP as well
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class ProposalNetworkSampler(Sampler):
    """Proposal-network sampling loop (ray_samplers.py:623-666)."""

    def __init__(
        self,
        num_proposal_samples_per_ray: Tuple[int, ...] = (64,),
        num_nerf_samples_per_ray: int = 32,
        num_proposal_network_iterations: int = 2,
        single_jitter: bool = False,
        update_sched: Callable = lambda x: 1,
        initial_sampler: Optional[Sampler] = None,
        pdf_sampler: Optional[PDFSampler] = None,
    ) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        if initial_sampler is None:
            raise NotImplementedError("pass initial_sampler=PowerSampler(...) as NeuRadarModel does")
        self.initial_sampler = initial_sampler
        self.pdf_sampler = pdf_sampler if pdf_sampler is not None else PDFSampler(include_original=False,
                                                                                  single_jitter=single_jitter)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0

    overlap_backward = os.environ.get("NRB_OVERLAP_PROPOSALS", "1") != "0"
    """Training: the proposal rounds run on a side stream, so that autograd replays their backward (which depends only on
    the losses on the proposal weights - the PDF sampler is not differentiated) concurrently with the field's backward.
    The proposal scatter is bound by L2 reductions, the field's backward by tensor-core latency and instruction issue:
    measured 4.95 -> 4.86 ms/step at config 2 (B200, CUDA graph).  A persistent proposal backward sized to leave room on
    every SM was measured too and is slower (it needs its full occupancy): DESIGN.md section 4."""

    def _on_side_stream(self, fn, ray_samples):
        dev = ray_samples.frustums.starts.device
        if not (self.overlap_backward and torch.is_grad_enabled() and dev.type == "cuda") or _data_parallel():
            return fn()
        main = torch.cuda.current_stream(dev)
        F.note_consumer_stream(dev)
        side = F.side_stream(dev, 2)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            out = fn()
        main.wait_stream(side)
        out.record_stream(main)
        return out

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def generate_ray_samples(
        self,
        ray_bundle: Optional[RayBundle] = None,
        density_fns: Optional[List[Callable]] = None,
        pass_ray_samples: bool = False,
    ) -> Tuple[RaySamples, List, List]:
        assert ray_bundle is not None
        assert density_fns is not None
        if not pass_ray_samples:
            density_fns = [lambda rs, f=f: f(rs.frustums.get_positions()) for f in density_fns]
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights = None
        ray_samples = None
        updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples)
            else:
                assert weights is not None
                annealed = weights if self._anneal == 1.0 else torch.pow(weights, self._anneal)
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, annealed, num_samples=num_samples)
            if is_prop:
                fn = density_fns[i_level]
                fused = getattr(fn, "density_and_weights", None)
                with torch.set_grad_enabled(updated and torch.is_grad_enabled()):
                    if fused is not None:  # proposal field of this package: density + weights in one kernel
                        weights = self._on_side_stream(lambda: fused(ray_samples)[1], ray_samples)
                    else:
                        weights = ray_samples.get_weights(fn(ray_samples))
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        assert ray_samples is not None
        return ray_samples, weights_list, ray_samples_list
