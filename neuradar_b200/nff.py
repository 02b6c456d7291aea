"""The per-ray hot path of NeuRadarModel as one module: proposal sampling -> field -> compositing.

Mirrors `NeuRadarModel.populate_modules` (field / proposal_fields / sampler / density_fns / renderers,
nerfstudio/models/neuradar.py:198-325), `_get_ray_samples` (:570-586), `_scale_pixel_area` (:996-1008),
`_render_weights` (:1010-1023) and `get_nff_outputs` (:495-548).  The decoders that consume its outputs (camera CNN,
lidar MLP, radar transformer) stay in PyTorch upstream of this module and are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import functional as F
from .fields import (
    FieldHeadNames,
    NeuRADField,
    NeuRADFieldConfig,
    NeuRADProposalField,
    NeuRADProposalFieldConfig,
)
from .ray_samplers import PowerSampler, ProposalNetworkSampler
from .rays import RayBundle, RaySamples

EPS = 1.0e-7


@dataclass
class SamplingSettings:
    """models/neuradar.py:118-138"""

    single_jitter: bool = True
    proposal_field_1: NeuRADProposalFieldConfig = field(default_factory=NeuRADProposalFieldConfig)
    proposal_field_2: NeuRADProposalFieldConfig = field(default_factory=NeuRADProposalFieldConfig)
    num_proposal_samples: Tuple[int, ...] = (128, 64)
    num_nerf_samples: int = 32
    power_lambda: float = -1.0
    power_scaling: float = 0.1
    sky_distance: float = 20000.0


@dataclass
class NeuRadarHotPathConfig:
    sampling: SamplingSettings = field(default_factory=SamplingSettings)
    field: NeuRADFieldConfig = field(default_factory=NeuRADFieldConfig)
    rgb_upsample_factor: int = 3
    static_scale: float = 100.0
    late_binding_density_fns: bool = True
    """Reproduce the reference's late-binding lambda list (models/neuradar.py:302): every proposal round queries
    the LAST proposal field.  False gives each round its own field."""


class DensityFn:
    """Callable `ray_samples -> density` that also exposes the fused density+weights kernel of its field."""

    def __init__(self, proposal_field: NeuRADProposalField):
        self._field = proposal_field

    def __call__(self, ray_samples: RaySamples) -> Tensor:
        return self._field.get_density(ray_samples)[0]

    def density_and_weights(self, ray_samples: RaySamples):
        return self._field.density_and_weights(ray_samples)


class NeuRadarHotPath(nn.Module):
    def __init__(self, config: NeuRadarHotPathConfig, actors=None):
        super().__init__()
        self.config = config
        s = config.sampling
        self.sampler = ProposalNetworkSampler(
            num_proposal_samples_per_ray=s.num_proposal_samples,
            num_nerf_samples_per_ray=s.num_nerf_samples,
            num_proposal_network_iterations=len(s.num_proposal_samples),
            single_jitter=s.single_jitter,
            initial_sampler=PowerSampler(lambda_=s.power_lambda, scaling=s.power_scaling),
            update_sched=lambda x: 0,
        )
        self.proposal_fields = nn.ModuleList(
            [conf.setup(actors=actors, static_scale=config.static_scale) for conf in (s.proposal_field_1, s.proposal_field_2)]
        )
        if config.late_binding_density_fns:
            self.density_fns = [DensityFn(self.proposal_fields[-1]) for _ in self.proposal_fields]
        else:
            self.density_fns = [DensityFn(f) for f in self.proposal_fields]
        self.field: NeuRADField = config.field.setup(actors=actors, static_scale=config.static_scale)

    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        groups: Dict[str, List[nn.Parameter]] = {"hashgrids": [], "fields": []}
        self.field.get_param_groups(groups)
        for f in self.proposal_fields:
            f.get_param_groups(groups)
        return groups

    def _scale_pixel_area(self, ray_bundle: RayBundle) -> None:
        is_lidar = ray_bundle.metadata.get("is_lidar")
        is_radar = ray_bundle.metadata.get("is_radar")
        up2 = self.config.rgb_upsample_factor**2
        scaling = torch.ones_like(ray_bundle.pixel_area)
        if is_lidar is not None and is_radar is not None:
            scaling[~(is_lidar | is_radar)] = up2
        elif is_lidar is not None:
            scaling[~is_lidar] = up2
        elif is_radar is not None:
            scaling[~is_radar] = up2
        else:
            scaling = up2
        ray_bundle.pixel_area = ray_bundle.pixel_area * scaling

    def _get_ray_samples(self, ray_bundle: RayBundle):
        sky = self.config.sampling.sky_distance
        if ray_bundle.fars is not None:
            ray_bundle.fars.clamp_max_(sky)
        else:
            ray_bundle.fars = torch.full_like(ray_bundle.pixel_area, sky)
        ray_bundle.nears = ray_bundle.nears if ray_bundle.nears is not None else torch.zeros_like(ray_bundle.fars)
        ray_samples, prop_weights, prop_ray_samples = self.sampler(ray_bundle, self.density_fns, pass_ray_samples=True)
        # "sky field": stretch the last sample to sky_distance (models/neuradar.py:578-582); the bins tensor is
        # shared by frustums.ends / spacing_ends, so one in-place edit of its last column updates every view
        ebins, sbins = ray_samples.euclidean_bins, ray_samples.spacing_bins
        dist_to_sky = sky - ebins[:, -1]
        ebins[:, -1] += dist_to_sky
        ray_samples.deltas[..., -1, 0] += dist_to_sky
        sbins[:, -1] = 1 - EPS
        return ray_samples, prop_ray_samples, prop_weights

    def get_nff_outputs(self, ray_bundle: RayBundle, calc_lidar_losses: bool = False) -> Dict[str, Tensor]:
        self._scale_pixel_area(ray_bundle)
        ray_samples, proposal_ray_samples, proposal_weights = self._get_ray_samples(ray_bundle)
        if self.field.can_render(ray_samples):
            # field + _render_weights + AccumulationRenderer + sky fix-up + FeatureRenderer + render_depth_simple as ONE
            # autograd node: three kernels forward, three backward, no [N,S,32] gradient tensor
            weights, features, depth, accumulation = self.field.render(ray_samples, trans_eps=0.0)
        else:
            outputs = self.field(ray_samples)
            N, S = ray_samples.shape
            alpha = outputs[FieldHeadNames.ALPHA].reshape(N, S)
            weights, features, depth, accumulation = F.alpha_composite(
                alpha, outputs[FieldHeadNames.FEATURE], ray_samples.intervals(), trans_eps=0.0, sky_sample=True
            )
        weights = weights[:, :-1, None]  # the sky sample is discarded for everything downstream (:515)
        nff_outputs = {"features": features, "depth": depth[:, None], "accumulation": accumulation[:, None]}
        for i, (prop_w, prop_rs) in enumerate(zip(proposal_weights, proposal_ray_samples)):
            steps = (prop_rs.frustums.starts + prop_rs.frustums.ends) / 2
            nff_outputs[f"prop_depth_{i}"] = F.accumulate(prop_w[..., 0], steps)
        if self.training:
            nff_outputs["weights_list"] = proposal_weights + [weights]
            nff_outputs["ray_samples_list"] = proposal_ray_samples + [ray_samples[..., :-1]]
        return nff_outputs

    def forward(self, ray_bundle: RayBundle) -> Dict[str, Tensor]:
        return self.get_nff_outputs(ray_bundle)


def bench_loss(out: Dict[str, Tensor]) -> Tensor:
    """Upstream loss of the synthetic train step: touches every output of the path."""
    loss = out["features"].pow(2).mean() + 1e-3 * out["depth"].mean()
    for w in out["weights_list"][:-1]:
        loss = loss + w.pow(2).mean()
    return loss


def training_losses(out: Dict[str, Tensor], interlevel_mult: float = 0.001, distortion_mult: float = 0.002) -> Tensor:
    """The sampler regularisers NeuRadarModel.get_loss_dict adds every step (models/neurad.py:524-545; multipliers
    :83-85): ZipNeRF interlevel loss of both proposal rounds + MipNeRF-360 distortion loss of the final level."""
    from .losses import distortion_loss, zipnerf_interlevel_loss

    wl, rl = out["weights_list"], out["ray_samples_list"]
    return interlevel_mult * zipnerf_interlevel_loss(wl, rl) + distortion_mult * distortion_loss(wl, rl)
