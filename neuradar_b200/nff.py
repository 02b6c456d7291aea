"""The per-ray hot path of NeuRadarModel as one module: proposal sampling -> field -> compositing.

Mirrors `NeuRadarModel.populate_modules` (field / proposal_fields / sampler / density_fns / renderers,
nerfstudio/models/neuradar.py:198-325), `_get_ray_samples` (:570-586), `_scale_pixel_area` (:996-1008),
`_render_weights` (:1010-1023) and `get_nff_outputs` (:495-548).  The decoders that consume its outputs (camera CNN,
lidar MLP, radar transformer) stay in PyTorch upstream of this module and are out of scope.
"""
from __future__ import annotations

import dataclasses
import os
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
class LossSettings:
    """The two loss settings the path itself reads (models/neurad.py:79,87)."""

    carving_epsilon: float = 0.1
    non_return_lidar_distance: float = 150.0


@dataclass
class NeuRadarHotPathConfig:
    sampling: SamplingSettings = field(default_factory=SamplingSettings)
    field: NeuRADFieldConfig = field(default_factory=NeuRADFieldConfig)
    rgb_upsample_factor: int = 3
    static_scale: float = 100.0
    late_binding_density_fns: bool = True
    """Reproduce the reference's late-binding lambda list (models/neuradar.py:302): every proposal round queries
    the LAST proposal field.  False gives each round its own field."""
    loss: "LossSettings" = dataclasses.field(default_factory=lambda: LossSettings())
    appearance_dim: int = 0
    """> 0 adds the per-sensor appearance code to the rendered features (models/neuradar.py:205-215,510-512)."""
    num_sensors: int = 1
    use_temporal_appearance: bool = False
    num_embeds_per_sensor: int = 1
    sequence_duration: float = 1.0
    gather_non_nearby: bool = False
    """Also emit `non_nearby_weights` / `non_nearby_lidar_ray_indices` exactly as the reference does (a nonzero()
    gather, i.e. a host synchronisation); `non_nearby_mask` and `non_nearby_weights_sq_sum` are always there."""


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
        # Overlapping the proposal backward with the field's pays while both gradient tables stay L2-resident (+1.8 % /
        # +3.4 % at configs 2 / 3); with the 512 MiB table of config 4 the two scatters fight over L2 instead (-3.7 %).
        if "NRB_OVERLAP_PROPOSALS" not in os.environ:
            scattered = self.field.hashgrid.static_grid.hash_table.numel() + self.proposal_fields[-1].hashgrid.static_grid.hash_table.numel()
            self.sampler.overlap_backward = scattered * 4 <= 96 << 20
        self.appearance_embedding = None
        if config.appearance_dim > 0:
            self.appearance_embedding = nn.Embedding(config.num_sensors * config.num_embeds_per_sensor, config.appearance_dim)

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
        grids = [self.field.hashgrid] + [p.hashgrid for p in self.proposal_fields]
        poses: dict = {}
        for h in grids:  # the actor poses at the rays' times are shared by the three encodings of this step
            h.pose_cache = poses
        try:
            return self._nff_outputs(ray_bundle, calc_lidar_losses)
        finally:
            for h in grids:
                h.pose_cache = None

    def _nff_outputs(self, ray_bundle: RayBundle, calc_lidar_losses: bool) -> Dict[str, Tensor]:
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
        if self.appearance_embedding is not None:  # models/neuradar.py:510-512,550-568
            nff_outputs["features"] = torch.cat([features, self._get_appearance_embedding(ray_bundle, features)], dim=-1)
        md = ray_bundle.metadata or {}
        carve = self.training and calc_lidar_losses and "is_lidar" in md and "directions_norm" in md
        lc = self.config.loss
        for i, (prop_w, prop_rs) in enumerate(zip(proposal_weights, proposal_ray_samples)):
            nff_outputs[f"prop_depth_{i}"] = F.weighted_depth(prop_w[..., 0], prop_rs.intervals())
            if carve:  # carving loss of the proposal rounds (:529-531): sum((w * (is_lidar & ~is_close_to_lidar))^2)
                nff_outputs[f"prop_weights_loss_{i}"] = F.carving_loss(
                    prop_w[..., 0], prop_rs.intervals(), md["is_lidar"], md["directions_norm"], md.get("did_return"),
                    lc.carving_epsilon, lc.non_return_lidar_distance)
        if self.training:
            nff_outputs["weights_list"] = proposal_weights + [weights]
            nff_outputs["ray_samples_list"] = proposal_ray_samples + [ray_samples[..., :-1]]
        if carve:  # :537-546
            iv = F.SampleIntervals(ray_samples.frustums.starts[:, :-1], ray_samples.frustums.ends[:, :-1])
            close = F.is_close_to_lidar(iv, md["is_lidar"], md["directions_norm"], md.get("did_return"), lc.carving_epsilon,
                                        lc.non_return_lidar_distance)
            mask = (~close) & md["is_lidar"].reshape(-1, 1).bool()
            nff_outputs["non_nearby_mask"] = mask
            # the sum of squares the reference forms from the gathered weights (:638), without the nonzero() round trip
            nff_outputs["non_nearby_weights_sq_sum"] = F.carving_loss(
                weights[..., 0], iv, md["is_lidar"], md["directions_norm"], md.get("did_return"), lc.carving_epsilon,
                lc.non_return_lidar_distance)
            if self.config.gather_non_nearby:  # exactly the reference's outputs; nonzero() synchronises the host
                idx = mask.nonzero(as_tuple=True)
                nff_outputs["non_nearby_weights"] = weights[..., 0][idx][:, None]
                lidar_start = md["is_lidar"].reshape(-1).int().argmax()
                nff_outputs["non_nearby_lidar_ray_indices"] = (idx[0] - lidar_start)[:, None]
        return nff_outputs

    def _get_appearance_embedding(self, ray_bundle: RayBundle, features: Tensor) -> Tensor:
        """models/neuradar.py:550-568: per-sensor (optionally time-interpolated) appearance code."""
        c = self.config
        sensor_idx = (ray_bundle.metadata or {}).get("sensor_idxs")
        if sensor_idx is None:
            assert not self.training, "Sensor sensor_idx must be present in metadata during training"
            sensor_idx = torch.zeros_like(features[..., :1], dtype=torch.long)
        if c.use_temporal_appearance:
            n_e = c.num_embeds_per_sensor
            time_idx = ray_bundle.times / c.sequence_duration * n_e
            before = time_idx.floor().clamp(0, n_e - 1)
            after = (before + 1).clamp(0, n_e - 1)
            ratio = time_idx - before
            before, after = (x + sensor_idx * n_e for x in (before, after))
            return (self.appearance_embedding(before.squeeze(-1).long()) * (1 - ratio)
                    + self.appearance_embedding(after.squeeze(-1).long()) * ratio)
        return self.appearance_embedding(sensor_idx.squeeze(-1))

    def point_heads(self, ray_bundle: RayBundle, depth: Tensor, world2lidar: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """The point heads on a rendered depth [N,1]: `points` for lidar / camera rays = o + d * depth,
        in the lidar frame when world2lidar is given (models/ad_model.py:103-108); for the rays flagged is_radar the
        cartesian position from the spherical direction (models/neuradar.py:463-473,1025-1029) that feeds the radar
        decoder's positional embedding.  One kernel, differentiable in depth."""
        md = ray_bundle.metadata or {}
        is_radar = md.get("is_radar")
        spher = md.get("directions_spher")
        if is_radar is not None and spher is None:
            is_radar = None
        pts = F.point_heads(depth, ray_bundle.origins.reshape(-1, 3), ray_bundle.directions.reshape(-1, 3), is_radar, spher,
                            world2lidar)
        out = {"points": pts}
        if is_radar is not None:
            out["radar_xyz"] = pts[is_radar.reshape(-1)]
        return out

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
