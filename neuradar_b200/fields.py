"""NeuRAD fields with the reference's API: `forward(ray_samples)`, `get_density(ray_samples)`, `get_outputs`.

NeuRADField / NeuRADFieldConfig                   <- nerfstudio/fields/neurad_field.py:45-152
NeuRADProposalField / NeuRADProposalFieldConfig   <- nerfstudio/fields/neurad_field.py:155-216
Field base API (`get_density`, `get_outputs`, `forward`) <- nerfstudio/fields/base_field.py:40-133
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, Optional, Tuple, Type

import torch
from torch import Tensor, nn

from . import functional as F
from .field_components import (
    MLP,
    ActorSettings,
    NeuRADHashEncoding,
    NeuRADHashEncodingConfig,
    SHEncoding,
    SigmoidDensity,
    StaticSettings,
    trunc_exp,
)
from .rays import RaySamples, per_ray_of


class FieldHeadNames(Enum):
    """Subset of nerfstudio/field_components/field_heads.py used on the path."""

    FEATURE = "feature"
    DENSITY = "density"
    SDF = "sdf"
    ALPHA = "alpha"

    # The reference model indexes a field's output dict with ITS enum (models/neuradar.py:1011): members of the two
    # classes must be interchangeable as dictionary keys, so equality and hashing go by member name / value.
    def __eq__(self, other):
        if isinstance(other, Enum) and type(other).__name__ == "FieldHeadNames":
            return self.value == other.value
        return NotImplemented

    def __hash__(self):
        return hash(self._name_)


def get_normalized_directions(directions: Tensor) -> Tensor:
    """(d + 1) / 2 (fields/base_field.py:136-142)."""
    return (directions + 1.0) / 2.0


@dataclass
class NeuRADFieldConfig:
    _target: Type = field(default_factory=lambda: NeuRADField)
    grid: NeuRADHashEncodingConfig = field(
        default_factory=lambda: NeuRADHashEncodingConfig(require_actor_grad=True, actor=ActorSettings(flip_prob=0.25))
    )
    geo_hidden_dim: int = 32
    geo_num_layers: int = 2
    nff_hidden_dim: int = 32
    nff_num_layers: int = 3
    nff_out_dim: int = 32
    num_multisamples: int = 1
    use_sdf: bool = True
    sdf_beta: float = 20.0
    learnable_beta: bool = True

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class NeuRADField(nn.Module):
    """hash grid -> geometry MLP -> (sdf | embedding) -> SH(direction) -> feature MLP + residual; alpha from the sdf."""

    def __init__(self, config: NeuRADFieldConfig, actors=None, static_scale: float = 1.0, implementation: str = "b200"):
        super().__init__()
        self.config = config
        self.implementation = implementation
        self.use_tensor_cores = True
        if config.num_multisamples != 1:
            raise NotImplementedError("num_multisamples must be 1")
        self.hashgrid: NeuRADHashEncoding = config.grid.setup(dynamic_actors=actors, static_scale=static_scale)
        self.geo_feat_dim = config.nff_out_dim
        self.mlp_geo = MLP(
            in_dim=self.hashgrid.get_out_dim(),
            num_layers=config.geo_num_layers,
            layer_width=config.geo_hidden_dim,
            out_dim=self.geo_feat_dim + 1,
        )
        self.direction_encoding = SHEncoding(levels=4)
        self.mlp_feature = MLP(
            in_dim=self.direction_encoding.get_out_dim() + self.geo_feat_dim,
            num_layers=config.nff_num_layers,
            layer_width=config.nff_hidden_dim,
            out_dim=config.nff_out_dim,
        )
        if config.use_sdf:
            self.sdf_to_density = SigmoidDensity(config.sdf_beta, learnable_beta=config.learnable_beta)

    def get_param_groups(self, param_groups: Dict):
        self.hashgrid.get_param_groups(param_groups)
        param_groups["fields"] += list(self.mlp_geo.parameters()) + list(self.mlp_feature.parameters())
        if self.config.use_sdf:
            param_groups["fields"] += list(self.sdf_to_density.parameters())

    def _tensor_core_path(self) -> bool:
        """The fused tcgen05 kernel is built for NeuRadar's default field: 32 hash features, geometry MLP
        32 -> 32 -> 33, feature MLP 48 -> 32 -> 32 -> 32, SDF head.  Other shapes run the generic MLP kernels."""
        c = self.config
        return (
            self.use_tensor_cores
            and c.use_sdf
            and self.hashgrid.get_out_dim() == 32
            and (c.geo_hidden_dim, c.geo_num_layers, c.nff_hidden_dim, c.nff_num_layers, c.nff_out_dim) == (32, 2, 32, 3, 32)
        )

    def _field_chunk(self, rays: F.RayData, iv: F.SampleIntervals, times: Optional[Tensor] = None):
        """hash encode + everything after it (ONE tcgen05 kernel forward, one backward) for a set of rays."""
        geo_l, feat_l = self.mlp_geo.layers, self.mlp_feature.layers
        weights = [geo_l[0].weight, geo_l[1].weight, feat_l[0].weight, feat_l[1].weight, feat_l[2].weight]
        biases = [geo_l[0].bias, geo_l[1].bias, feat_l[0].bias, feat_l[1].bias, feat_l[2].bias]
        beta, beta_min = self.sdf_to_density.beta, self.sdf_to_density.beta_min_value
        grid = self.hashgrid.static_grid
        with_actors = self.hashgrid.has_actors and times is not None
        if grid.features_per_level in (2, 4) and (not with_actors or self.hashgrid.can_assign_in_kernel()):
            # the 32 hash features are gathered inside the MLP kernel and never reach HBM; samples inside an actor box
            # (assigned by one kernel) read their actor's grid there
            x3, std = F.frustum_gaussians(rays, iv, self.hashgrid.static_scale)
            sh = self.direction_encoding(get_normalized_directions(rays.directions))
            actors = self.hashgrid.assign_actors(rays, iv, times) if with_actors else None
            return F.field_fused(grid.hash_table, None, x3, std, sh, iv.num_samples, grid.spec, weights, biases, beta, beta_min,
                                 actors)
        features, sample_dirs = self.hashgrid.encode_samples(rays, iv, times)
        if sample_dirs is None:  # directions are per ray: 16 SH values per ray, indexed by row / S in the kernel
            sh = self.direction_encoding(get_normalized_directions(rays.directions))
            sh_group = iv.num_samples
        else:  # some samples were rotated into an actor frame: one SH row per sample
            sh = self.direction_encoding(get_normalized_directions(sample_dirs.reshape(-1, 3)))
            sh_group = 1
        return F.field_fused(None, features, None, None, sh, sh_group, None, weights, biases, beta, beta_min)

    def can_render(self, ray_samples: RaySamples) -> bool:
        """True when `render` applies: default field shape, a 32-feature static grid and no dynamic actors."""
        grid = self.hashgrid.static_grid
        return (self._tensor_core_path() and grid.features_per_level in (2, 4) and len(ray_samples.shape) == 2
                and (not self.hashgrid.has_actors or (self.hashgrid.can_assign_in_kernel() and ray_samples.times is not None)))

    def render(self, ray_samples: RaySamples, trans_eps: float = 0.0):
        """Field + compositing tail in one autograd node (models/neuradar.py:500-517): returns (weights [N,S] after the
        sky fix-up, features [N,32], depth [N], accumulation [N]).  The [N,S,32] feature gradient is never formed."""
        rays, iv = per_ray_of(ray_samples)
        geo_l, feat_l = self.mlp_geo.layers, self.mlp_feature.layers
        weights = [geo_l[0].weight, geo_l[1].weight, feat_l[0].weight, feat_l[1].weight, feat_l[2].weight]
        biases = [geo_l[0].bias, geo_l[1].bias, feat_l[0].bias, feat_l[1].bias, feat_l[2].bias]
        grid = self.hashgrid.static_grid
        x3, std = F.frustum_gaussians(rays, iv, self.hashgrid.static_scale)
        sh = self.direction_encoding(get_normalized_directions(rays.directions))
        actors = None
        if self.hashgrid.has_actors:
            actors = self.hashgrid.assign_actors(rays, iv, ray_samples.times.reshape(rays.num_rays, -1)[:, 0])
        return F.field_render(grid.hash_table, x3, std, sh, iv, grid.spec, weights, biases, self.sdf_to_density.beta,
                              self.sdf_to_density.beta_min_value, trans_eps, actors)

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, Tensor]:
        if compute_normals:
            raise NotImplementedError("normals are not rendered on the NeuRadar path")
        rays, iv = per_ray_of(ray_samples)
        N, S = rays.num_rays, iv.num_samples
        shape = tuple(ray_samples.shape) if len(ray_samples.shape) == 2 else (N, S)
        times = None
        if self.hashgrid.has_actors:
            if ray_samples.times is None:
                raise ValueError("ray_samples.times is required in a scene with dynamic actors")
            times = ray_samples.times.reshape(N, -1)[:, 0]
        if self._tensor_core_path():
            feature, sdf, alpha = self._field_chunk(rays, iv, times)
            return {
                FieldHeadNames.FEATURE: feature.view(*shape, 32),
                FieldHeadNames.SDF: sdf.view(*shape, 1),
                FieldHeadNames.ALPHA: alpha.view(*shape, 1),
            }
        features, sample_dirs = self.hashgrid.encode_samples(rays, iv, times)
        geo = self.mlp_geo(features)
        geo_out, geo_embedding = torch.split(geo, [1, self.geo_feat_dim], dim=-1)
        if sample_dirs is None:  # per-ray directions: 16 SH values per ray, broadcast over the samples
            sh = self.direction_encoding(get_normalized_directions(rays.directions))
            sh = sh[:, None, :].expand(N, S, 16).reshape(N * S, 16)
        else:
            sh = self.direction_encoding(get_normalized_directions(sample_dirs.reshape(-1, 3)))
        feature = geo_embedding + self.mlp_feature(torch.cat([geo_embedding, sh], dim=-1))
        outputs = {FieldHeadNames.FEATURE: feature.view(*shape, self.config.nff_out_dim)}
        geo_out = geo_out.reshape(*shape, 1)
        if self.config.use_sdf:
            outputs[FieldHeadNames.SDF] = geo_out
            outputs[FieldHeadNames.ALPHA] = self.sdf_to_density(geo_out)
        else:
            outputs[FieldHeadNames.DENSITY] = trunc_exp(geo_out)
        return outputs


@dataclass
class NeuRADProposalFieldConfig:
    _target: Type = field(default_factory=lambda: NeuRADProposalField)
    grid: NeuRADHashEncodingConfig = field(
        default_factory=lambda: NeuRADHashEncodingConfig(
            static=StaticSettings(log2_hashmap_size=20, num_levels=6, max_res=4096, base_res=128, hashgrid_dim=1),
            actor=ActorSettings(log2_hashmap_size=15, num_levels=4, base_res=64, max_res=1024, hashgrid_dim=1),
            require_actor_grad=False,
        )
    )
    hidden_dim: int = 16

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class NeuRADProposalField(nn.Module):
    """hash grid -> Linear(L*F, 1, bias=False) -> trunc_exp.  `get_density` runs the fused proposal kernel."""

    def __init__(self, config: NeuRADProposalFieldConfig, actors=None, static_scale: float = 1.0,
                 implementation: str = "b200"):
        super().__init__()
        self.config = config
        self.implementation = implementation
        self.hashgrid: NeuRADHashEncoding = config.grid.setup(dynamic_actors=actors, static_scale=static_scale)
        self.density_decoder = nn.Linear(self.hashgrid.get_out_dim(), 1, bias=False)

    def get_param_groups(self, param_groups: Dict):
        self.hashgrid.get_param_groups(param_groups)
        param_groups["fields"] += list(self.density_decoder.parameters())

    def density_and_weights(self, ray_samples: RaySamples) -> Tuple[Tensor, Tensor]:
        """One kernel for get_density + RaySamples.get_weights: ([N,S,1], [N,S,1])."""
        rays, iv = per_ray_of(ray_samples)
        if self.hashgrid.has_actors:
            if ray_samples.times is None:
                raise ValueError("ray_samples.times is required in a scene with dynamic actors")
            if self.hashgrid.can_assign_in_kernel(proposal=True):  # one assignment kernel + the fused round, no host sync
                grid = self.hashgrid.static_grid
                actors = self.hashgrid.assign_actors(rays, iv, ray_samples.times.reshape(rays.num_rays, -1)[:, 0])
                dens, w = F.proposal_round(grid.hash_table, self.density_decoder.weight, rays, iv, grid.spec,
                                           self.hashgrid.static_scale, actors=actors)
                return dens.unsqueeze(-1), w.unsqueeze(-1)
            # unfused: the torch bookkeeping rewrites features between the grid and the decoder
            feats, _ = self.hashgrid.encode_samples(rays, iv, ray_samples.times.reshape(rays.num_rays, -1)[:, 0])
            dens = trunc_exp(self.density_decoder(feats)).view(rays.num_rays, iv.num_samples)
            return dens.unsqueeze(-1), F.density_weights(dens, iv).unsqueeze(-1)
        grid = self.hashgrid.static_grid
        dens, w = F.proposal_round(grid.hash_table, self.density_decoder.weight, rays, iv, grid.spec,
                                   self.hashgrid.static_scale)
        return dens.unsqueeze(-1), w.unsqueeze(-1)

    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, None]:
        dens, _ = self.density_and_weights(ray_samples)
        return dens.view(*ray_samples.shape, 1), None

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None) -> dict:
        return {}

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, Tensor]:
        density, _ = self.get_density(ray_samples)
        return {FieldHeadNames.DENSITY: density}
