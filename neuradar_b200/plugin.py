"""Drop-in door into an existing nerfstudio / neurad-studio installation.

Nothing here is imported by the kernels or the benchmark; it only runs where the reference (`nerfstudio`) is
installed.  Two doors:

1. `convert_neuradar_model(model)`: swap the hot-path submodules of an already constructed `NeuRadarModel`
   (`field`, `proposal_fields`, `density_fns`, `sampler`, feature/accumulation renderers) for the B200 ones, carrying
   the parameters over (state-dict keys are identical to the reference's torch path).
2. `method_spec()`: a `MethodSpecification` "neuradar-b200" for the plugin registry
   (`nerfstudio/plugins/registry.py:35-83`), i.e. `NERFSTUDIO_METHOD_CONFIGS="neuradar-b200=neuradar_b200.plugin:method_spec"`.
   Its model class is `NeuRadarModel` whose `populate_modules` ends with `convert_neuradar_model(self)`.
"""
from __future__ import annotations

import copy
from typing import Any

import torch
from torch import nn

from . import nerfacc_compat
from .field_components import ActorSettings, NeuRADHashEncodingConfig, StaticSettings
from .fields import NeuRADField, NeuRADFieldConfig, NeuRADProposalField, NeuRADProposalFieldConfig
from .nff import DensityFn
from .ray_samplers import PDFSampler, PowerSampler, ProposalNetworkSampler
from .renderers import AccumulationRenderer, FeatureRenderer


def _grid_config(ref_grid_cfg: Any) -> NeuRADHashEncodingConfig:
    s, a = ref_grid_cfg.static, ref_grid_cfg.actor
    return NeuRADHashEncodingConfig(
        static=StaticSettings(hashgrid_dim=s.hashgrid_dim, num_levels=s.num_levels, base_res=s.base_res, max_res=s.max_res,
                              log2_hashmap_size=s.log2_hashmap_size),
        actor=ActorSettings(flip_prob=a.flip_prob, actor_scale=a.actor_scale, hashgrid_dim=a.hashgrid_dim,
                            num_levels=a.num_levels, base_res=a.base_res, max_res=a.max_res,
                            log2_hashmap_size=a.log2_hashmap_size, use_4d_hashgrid=a.use_4d_hashgrid),
        disable_actors=ref_grid_cfg.disable_actors,
        require_actor_grad=ref_grid_cfg.require_actor_grad,
    )


def _carry_parameters(dst: nn.Module, src: nn.Module) -> None:
    """Same names, same shapes: load the reference module's state dict (its actor-trajectory entries have no
    counterpart here and must be the only leftovers)."""
    result = dst.load_state_dict(src.state_dict(), strict=False)
    bad = [k for k in result.unexpected_keys if ".actors." not in k and "actor_grids" not in k]
    if bad or result.missing_keys:
        raise RuntimeError(f"state dicts do not line up: missing={result.missing_keys} unexpected={bad}")


def _static_scale(ref_hashgrid: Any) -> float:
    return float(ref_hashgrid.static_contraction.scale)


def convert_field(ref_field: Any) -> NeuRADField:
    """`nerfstudio.fields.neurad_field.NeuRADField` (torch or tcnn implementation) -> B200 NeuRADField."""
    c = ref_field.config
    cfg = NeuRADFieldConfig(grid=_grid_config(c.grid), geo_hidden_dim=c.geo_hidden_dim, geo_num_layers=c.geo_num_layers,
                            nff_hidden_dim=c.nff_hidden_dim, nff_num_layers=c.nff_num_layers, nff_out_dim=c.nff_out_dim,
                            num_multisamples=c.num_multisamples, use_sdf=c.use_sdf, sdf_beta=c.sdf_beta,
                            learnable_beta=c.learnable_beta)
    new = NeuRADField(cfg, actors=ref_field.hashgrid.actors, static_scale=_static_scale(ref_field.hashgrid))
    _carry_parameters(new, ref_field)
    return new.to(next(ref_field.parameters()).device)


def convert_proposal_field(ref_field: Any) -> NeuRADProposalField:
    c = ref_field.config
    cfg = NeuRADProposalFieldConfig(grid=_grid_config(c.grid), hidden_dim=c.hidden_dim)
    new = NeuRADProposalField(cfg, actors=ref_field.hashgrid.actors, static_scale=_static_scale(ref_field.hashgrid))
    _carry_parameters(new, ref_field)
    return new.to(next(ref_field.parameters()).device)


def convert_sampler(ref_sampler: Any, power_lambda: float, power_scaling: float) -> ProposalNetworkSampler:
    new = ProposalNetworkSampler(
        num_proposal_samples_per_ray=ref_sampler.num_proposal_samples_per_ray,
        num_nerf_samples_per_ray=ref_sampler.num_nerf_samples_per_ray,
        num_proposal_network_iterations=ref_sampler.num_proposal_network_iterations,
        single_jitter=ref_sampler.pdf_sampler.single_jitter,
        update_sched=ref_sampler.update_sched,
        initial_sampler=PowerSampler(lambda_=power_lambda, scaling=power_scaling,
                                     single_jitter=ref_sampler.initial_sampler.single_jitter),
        pdf_sampler=PDFSampler(include_original=False, single_jitter=ref_sampler.pdf_sampler.single_jitter,
                               histogram_padding=ref_sampler.pdf_sampler.histogram_padding),
    )
    new._anneal, new._steps_since_update, new._step = ref_sampler._anneal, ref_sampler._steps_since_update, ref_sampler._step
    return new


def convert_neuradar_model(model: Any) -> Any:
    """In-place swap of the hot path of a constructed NeuRadarModel (nerfstudio/models/neuradar.py:198-325)."""
    nerfacc_compat.install()
    model.field = convert_field(model.field)
    model.proposal_fields = nn.ModuleList([convert_proposal_field(f) for f in model.proposal_fields])
    # the reference builds `density_fns` from late-binding lambdas: every round queries the LAST proposal field
    # (models/neuradar.py:302); keep that behaviour
    model.density_fns = [DensityFn(model.proposal_fields[-1]) for _ in model.proposal_fields]
    s = model.config.sampling
    model.sampler = convert_sampler(model.sampler, s.power_lambda, s.power_scaling)
    model.renderer_feat = FeatureRenderer()
    model.renderer_accumulation = AccumulationRenderer()
    model.config.implementation = "torch"  # parameter names of the torch path are the compatibility surface
    return model


def method_spec():
    """MethodSpecification for `ns-train neuradar-b200` (needs nerfstudio importable)."""
    from nerfstudio.configs.method_configs import method_configs
    from nerfstudio.models.neuradar import NeuRadarModel
    from nerfstudio.plugins.types import MethodSpecification

    class B200NeuRadarModel(NeuRadarModel):
        def populate_modules(self):
            super().populate_modules()
            convert_neuradar_model(self)

    config = copy.deepcopy(method_configs["neuradar"])
    config.method_name = "neuradar-b200"
    config.pipeline.model.implementation = "torch"
    config.pipeline.model._target = B200NeuRadarModel
    return MethodSpecification(config=config, description="NeuRadar with the B200-native per-ray hot path")
