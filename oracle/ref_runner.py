"""Drive the UNMODIFIED reference (mrafidashti/neuradar) torch path on a synthetic workload: test / baseline
infrastructure, imported only by tests/, bench.py (`--impl reference`, `cpu_baseline`) and `__graft_entry__`.

What is built is what `NeuRadarModel.populate_modules` builds for the hot path (nerfstudio/models/neuradar.py:198-325):
`NeuRADField`, two `NeuRADProposalField`s, `ProposalNetworkSampler(PowerSampler)`, the late-binding `density_fns`, and
what `get_nff_outputs` does with them (:495-548), with `implementation="torch"` and autocast off.
Compositing uses `RaySamples.get_weights_and_transmittance_from_alphas` (cameras/rays.py:226-248) because the model's
`_render_weights` returns a constant 0.5 on CPU (:1012-1014).
"""
from __future__ import annotations

import contextlib
import os
import sys
import time
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import ref_shim


def available() -> bool:
    return ref_shim.available()


class ReferencePath:
    """The reference's hot-path objects for one workload (CPU, fp32 torch path)."""

    def __init__(self, main: Tuple[int, int, int, int, int] = (16, 2, 19, 16, 1024), prop_log2: int = 20,
                 proposal_samples: Tuple[int, ...] = (64, 48), nerf_samples: int = 48, static_scale: float = 100.0,
                 seed: int = 42, late_binding: bool = True, actors=None):
        """`actors`: None (the reference's own DynamicActors with no trajectories) or an object with the interface
        NeuRADHashEncoding uses (n_actors, get_boxes2world, actor_bounds, actor_to_id), e.g. synthetic.SyntheticActors."""
        self.root = ref_shim.install()
        with contextlib.redirect_stdout(sys.stderr):  # the reference print()s notices; stdout belongs to the caller
            self._build(main, prop_log2, proposal_samples, nerf_samples, static_scale, seed, late_binding, actors)

    def _build(self, main, prop_log2, proposal_samples, nerf_samples, static_scale, seed, late_binding, actors=None):
        from nerfstudio.field_components.neurad_encoding import NeuRADHashEncodingConfig, StaticSettings
        from nerfstudio.fields.neurad_field import (NeuRADField, NeuRADFieldConfig, NeuRADProposalField,
                                                     NeuRADProposalFieldConfig)
        from nerfstudio.model_components.dynamic_actors import DynamicActors, DynamicActorsConfig
        from nerfstudio.model_components.ray_samplers import PowerSampler, ProposalNetworkSampler

        torch.manual_seed(seed)
        L, F, T, r0, r1 = main
        if actors is None:
            actors = DynamicActors(DynamicActorsConfig(), trajectories=[])
        fcfg = NeuRADFieldConfig()
        fcfg.grid = NeuRADHashEncodingConfig(
            static=StaticSettings(hashgrid_dim=F, num_levels=L, base_res=r0, max_res=r1, log2_hashmap_size=T),
            require_actor_grad=True)
        self.field = NeuRADField(fcfg, actors, static_scale=static_scale, implementation="torch")
        self.proposal_fields = []
        for _ in range(2):
            pcfg = NeuRADProposalFieldConfig()
            pcfg.grid.static.log2_hashmap_size = prop_log2
            self.proposal_fields.append(NeuRADProposalField(pcfg, actors, static_scale, implementation="torch"))
        self.sampler = ProposalNetworkSampler(
            num_proposal_samples_per_ray=tuple(proposal_samples), num_nerf_samples_per_ray=nerf_samples,
            num_proposal_network_iterations=len(proposal_samples), single_jitter=True,
            initial_sampler=PowerSampler(lambda_=-1.0, scaling=0.1), update_sched=lambda x: 0)
        if late_binding:  # models/neuradar.py:302: `lambda x: field.get_density(x)[0]` in a loop -> every round sees the LAST field
            self.density_fns = [lambda rs, f=self.proposal_fields[-1]: f.get_density(rs)[0] for _ in self.proposal_fields]
        else:
            self.density_fns = [lambda rs, f=f: f.get_density(rs)[0] for f in self.proposal_fields]
        self.modules = [self.field, *self.proposal_fields, self.sampler]

    def train(self, mode: bool = True) -> None:
        for m in self.modules:
            m.train(mode)

    def parameters(self) -> List[torch.nn.Parameter]:
        return [p for m in self.modules[:3] for p in m.parameters()]

    def named_parameters(self) -> Dict[str, torch.nn.Parameter]:
        out = {f"field.{k}": v for k, v in self.field.named_parameters()}
        for i, f in enumerate(self.proposal_fields):
            out.update({f"proposal_fields.{i}.{k}": v for k, v in f.named_parameters()})
        return out

    def load_from(self, state: Dict[str, Tensor]) -> None:
        """Copy parameters by name (the B200 modules keep the reference's torch-path parameter names)."""
        mine = self.named_parameters()
        with torch.no_grad():
            for k, v in state.items():
                if k in mine:
                    mine[k].copy_(v.detach().cpu())

    def forward(self, rays: Dict[str, Tensor], pixel_area: Tensor, sky_distance: float = 20000.0):
        """get_nff_outputs on CPU (models/neuradar.py:495-548,570-586): returns the dict the B200 hot path returns."""
        from nerfstudio.cameras.rays import RayBundle, RaySamples
        from nerfstudio.field_components.field_heads import FieldHeadNames

        rb = RayBundle(origins=rays["origins"].clone(), directions=rays["directions"].clone(), pixel_area=pixel_area.clone(),
                       nears=torch.zeros_like(rays["nears"]), fars=rays["fars"].clone().clamp_max(sky_distance),
                       times=rays["times"].clone(), metadata={})
        ray_samples, weights_list, rs_list = self.sampler(rb, self.density_fns, pass_ray_samples=True)
        dist = sky_distance - ray_samples.frustums.ends[..., -1, 0]
        ray_samples.frustums.ends[..., -1, 0] += dist
        ray_samples.deltas[..., -1, 0] += dist
        ray_samples.spacing_ends[..., -1, 0] = 1 - 1e-7
        fo = self.field(ray_samples)
        w = RaySamples.get_weights_and_transmittance_from_alphas(fo[FieldHeadNames.ALPHA], weights_only=True)[..., 0]
        acc = torch.sum(w, dim=-1, keepdim=True)
        w = torch.cat((w[..., :-1], w[..., -1:] + 1 - acc), dim=-1).unsqueeze(-1)
        feats = torch.sum(w * fo[FieldHeadNames.FEATURE], dim=-2)
        w, rs_ns = w[..., :-1, :], ray_samples[..., :-1]
        depth = torch.sum(w * (rs_ns.frustums.starts + rs_ns.frustums.ends) / 2, dim=-2)
        return {"features": feats, "depth": depth, "accumulation": acc, "weights_list": list(weights_list) + [w],
                "ray_samples_list": list(rs_list) + [rs_ns]}


def bench_loss(out) -> Tensor:
    """Upstream loss of the synthetic train step."""
    loss = out["features"].pow(2).mean() + 1e-3 * out["depth"].mean()
    for w in out["weights_list"][:-1]:
        loss = loss + w.pow(2).mean()
    return loss


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_reference(rays: Dict[str, Tensor], pixel_area: Tensor, path: ReferencePath, steps: int, warmup: int,
                   train: bool = True) -> Tuple[float, float, int]:
    """(rays/s, ms per step, threads) of the reference torch path on the host CPU, all cores."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    path.train(train)
    n = rays["origins"].shape[0]

    def step():
        if train:
            for p in path.parameters():
                p.grad = None
            bench_loss(path.forward(rays, pixel_area)).backward()
        else:
            with torch.no_grad():
                path.forward(rays, pixel_area)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt * 1e3, cores
