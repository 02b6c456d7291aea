"""Helpers shared by the GPU parity tests, `__graft_entry__.smoke()` and bench.py's CPU baseline: synthetic rays of
the BASELINE shapes, oracle <-> module parameter plumbing, and a whole-path parity run."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from neuradar_b200.synthetic import build_hot_path, scaled_pixel_area, synthetic_rays  # noqa: F401 (re-exported)
from oracle import neuradar_oracle as O


def make_ray_bundle(rays: Dict[str, Tensor], device: str):
    from neuradar_b200 import RayBundle

    return RayBundle(
        origins=rays["origins"].to(device).clone(),
        directions=rays["directions"].to(device).clone(),
        pixel_area=rays["pixel_area"].to(device).clone(),
        nears=rays["nears"].to(device).clone(),
        fars=rays["fars"].to(device).clone(),
        times=rays["times"].to(device).clone(),
        metadata={"is_lidar": rays["is_lidar"].to(device), "is_radar": rays["is_radar"].to(device)},
    )


def oracle_params(model) -> Tuple[O.FieldParams, List[O.ProposalParams], List[Tensor]]:
    """CPU copies of a NeuRadarHotPath's parameters in oracle form, plus the flat list of leaf tensors."""
    f = model.field
    sg = f.hashgrid.static_grid

    def leaf(t):
        return t.detach().cpu().clone().requires_grad_(True)

    fld = O.FieldParams(
        grid=O.GridParams(leaf(sg.hash_table), sg.scalings.detach().cpu().clone(), sg.log2_hashmap_size),
        geo_w=[leaf(l.weight) for l in f.mlp_geo.layers],
        geo_b=[leaf(l.bias) for l in f.mlp_geo.layers],
        feat_w=[leaf(l.weight) for l in f.mlp_feature.layers],
        feat_b=[leaf(l.bias) for l in f.mlp_feature.layers],
        beta=leaf(f.sdf_to_density.beta),
        static_scale=f.hashgrid.static_scale,
    )
    props = []
    for p in model.proposal_fields:
        g = p.hashgrid.static_grid
        props.append(
            O.ProposalParams(
                O.GridParams(leaf(g.hash_table), g.scalings.detach().cpu().clone(), g.log2_hashmap_size),
                leaf(p.density_decoder.weight),
                p.hashgrid.static_scale,
            )
        )
    return fld, props


def named_leaves(fld: O.FieldParams, props: Sequence[O.ProposalParams]) -> Dict[str, Tensor]:
    """Oracle leaves keyed by the module's parameter names."""
    out = {"field.hashgrid.static_grid.hash_table": fld.grid.table, "field.sdf_to_density.beta": fld.beta}
    for k, (w, b) in enumerate(zip(fld.geo_w, fld.geo_b)):
        out[f"field.mlp_geo.layers.{k}.weight"], out[f"field.mlp_geo.layers.{k}.bias"] = w, b
    for k, (w, b) in enumerate(zip(fld.feat_w, fld.feat_b)):
        out[f"field.mlp_feature.layers.{k}.weight"], out[f"field.mlp_feature.layers.{k}.bias"] = w, b
    for i, p in enumerate(props):
        out[f"proposal_fields.{i}.hashgrid.static_grid.hash_table"] = p.grid.table
        out[f"proposal_fields.{i}.density_decoder.weight"] = p.decoder_w
    return out


class FixedJitter:
    """Context manager that makes torch.rand return pre-drawn tensors (moved to the requested device), so that the
    CUDA samplers and the CPU oracle consume identical jitter."""

    def __init__(self, draws: Sequence[Tensor]):
        self.draws = list(draws)
        self._orig = None

    def __enter__(self):
        self._orig = torch.rand
        it = iter(self.draws)

        def fake_rand(*size, **kw):
            t = next(it)
            shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
            assert tuple(t.shape) == shape, (t.shape, shape)
            return t.to(kw.get("device", "cpu"))

        torch.rand = fake_rand
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig
        return False


def rel_err(a: Tensor, b: Tensor) -> float:
    """max |a-b| relative to max |b| (scale-relative; grads of untouched table rows are exactly 0 in both)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def outside_bar(a: Tensor, b: Tensor, rel: float = 1e-3) -> int:
    """Number of elements that break the element-wise parity bar of north_star: |a - b| <= rel * |b| + rel * rms(b)."""
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    rms = float(b.pow(2).mean().sqrt())
    return int(((a - b).abs() > rel * b.abs() + rel * rms).sum())


def run_path_parity(num_rays: int = 256, device: str = "cuda:0", seed: int = 3, train: bool = True,
                    log2_main: int = 14, log2_prop: int = 14, tol: float = 1e-3, with_losses: bool = False) -> Dict[str, object]:
    """Run the CUDA hot path and the CPU oracle on the same rays, parameters and jitter; compare everything.
    with_losses adds the interlevel + distortion regularisers (at 100x their default multipliers so that their
    gradients are not hidden under the synthetic loss's)."""
    from neuradar_b200 import bench_loss, training_losses

    S0, S1, S2 = 64, 48, 48
    model = build_hot_path(log2_main=log2_main, log2_prop=log2_prop, num_proposal_samples=(S0, S1), num_nerf_samples=S2,
                           late_binding=True, table_gain=(300.0, 2000.0), seed=seed, device=device)
    model.train(train)
    rays = synthetic_rays(num_rays, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    jit = [torch.rand((num_rays, S0 + 1), generator=g), torch.rand((num_rays, 1), generator=g),
           torch.rand((num_rays, 1), generator=g)]
    fld, props = oracle_params(model)
    props_used = [props[-1], props[-1]]  # late binding: both rounds query the last proposal field

    rb = make_ray_bundle(rays, device)
    if train:
        with FixedJitter(jit):
            out = model(rb)
    else:
        with torch.no_grad():
            out = model(rb)
    cfg = O.PathConfig(num_proposal_samples=(S0, S1), num_nerf_samples=S2)
    ref = O.nff_forward(fld, props_used, rays["origins"], rays["directions"], scaled_pixel_area(rays), rays["nears"],
                        rays["fars"], cfg, jit if train else None, composite_eps=0.0)
    report: Dict[str, object] = {}
    report["features"] = rel_err(out["features"], ref.features)
    report["depth"] = rel_err(out["depth"], ref.depth)
    report["accumulation"] = rel_err(out["accumulation"], ref.accumulation)
    if train:
        for i in range(3):
            report[f"weights_{i}"] = rel_err(out["weights_list"][i], ref.weights_list[i])
        loss = bench_loss(out)
        ref_loss = O.bench_loss(ref)
        if with_losses:
            reg, ref_reg = training_losses(out, 0.1, 0.2), O.training_losses(ref, 0.1, 0.2)
            report["regularisers"] = abs(reg.item() - ref_reg.item()) / abs(ref_reg.item())
            loss, ref_loss = loss + reg, ref_loss + ref_reg
        loss.backward()
        ref_loss.backward()
        report["loss"] = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
        leaves = named_leaves(fld, props)
        for name, p in model.named_parameters():
            ref_g = leaves[name].grad
            if ref_g is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                continue
            report["grad:" + name] = rel_err(p.grad, ref_g)
    report["ok"] = all(v <= tol for k, v in report.items() if isinstance(v, float))
    report["worst"] = max((v, k) for k, v in report.items() if isinstance(v, float))
    return report
