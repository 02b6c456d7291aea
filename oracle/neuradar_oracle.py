"""CPU oracle for the NeuRadar per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain fp32 torch ops on the CPU, the algorithm of the reference's
`implementation="torch"` path.  It exists to check the CUDA kernels and is
imported only by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py`.  Nothing under `neuradar_b200/` may import it: the product has no CPU path.

Pinning: the reference's own tests hold almost no numeric fixtures for this path, so
the oracle is pinned against outputs of the reference itself, executed in the build container by
`tests/golden/make_golden.py` and committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py`
compares every function below with those vectors (bit-exact for integer outputs, <=1e-6 otherwise)
and with the reference's three known-answer tests (`tests/cameras/test_rays.py:11-30`,
`tests/utils/test_math.py:8-16`, the hash KATs of SURVEY.md 8c).
The nerfacc compositing contract (`render_weight_from_alpha`, third-party, not vendored in the
reference) has no fixture anywhere: its restatement below (`alpha_weights(eps=0)`) is "parity
unpinned" and is anchored on the in-tree twin `RaySamples.get_weights_and_transmittance_from_alphas`
(`alpha_weights(eps=1e-7)`), which IS pinned.

All `file:line` citations are relative to the reference checkout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

PRIME_Y = 2654435761  # nerfstudio/field_components/encodings.py:418
PRIME_Z = 805459861

# Corner order of HashEncoding.pytorch_fwd (encodings.py:436-443); 1 = ceil, 0 = floor per (x, y, z).
CORNERS = (
    (1, 1, 1),
    (1, 0, 1),
    (0, 0, 1),
    (0, 1, 1),
    (1, 1, 0),
    (1, 0, 0),
    (0, 0, 0),
    (0, 1, 0),
)


# --------------------------------------------------------------------------------------------
# H1-H4: multiresolution hash grid
# --------------------------------------------------------------------------------------------
def level_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """Per-level grid resolutions, HashEncoding.__init__ (encodings.py:348-350).

    The growth factor is a float64 numpy scalar but the power is evaluated by torch on an int64
    `arange`, which yields float32 - e.g. the top level of the 32..8192 grid is 8191, not 8192.
    """
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    return torch.floor(min_res * growth**levels)


def hash_coords(coords: Tensor, log2_hashmap_size: int) -> Tensor:
    """Spatial hash of integer grid coordinates `[..., L, 3]` -> `[..., L]` int64 (encodings.py:406-423).

    int32 coordinates times int64 primes, xor, Python-style modulo, plus the per-level offset.
    """
    table_size = 1 << log2_hashmap_size
    num_levels = coords.shape[-2]
    c = coords.to(torch.int64)
    h = c[..., 0] ^ (c[..., 1] * PRIME_Y) ^ (c[..., 2] * PRIME_Z)
    h = torch.remainder(h, table_size)
    return h + torch.arange(num_levels, dtype=torch.int64) * table_size


def hash_corner_indices(x: Tensor, scalings: Tensor, log2_hashmap_size: int) -> Tuple[Tensor, Tensor]:
    """Table rows of the 8 cell corners and the in-cell offsets (encodings.py:428-443).

    Returns `idx [M, L, 8]` int64 (corner order `CORNERS`) and `offset [M, L, 3]` fp32.
    """
    scaled = x[..., None, :] * scalings.view(-1, 1)
    hi = torch.ceil(scaled).to(torch.int32)
    lo = torch.floor(scaled).to(torch.int32)
    offset = scaled - lo
    idx = []
    for cx, cy, cz in CORNERS:
        corner = torch.stack(
            [(hi if cx else lo)[..., 0], (hi if cy else lo)[..., 1], (hi if cz else lo)[..., 2]], dim=-1
        )
        idx.append(hash_coords(corner, log2_hashmap_size))
    return torch.stack(idx, dim=-1), offset


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_hashmap_size: int) -> Tensor:
    """Trilinear interpolation of hashed features, `[M, 3] -> [M, L*F]` (encodings.py:425-466)."""
    idx, o = hash_corner_indices(x, scalings, log2_hashmap_size)
    f = [table[idx[..., k]] for k in range(8)]  # each [M, L, F]
    ox, oy, oz = o[..., 0:1], o[..., 1:2], o[..., 2:3]
    f03 = f[0] * ox + f[3] * (1 - ox)
    f12 = f[1] * ox + f[2] * (1 - ox)
    f56 = f[5] * ox + f[6] * (1 - ox)
    f47 = f[4] * ox + f[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    out = f0312 * oz + f4756 * (1 - oz)
    return out.flatten(-2, -1)


# --------------------------------------------------------------------------------------------
# H5-H7: sample gaussians, contraction, anti-alias level weights
# --------------------------------------------------------------------------------------------
def fast_isotropic_gaussian(
    origins: Tensor, directions: Tensor, starts: Tensor, ends: Tensor, pixel_area: Tensor
) -> Tuple[Tensor, Tensor]:
    """One-multisample gaussian of each frustum (cameras/rays.py:109-124).

    origins/directions `[N, 3]`, pixel_area `[N, 1]`, starts/ends `[N, S]` -> mean `[N, S, 3]`, std `[N, S, 1]`.
    """
    dist = (ends - starts) / 2
    t = starts + 1.0 * dist
    mean = origins[:, None, :] + directions[:, None, :] * t[..., None]
    area = pixel_area[:, None, :] * t[..., None].pow(2)
    std = (area * dist[..., None]).pow(1 / 3)
    return mean, std


def contract(mean: Tensor, std: Tensor, scale: float) -> Tuple[Tensor, Tensor]:
    """L-inf ZipNeRF-style contraction into [0,1]^3 (spatial_distortions.py:103-113,132-136)."""
    mean = mean / scale
    std = std / scale
    mag = torch.linalg.norm(mean, ord=float("inf"), dim=-1)[..., None]
    inside = mag < 1
    cm = mag.clamp_min(1.0)
    mean = torch.where(inside, mean, (2 - (1 / cm)) * (mean / cm))
    std = torch.where(inside, std, std * ((2 * cm - 1).pow(1 / 3) / cm) ** 2)
    return (mean + 2.0) / 4.0, std / 4.0


def level_weights(scalings: Tensor, std: Tensor) -> Tensor:
    """Anti-aliasing down-weighting per level, `[..., 1] -> [..., L]` (neurad_encoding.py:314)."""
    return 1 / (scalings * 2 * std).clamp_min(1.0)


def neurad_hash_encode(
    mean: Tensor, std: Tensor, table: Tensor, scalings: Tensor, log2_hashmap_size: int, static_scale: float
) -> Tensor:
    """Static branch of NeuRADHashEncoding.forward (neurad_encoding.py:152-176,277-280,309-316).

    mean `[N, S, 3]`, std `[N, S, 1]` (world units) -> `[N*S, L*F]`.
    """
    num_levels = scalings.numel()
    cmean, cstd = contract(mean, std, static_scale)
    feats = hash_encode(cmean.reshape(-1, 3), table, scalings, log2_hashmap_size)
    w = level_weights(scalings, cstd.reshape(-1, 1))
    feats = feats.view(feats.shape[0], num_levels, -1) * w[..., None]
    return feats.flatten(-2, -1)


# --------------------------------------------------------------------------------------------
# M1-M4: tiny MLPs, spherical harmonics, the two fields
# --------------------------------------------------------------------------------------------
def mlp(x: Tensor, weights: Sequence[Tensor], biases: Sequence[Optional[Tensor]]) -> Tensor:
    """Linear+ReLU chain without output activation (field_components/mlp.py:159-178)."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = torch.nn.functional.linear(x, w, b)
        if i < n - 1:
            x = torch.relu(x)
    return x


def sh16(d: Tensor) -> Tensor:
    """Degree-4 real spherical harmonics basis, `[M, 3] -> [M, 16]` (utils/math.py:31-94).  No gradient."""
    d = d.detach()
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x**2, y**2, z**2
    out = torch.zeros((*d.shape[:-1], 16))
    out[..., 0] = 0.28209479177387814
    out[..., 1] = 0.4886025119029199 * y
    out[..., 2] = 0.4886025119029199 * z
    out[..., 3] = 0.4886025119029199 * x
    out[..., 4] = 1.0925484305920792 * x * y
    out[..., 5] = 1.0925484305920792 * y * z
    out[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
    out[..., 7] = 1.0925484305920792 * x * z
    out[..., 8] = 0.5462742152960396 * (xx - yy)
    out[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
    out[..., 10] = 2.890611442640554 * x * y * z
    out[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
    out[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
    out[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
    out[..., 14] = 1.445305721320277 * z * (xx - yy)
    out[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return out


class _TruncExp(torch.autograd.Function):
    """exp whose backward clamps the argument to [-15, 15] (field_components/activations.py:28-41)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


@dataclass
class GridParams:
    """One hash grid: `table [L*T, F]`, fp32 `scalings [L]`, log2 of rows per level."""

    table: Tensor
    scalings: Tensor
    log2_hashmap_size: int


@dataclass
class FieldParams:
    """Parameters of NeuRADField (fields/neurad_field.py:79-120) under their torch-path names."""

    grid: GridParams
    geo_w: List[Tensor]
    geo_b: List[Tensor]
    feat_w: List[Tensor]
    feat_b: List[Tensor]
    beta: Tensor
    static_scale: float = 100.0
    beta_min: float = 1e-4


@dataclass
class ProposalParams:
    """Parameters of NeuRADProposalField (fields/neurad_field.py:185-203)."""

    grid: GridParams
    decoder_w: Tensor  # [1, L*F], no bias
    static_scale: float = 100.0


def proposal_density(
    p: ProposalParams, origins: Tensor, directions: Tensor, pixel_area: Tensor, starts: Tensor, ends: Tensor
) -> Tensor:
    """NeuRADProposalField.get_density (fields/neurad_field.py:208-213) -> `[N, S, 1]`."""
    mean, std = fast_isotropic_gaussian(origins, directions, starts, ends, pixel_area)
    feats = neurad_hash_encode(mean, std, p.grid.table, p.grid.scalings, p.grid.log2_hashmap_size, p.static_scale)
    x = torch.nn.functional.linear(feats, p.decoder_w)
    return trunc_exp(x).view(*starts.shape, 1)


def field_forward(
    p: FieldParams, origins: Tensor, directions: Tensor, pixel_area: Tensor, starts: Tensor, ends: Tensor
) -> Dict[str, Tensor]:
    """NeuRADField.forward (fields/neurad_field.py:128-152) -> FEATURE `[N,S,32]`, SDF, ALPHA `[N,S,1]`."""
    n, s = starts.shape
    mean, std = fast_isotropic_gaussian(origins, directions, starts, ends, pixel_area)
    feats = neurad_hash_encode(mean, std, p.grid.table, p.grid.scalings, p.grid.log2_hashmap_size, p.static_scale)
    geo = mlp(feats, p.geo_w, p.geo_b)
    sdf, emb = torch.split(geo, [1, geo.shape[-1] - 1], dim=-1)
    dirs = directions[:, None, :].expand(n, s, 3).reshape(-1, 3)
    dir_emb = sh16((dirs + 1.0) / 2.0)  # base_field.py:136-142
    feature = emb + mlp(torch.cat([emb, dir_emb], dim=-1), p.feat_w, p.feat_b)
    sdf = sdf.view(n, s, 1)
    alpha = torch.sigmoid(-sdf * (p.beta.abs() + p.beta_min))  # model_components/utils.py:30-41
    return {"feature": feature.view(n, s, -1), "sdf": sdf, "alpha": alpha}


# --------------------------------------------------------------------------------------------
# S1-S3: samplers
# --------------------------------------------------------------------------------------------
def power_fn(x: Tensor, lam: float) -> Tensor:
    """ZipNeRF power transform for finite lam not in {0, 1} (utils/math.py:541-558)."""
    lam_1 = abs(lam - 1)
    return (lam_1 / lam) * ((x / lam_1 + 1) ** lam - 1)


def inv_power_fn(x: Tensor, lam: float, eps: float = 1e-10) -> Tensor:
    """Inverse of `power_fn` (utils/math.py:561-580)."""
    lam_1 = abs(lam - 1)
    return ((x * lam / lam_1 + 1).clamp_min(eps) ** (1 / lam) - 1) * lam_1


@dataclass
class Spacing:
    """The `spacing_to_euclidean_fn` closure of SpacedSampler (ray_samplers.py:117-120) as data."""

    s_near: Tensor  # [N, 1]
    s_far: Tensor  # [N, 1]
    lam: float
    scaling: float

    def to_euclidean(self, x: Tensor) -> Tensor:
        return inv_power_fn(x * self.s_far + (1 - x) * self.s_near, self.lam) / self.scaling


def spaced_bins(
    nears: Tensor, fars: Tensor, num_samples: int, jitter: Optional[Tensor], lam: float = -1.0, scaling: float = 0.1
) -> Tuple[Tensor, Tensor, Spacing]:
    """PowerSampler initial bins (ray_samplers.py:80-132,838-852).

    `jitter [N, 1]` is the `torch.rand` draw of training mode with single_jitter, or None in eval.
    Returns spacing bins `[N, S+1]` (or `[1, S+1]` in eval), euclidean bins `[N, S+1]`, the closure data.
    """
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, ...]
    if jitter is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * jitter
    sp = Spacing(power_fn(nears * scaling, lam), power_fn(fars * scaling, lam), lam, scaling)
    return bins, sp.to_euclidean(bins), sp


def pdf_cdf(weights: Tensor, padding: float = 0.01, eps: float = 1e-5) -> Tensor:
    """Piecewise-constant CDF `[N, S] -> [N, S+1]` (ray_samplers.py:308-319)."""
    w = weights + padding
    total = torch.sum(w, dim=-1, keepdim=True)
    pad = torch.relu(eps - total)
    w = w + pad / w.shape[-1]
    total = total + pad
    pdf = w / total
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def pdf_u(num_rays: int, num_samples: int, jitter: Optional[Tensor]) -> Tensor:
    """Stratified CDF query points `[N, S_out+1]` (ray_samplers.py:321-335); jitter `[N,1]` = torch.rand draw."""
    nb = num_samples + 1
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb).expand(num_rays, nb).clone()
    if jitter is not None:
        return (u + jitter / nb).contiguous()
    return (u + 1.0 / (2 * nb)).contiguous()


def pdf_invert(cdf: Tensor, u: Tensor, existing_bins: Tensor) -> Tuple[Tensor, Tensor]:
    """Inverse-CDF lookup (ray_samplers.py:349-359) -> new spacing bins `[N, S_out+1]`, `inds` int64."""
    inds = torch.searchsorted(cdf, u, side="right")
    last = existing_bins.shape[-1] - 1
    below = torch.clamp(inds - 1, 0, last)
    above = torch.clamp(inds, 0, last)
    cdf0 = torch.gather(cdf, -1, below)
    bin0 = torch.gather(existing_bins, -1, below)
    cdf1 = torch.gather(cdf, -1, above)
    bin1 = torch.gather(existing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf0) / (cdf1 - cdf0), 0), 0, 1)
    return (bin0 + t * (bin1 - bin0)).detach(), inds


def pdf_sample(
    weights: Tensor, existing_bins: Tensor, num_samples: int, jitter: Optional[Tensor]
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """PDFSampler.generate_ray_samples with include_original=False (ray_samplers.py:280-376).

    weights `[N, S_in]`, existing spacing bins `[N, S_in+1]` -> (bins `[N, S_out+1]`, inds, cdf, u).
    """
    cdf = pdf_cdf(weights)
    u = pdf_u(weights.shape[0], num_samples, jitter)
    bins, inds = pdf_invert(cdf, u, existing_bins.expand(weights.shape[0], -1))
    return bins, inds, cdf, u


# --------------------------------------------------------------------------------------------
# C1-C6: compositing, renderers, point heads
# --------------------------------------------------------------------------------------------
def density_weights(densities: Tensor, deltas: Tensor) -> Tensor:
    """RaySamples.get_weights (cameras/rays.py:188-210); `[N, S, 1]` in and out."""
    dd = deltas * densities
    alphas = 1 - torch.exp(-dd)
    trans = torch.cumsum(dd[..., :-1, :], dim=-2)
    trans = torch.cat([torch.zeros((*trans.shape[:1], 1, 1)), trans], dim=-2)
    return torch.nan_to_num(alphas * torch.exp(-trans))


def alpha_weights(alphas: Tensor, eps: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Weights and transmittance from alphas, `[N, S] -> ([N, S], [N, S+1])`.

    eps=1e-7: RaySamples.get_weights_and_transmittance_from_alphas (cameras/rays.py:226-248).
    eps=0:    the dense-input contract of nerfacc.render_weight_from_alpha as called at
              models/neuradar.py:1016 (nerfacc==0.5.2, pyproject.toml:36; third party, parity unpinned).
    """
    ones = torch.ones((alphas.shape[0], 1))
    trans = torch.cumprod(torch.cat([ones, 1.0 - alphas + eps], dim=1), dim=1)
    return alphas * trans[:, :-1], trans


def composite(
    alphas: Tensor, features: Tensor, starts: Tensor, ends: Tensor, eps: float = 0.0
) -> Dict[str, Tensor]:
    """Compositing tail of get_nff_outputs (models/neuradar.py:504-517, models/neurad.py:721-728).

    alphas `[N, S]`, features `[N, S, C]`, starts/ends `[N, S]` (last sample = sky sample).
    """
    w, _ = alpha_weights(alphas, eps)
    acc = torch.sum(w[..., None], dim=-2)  # AccumulationRenderer, renderers.py:349
    w = torch.cat((w[..., :-1], w[..., -1:] + 1 - acc), dim=-1).unsqueeze(-1)
    feat = torch.sum(features * w, dim=-2)  # FeatureRenderer, renderers.py:85
    w = w[..., :-1, :]
    steps = (starts[..., :-1, None] + ends[..., :-1, None]) / 2
    depth = torch.sum(w * steps, dim=-2)  # accumulate_along_rays(weights, steps)
    return {"features": feat, "depth": depth, "accumulation": acc, "weights": w}


def expected_depth(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """DepthRenderer(method="expected"), dense branch (renderers.py:399-414); weights `[N,S,1]`."""
    steps = (starts[..., None] + ends[..., None]) / 2
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
    return torch.clip(depth, steps.min(), steps.max())


def median_depth(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """DepthRenderer(method="median") (renderers.py:386-397)."""
    steps = (starts + ends) / 2
    cw = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * 0.5
    idx = torch.clamp(torch.searchsorted(cw, split, side="left"), 0, steps.shape[-1] - 1)
    return torch.gather(steps, dim=-1, index=idx)


def lidar_points(origins: Tensor, directions: Tensor, depth: Tensor) -> Tensor:
    """Lidar point head before the sensor-frame transform (models/ad_model.py:105)."""
    return origins + directions * depth


def radar_points(depth: Tensor, theta: Tensor, phi: Tensor) -> Tensor:
    """Radar point head `_get_cartesian_coords` (models/neuradar.py:469-473,1025-1029); all `[N, 1]`."""
    x = depth * torch.cos(phi) * torch.cos(theta)
    y = depth * torch.sin(phi) * torch.cos(theta)
    z = depth * torch.sin(theta)
    return torch.cat((x, y, z), dim=-1)


def is_close_to_lidar(starts: Tensor, ends: Tensor, is_lidar: Tensor, directions_norm: Tensor,
                      did_return: Optional[Tensor] = None, carving_epsilon: float = 0.1,
                      non_return_lidar_distance: float = 150.0) -> Tensor:
    """`NeuRadarModel._compute_is_close_to_lidar` (models/neuradar.py:971-994) restated densely: starts / ends [N,S],
    is_lidar / did_return [N,1] bool, directions_norm [N,1]; returns bool [N,S] (False for every non-lidar ray)."""
    mid = (starts + ends) * 0.5
    close_to_hit = (directions_norm - mid).abs() < carving_epsilon
    if did_return is not None:
        in_range = mid < non_return_lidar_distance
        close = (did_return & close_to_hit) | ((~did_return) & in_range)
    else:
        close = close_to_hit
    return close & is_lidar


def carving_loss(weights: Tensor, close: Tensor, is_lidar: Tensor) -> Tensor:
    """`((prop_w * weights_mask) ** 2).sum()` with weights_mask = ~is_close_to_lidar & is_lidar (models/neuradar.py:529-531)."""
    return ((weights * ((~close) & is_lidar)) ** 2).sum()


# --------------------------------------------------------------------------------------------
# the whole path: NeuRadarModel._get_ray_samples + get_nff_outputs
# --------------------------------------------------------------------------------------------
@dataclass
class PathConfig:
    """Sampling settings (models/neuradar.py:120-138)."""

    num_proposal_samples: Tuple[int, ...] = (128, 64)
    num_nerf_samples: int = 32
    power_lambda: float = -1.0
    power_scaling: float = 0.1
    sky_distance: float = 20000.0
    anneal: float = 1.0


@dataclass
class PathOutputs:
    features: Tensor  # [N, C]
    depth: Tensor  # [N, 1]
    accumulation: Tensor  # [N, 1]
    weights_list: List[Tensor] = field(default_factory=list)  # per round [N, S, 1]
    sbins_list: List[Tensor] = field(default_factory=list)  # spacing bins [N, S+1]
    ebins_list: List[Tensor] = field(default_factory=list)  # euclidean bins [N, S+1]
    inds_list: List[Tensor] = field(default_factory=list)  # searchsorted indices of each PDF round
    prop_depths: List[Tensor] = field(default_factory=list)


def nff_forward(
    fld: FieldParams,
    proposals: Sequence[ProposalParams],
    origins: Tensor,
    directions: Tensor,
    pixel_area: Tensor,
    nears: Tensor,
    fars: Tensor,
    cfg: PathConfig,
    jitters: Optional[Sequence[Tensor]],
    composite_eps: float = 0.0,
) -> PathOutputs:
    """`_get_ray_samples` + `get_nff_outputs` without the appearance embedding
    (models/neuradar.py:495-548,570-586; ray_samplers.py:623-666).

    `proposals[i]` is the field queried in round i.  The reference model's late-binding lambda list
    (neuradar.py:302) makes both rounds query `proposal_fields[1]`; callers reproduce that by passing
    the same object twice.  `jitters` holds one `[N,1]` uniform draw per round (training) or is None (eval).
    `pixel_area` is expected already scaled by `_scale_pixel_area`.
    """
    fars = fars.clamp_max(cfg.sky_distance)
    n_rounds = len(cfg.num_proposal_samples)
    out_w, out_sb, out_eb, out_inds, out_pd = [], [], [], [], []
    weights = None
    sbins = ebins = None
    spacing = None
    for level in range(n_rounds + 1):
        is_prop = level < n_rounds
        s = cfg.num_proposal_samples[level] if is_prop else cfg.num_nerf_samples
        jit = None if jitters is None else jitters[level]
        if level == 0:
            sbins, ebins, spacing = spaced_bins(nears, fars, s, jit, cfg.power_lambda, cfg.power_scaling)
            sbins = sbins.expand(origins.shape[0], -1)
        else:
            annealed = torch.pow(weights, cfg.anneal)
            sbins, inds, _, _ = pdf_sample(annealed[..., 0], sbins, s, jit)
            ebins = spacing.to_euclidean(sbins)
            out_inds.append(inds)
        if is_prop:
            starts, ends = ebins[:, :-1], ebins[:, 1:]
            dens = proposal_density(proposals[level], origins, directions, pixel_area, starts, ends)
            weights = density_weights(dens, (ends - starts)[..., None])
            out_w.append(weights)
            out_sb.append(sbins)
            out_eb.append(ebins)
            out_pd.append(torch.sum(weights * ((starts + ends) / 2)[..., None], dim=-2))
    # sky sample: the final bin is stretched to sky_distance (neuradar.py:578-582)
    last = ebins[:, -1:]
    ebins = torch.cat([ebins[:, :-1], last + (cfg.sky_distance - last)], dim=-1)
    sbins = torch.cat([sbins[:, :-1], torch.full_like(sbins[:, -1:], 1 - 1e-7)], dim=-1)
    starts, ends = ebins[:, :-1], ebins[:, 1:]
    f = field_forward(fld, origins, directions, pixel_area, starts, ends)
    comp = composite(f["alpha"][..., 0], f["feature"], starts, ends, composite_eps)
    out_w.append(comp["weights"])
    out_sb.append(sbins[:, :-1])
    out_eb.append(ebins[:, :-1])
    return PathOutputs(
        comp["features"], comp["depth"], comp["accumulation"], out_w, out_sb, out_eb, out_inds, out_pd
    )


def bench_loss(out: PathOutputs) -> Tensor:
    """Upstream loss of the synthetic train step: touches every output of the path."""
    loss = out.features.pow(2).mean() + 1e-3 * out.depth.mean()
    for w in out.weights_list[:-1]:
        loss = loss + w.pow(2).mean()
    return loss


# --------------------------------------------------------------------------------------------
# H8: dynamic actors (field_components/neurad_encoding.py:152-307)
# --------------------------------------------------------------------------------------------
def pose_inverse(pose: Tensor) -> Tensor:
    """[..., 3|4, 4] -> [..., 3, 4] inverse of a rigid transform (utils/poses.py:42-55)."""
    R = pose[..., :3, :3]
    t = pose[..., :3, 3:]
    Ri = R.transpose(-2, -1)
    return torch.cat([Ri, -Ri.matmul(t)], dim=-1)


def transform_points(points: Tensor, transforms: Tensor, with_translation: bool = True) -> Tensor:
    """cameras/lidars.py:507-519."""
    out = (points.unsqueeze(-2) @ transforms[..., :3, :3].swapaxes(-2, -1)).squeeze(-2)
    return out + transforms[..., :3, 3] if with_translation else out


def actor_sample_indices(mean: Tensor, boxes2world: Tensor, valid: Tensor, bounds: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(ray, sample, actor) triples of the samples inside actor boxes (neurad_encoding.py:231-275).

    mean [N,S,3] world-space sample centres, boxes2world [N,A,4,4] at each ray's time, valid [N,A], bounds [A,3]."""
    eps = 1.0e-7
    radii = bounds.norm(dim=-1)
    p0 = mean[:, 0, :]
    line = mean[:, -1, :] - p0
    line = (line / (torch.linalg.norm(line, dim=-1, keepdim=True) + eps)).unsqueeze(-2)
    centers = boxes2world[..., :3, 3]
    dist = torch.linalg.norm(torch.cross(centers - p0.unsqueeze(-2), line.expand_as(centers), dim=-1), dim=-1)
    r, a = ((dist < radii) & valid).nonzero(as_tuple=False).T
    empty = torch.empty(0, dtype=torch.int64)
    if r.shape[0] == 0:
        return empty, empty, empty
    within = torch.linalg.norm(mean[r] - centers[r, a].unsqueeze(-2), dim=-1) < radii[a].unsqueeze(-1)
    k, s = within.nonzero(as_tuple=False).T
    ri, ai = r[k], a[k]
    w2b = pose_inverse(boxes2world)
    inside = (transform_points(mean[ri, s], w2b[ri, ai]).abs() < bounds[ai]).all(dim=-1)
    return ri[inside], s[inside], ai[inside]


def neurad_hash_encode_with_actors(
    mean: Tensor, std: Tensor, directions: Tensor, static: GridParams, actor_grids: Sequence[GridParams],
    static_scale: float, actor_scale: float, boxes2world: Tensor, valid: Tensor, bounds: Tensor, actor_to_id: Tensor,
    ray_flip: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    """NeuRADHashEncoding.forward with per-actor 3-D grids (the torch path, neurad_encoding.py:152-229,295-316).

    mean [N,S,3], std [N,S,1] world-space gaussians, directions [N,S,3]; ray_flip [N] of +-1 (training) or None.
    Returns features [N*S, L*F] and the per-sample directions (actor samples rotated into the box frame)."""
    n, s, _ = mean.shape
    feats = neurad_hash_encode(mean, std, static.table, static.scalings, static.log2_hashmap_size, static_scale)
    ri, si, ai = actor_sample_indices(mean, boxes2world, valid, bounds)
    directions = directions.clone()
    if ri.shape[0] == 0:
        return feats, directions
    w2b = pose_inverse(boxes2world)[ri, ai]
    pos = transform_points(mean[ri, si], w2b)
    d = transform_points(directions[ri, si], w2b, with_translation=False)
    d = d / (torch.linalg.norm(d, dim=-1, keepdim=True) + 1.0e-7)
    if ray_flip is not None:
        f = ray_flip[ri]
        pos = torch.cat([pos[:, :1] * f[:, None], pos[:, 1:]], dim=-1)
        d = torch.cat([d[:, :1] * f[:, None], d[:, 1:]], dim=-1)
    directions[ri, si] = d
    cpos, cstd = contract(pos, std[ri, si], actor_scale)
    out_dim = feats.shape[-1]
    grid_id = actor_to_id[ai]
    actor_feats = torch.zeros((ri.shape[0], out_dim))
    for gid in grid_id.unique():
        g = actor_grids[int(gid)]
        m = grid_id == gid
        f = hash_encode(cpos[m], g.table, g.scalings, g.log2_hashmap_size)
        w = level_weights(g.scalings, cstd[m])
        f = (f.view(f.shape[0], g.scalings.numel(), -1) * w[..., None]).flatten(-2, -1)
        actor_feats = actor_feats.index_put((m.nonzero(as_tuple=True)[0],), torch.nn.functional.pad(f, (0, out_dim - f.shape[-1])))
    feats = feats.view(n, s, out_dim).index_put((ri, si), actor_feats)
    return feats.view(n * s, out_dim), directions


def field_forward_with_actors(p: FieldParams, actor_grids: Sequence[GridParams], actor_scale: float, origins: Tensor,
                              directions: Tensor, pixel_area: Tensor, starts: Tensor, ends: Tensor, boxes2world: Tensor,
                              valid: Tensor, bounds: Tensor, actor_to_id: Tensor, ray_flip: Optional[Tensor] = None):
    """NeuRADField.forward (fields/neurad_field.py:128-152) in a scene with dynamic actors."""
    n, s = starts.shape
    mean, std = fast_isotropic_gaussian(origins, directions, starts, ends, pixel_area)
    dirs = directions[:, None, :].expand(n, s, 3)
    feats, dirs = neurad_hash_encode_with_actors(mean, std, dirs, p.grid, actor_grids, p.static_scale, actor_scale,
                                                 boxes2world, valid, bounds, actor_to_id, ray_flip)
    geo = mlp(feats, p.geo_w, p.geo_b)
    sdf, emb = torch.split(geo, [1, geo.shape[-1] - 1], dim=-1)
    dir_emb = sh16((dirs.reshape(-1, 3) + 1.0) / 2.0)
    feature = emb + mlp(torch.cat([emb, dir_emb], dim=-1), p.feat_w, p.feat_b)
    sdf = sdf.view(n, s, 1)
    alpha = torch.sigmoid(-sdf * (p.beta.abs() + p.beta_min))
    return {"feature": feature.view(n, s, -1), "sdf": sdf, "alpha": alpha, "grid_features": feats, "directions": dirs}


# --------------------------------------------------------------------------------------------
# losses that run on the path's outputs every training step
# --------------------------------------------------------------------------------------------
def distortion_loss(sdist: Tensor, w: Tensor) -> Tensor:
    """MipNeRF-360 distortion loss of the final level, mean over rays (model_components/losses.py:137-156).

    sdist [N, S+1] spacing-domain bin edges, w [N, S] weights."""
    ut = (sdist[..., 1:] + sdist[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    intra = torch.sum(w**2 * (sdist[..., 1:] - sdist[..., :-1]), dim=-1) / 3
    return torch.mean(inter + intra)


def _blur_stepfun(x: Tensor, y: Tensor, r: float) -> Tuple[Tensor, Tensor]:
    """Convolve the step function (x, y) with a box of half-width r -> piecewise linear (losses.py:620-629).

    Note: on the host torch.cumsum carries its running sum in fp64 (rounded to fp32 per element); the two nested
    prefix sums cancel heavily, so the CUDA kernel carries in fp64 too (csrc/losses.cu)."""
    xr, order = torch.sort(torch.cat([x - r, x + r], dim=-1))
    zero = torch.zeros_like(y[..., :1])
    y1 = (torch.cat([y, zero], dim=-1) - torch.cat([zero, y], dim=-1)) / (2 * r)
    y2 = torch.cat([y1, -y1], dim=-1).take_along_dim(order[..., :-1], dim=-1)
    yr = torch.cumsum((xr[..., 1:] - xr[..., :-1]) * torch.cumsum(y2, dim=-1), dim=-1).clamp_min(0)
    return xr, torch.cat([torch.zeros_like(yr[..., :1]), yr], dim=-1)


def _interp_quad(x: Tensor, xp: Tensor, fpdf: Tensor, fcdf: Tensor) -> Tensor:
    """Piecewise-quadratic CDF of a piecewise-linear PDF evaluated at sorted queries (losses.py:632-645)."""
    right = torch.searchsorted(xp, x)
    left = (right - 1).clamp_min(0)
    right = right.clamp_max(xp.shape[-1] - 1)
    xp0, xp1 = xp.take_along_dim(left, dim=-1), xp.take_along_dim(right, dim=-1)
    p0, p1 = fpdf.take_along_dim(left, dim=-1), fpdf.take_along_dim(right, dim=-1)
    c0 = fcdf.take_along_dim(left, dim=-1)
    off = torch.clip(torch.nan_to_num((x - xp0) / (xp1 - xp0), 0), 0, 1)
    return c0 + (x - xp0) * (p0 + p1 * off + p0 * (1 - off)) * 0.5


def zipnerf_interlevel_loss(c: Tensor, w: Tensor, proposals: Sequence[Tuple[Tensor, Tensor]],
                            pulse_widths: Sequence[float] = (0.03, 0.003)) -> Tensor:
    """Anti-aliased interlevel loss (losses.py:648-705).

    c [N, S+1], w [N, S]: spacing bins and weights of the FINAL level (treated as constants);
    proposals: (cp [N, Sp+1], wp [N, Sp]) per proposal round; gradients flow to wp only."""
    c, w = c.detach(), w.detach()
    w = torch.cat([w[..., :-1], w[..., -1:] + (1 - torch.sum(w, dim=-1, keepdim=True))], dim=-1)
    w_norm = w / (c[..., 1:] - c[..., :-1])
    loss = 0
    for (cp, wp), r in zip(proposals, pulse_widths):
        c_, w_ = _blur_stepfun(c, w_norm, r)
        area = 0.5 * (w_[..., 1:] + w_[..., :-1]) * (c_[..., 1:] - c_[..., :-1])
        cdf = torch.cat([torch.zeros_like(area[..., :1]), torch.cumsum(area, dim=-1)], dim=-1)
        zero, one = torch.zeros_like(c_[..., :1]), torch.ones_like(c_[..., :1])
        c_ = torch.cat([zero, c_, one], dim=-1)
        w_ = torch.cat([zero, w_, zero], dim=-1)
        cdf = torch.cat([zero, cdf, one], dim=-1)
        w_s = torch.diff(_interp_quad(cp, c_, w_, cdf), dim=-1)
        loss = loss + ((w_s - wp).clamp_min(0) ** 2 / (wp + 1e-5)).sum(dim=-1).mean()
    return loss


def training_losses(out: PathOutputs, interlevel_mult: float = 0.001, distortion_mult: float = 0.002) -> Tensor:
    """The two sampler regularisers of NeuRadarModel.get_loss_dict (models/neurad.py:524-545, defaults :83-85)."""
    w_final = out.weights_list[-1][..., 0]
    proposals = [(sb, w[..., 0]) for sb, w in zip(out.sbins_list[:-1], out.weights_list[:-1])]
    inter = zipnerf_interlevel_loss(out.sbins_list[-1], w_final, proposals)
    return interlevel_mult * inter + distortion_mult * distortion_loss(out.sbins_list[-1], w_final)


# --------------------------------------------------------------------------------------------
# the optimiser step of the "hashgrids" / "fields" parameter groups
# --------------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, betas: Tuple[float, float] = (0.9, 0.999),
              eps: float = 1e-15, weight_decay: float = 0.0, decoupled: bool = False, grad_mult: float = 1.0,
              loss_scale: float = 1.0) -> None:
    """One in-place Adam (decoupled=False) / AdamW (True) update as torch.optim runs it for the reference
    (configs/method_configs.py:393-400 -> engine/optimizers.py:47-52,159-181 -> torch/optim/adam.py
    _single_tensor_adam).  `step` is the 1-based count of this update; the gradient is first multiplied by
    grad_mult / loss_scale (data-parallel average, GradScaler.unscale_)."""
    g = g * (grad_mult / loss_scale)
    if decoupled:
        p.mul_(1 - lr * weight_decay)
    elif weight_decay != 0:
        g = g.add(p, alpha=weight_decay)
    m.lerp_(g, 1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# ------------------------------------------------------------------------------------------------
# radar ray generation
# ------------------------------------------------------------------------------------------------
def radar_rays(radar_to_worlds: Tensor, min_az: Tensor, max_az: Tensor, az_step: Tensor, min_el: Tensor, max_el: Tensor,
               el_step: Tensor, scan_indices: Tensor) -> Dict[str, Tensor]:
    """Radars._generate_rays_from_fov (nerfstudio/cameras/radars.py:268-357), scan by scan like the reference:
    arange x arange meshgrid ("ij") of azimuth / elevation, spherical -> cartesian (:312-315), rotate + translate by the
    scan's radar_to_world (cameras/lidars.py:507-519), subtract the origin, normalize_with_norm (camera_utils.py:596-610),
    pixel_area = (azimuth_step / 5)(elevation_step / 5) (:322-328)."""
    eps = torch.finfo(torch.float32).eps
    spher, scans = [], []
    for idx in scan_indices.tolist():
        az = torch.arange(float(min_az[idx]), float(max_az[idx]), float(az_step[idx]))
        el = torch.arange(float(min_el[idx]), float(max_el[idx]), float(el_step[idx]))
        ga, ge = torch.meshgrid(az, el, indexing="ij")
        spher.append(torch.stack((ga.flatten(), ge.flatten()), dim=1))
        scans.append(torch.full((ga.numel(),), idx, dtype=torch.int64))
    spher = torch.cat(spher) if spher else torch.zeros((0, 2))
    scans = torch.cat(scans) if scans else torch.zeros((0,), dtype=torch.int64)
    r2w = radar_to_worlds[scans]
    origins = r2w[:, :3, 3]
    d = torch.stack([torch.cos(spher[:, 1]) * torch.cos(spher[:, 0]), torch.cos(spher[:, 1]) * torch.sin(spher[:, 0]),
                     torch.sin(spher[:, 1])], dim=-1)
    p = (d.unsqueeze(-2) @ r2w[:, :3, :3].swapaxes(-2, -1)).squeeze(-2) + origins
    v = p - origins
    norm = torch.maximum(torch.linalg.vector_norm(v, dim=-1, keepdim=True), torch.tensor([eps]))
    pixel_area = (az_step[scans] / 5) * (el_step[scans] / 5)
    return {"origins": origins, "directions": v / norm, "pixel_area": pixel_area.reshape(-1, 1), "directions_spher": spher,
            "directions_norm": norm, "ray_scan": scans}
