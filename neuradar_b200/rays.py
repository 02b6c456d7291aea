"""Exchange types of the hot path: GaussiansStd, Frustums, RaySamples, RayBundle.

Same field names, shapes and broadcast-view semantics as the reference containers
(nerfstudio/cameras/rays.py:33-357 on nerfstudio/utils/tensor_dataclass.py:28-391, nerfstudio/utils/math.py:114-145):
the batch shape is every dimension but the last, fields are broadcast to it as stride-0 views, and indexing /
reshaping acts on the batch dimensions of every tensor field.  The arithmetic methods (`get_weights`,
`get_weights_and_transmittance_from_alphas`, `get_fast_isotropic_gaussian`) run the CUDA kernels.

Samplers in this package additionally attach the un-broadcast per-ray tensors and the [N, S+1] bin tensors the
samples were cut from (`RaySamples.ray_data`, `.euclidean_bins`, `.spacing_bins`), so that kernels read 40 bytes per
ray instead of materialising [N, S, 3] positions.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple, Union

import torch
from torch import Tensor

from . import functional as F


# ------------------------------------------------------------------------------------------------
# Duck-typed accessors: the kernels need the un-broadcast per-ray tensors.  These work on this package's containers
# AND on the reference's own nerfstudio.cameras.rays.{RayBundle,RaySamples,Frustums} (same attribute names), which is
# what makes the modules drop in under the unmodified NeuRadarModel.
# ------------------------------------------------------------------------------------------------
def ray_data_of(ray_bundle) -> F.RayData:
    """Per-ray tensors of a RayBundle-like object."""
    return F.RayData(
        ray_bundle.origins.reshape(-1, 3),
        ray_bundle.directions.reshape(-1, 3),
        ray_bundle.pixel_area.reshape(-1),
        None if ray_bundle.nears is None else ray_bundle.nears.reshape(-1),
        None if ray_bundle.fars is None else ray_bundle.fars.reshape(-1),
    )


def frustums_per_ray(frustums) -> Tuple[F.RayData, F.SampleIntervals]:
    """Un-broadcast view of Frustums-like [N rays, S samples] (or of a flat batch: every sample its own ray)."""
    shape = tuple(frustums.origins.shape[:-1])
    if len(shape) == 2 and (frustums.origins.stride(1) == 0 or shape[1] == 1):
        rays = F.RayData(frustums.origins[:, 0, :], frustums.directions[:, 0, :], frustums.pixel_area[:, 0, 0])
        return rays, F.SampleIntervals(frustums.starts, frustums.ends)
    rays = F.RayData(frustums.origins.reshape(-1, 3), frustums.directions.reshape(-1, 3), frustums.pixel_area.reshape(-1))
    return rays, F.SampleIntervals(frustums.starts.reshape(-1, 1), frustums.ends.reshape(-1, 1))


def per_ray_of(ray_samples) -> Tuple[F.RayData, F.SampleIntervals]:
    """(per-ray data, sample intervals) of a RaySamples-like object; uses what this package's samplers attached."""
    rd = getattr(ray_samples, "ray_data", None)
    if rd is not None and ray_samples.frustums.starts.dim() == 3:
        return rd, F.SampleIntervals(ray_samples.frustums.starts, ray_samples.frustums.ends)
    return frustums_per_ray(ray_samples.frustums)


@dataclass
class GaussiansStd:
    """Isotropic gaussians: mean [*batch, dim], std [*batch, 1] (utils/math.py:114-145)."""

    mean: Tensor
    std: Tensor

    def nelement(self) -> int:
        return self.mean.nelement()

    def __getitem__(self, index) -> "GaussiansStd":
        return GaussiansStd(mean=self.mean[index], std=self.std[index])

    def __len__(self) -> int:
        return len(self.mean)

    @property
    def dtype(self) -> torch.dtype:
        return self.mean.dtype


class TensorDataclass:
    """Minimal broadcast-view container (semantics of utils/tensor_dataclass.py:68-147)."""

    _shape: tuple

    def __post_init__(self) -> None:
        shapes = []
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if isinstance(v, Tensor):
                shapes.append(v.shape[:-1])
            elif isinstance(v, TensorDataclass):
                shapes.append(v.shape)
            elif isinstance(v, dict):
                shapes.extend(t.shape[:-1] for t in v.values() if isinstance(t, Tensor))
        if not shapes:
            raise ValueError("TensorDataclass must have at least one tensor")
        batch = torch.broadcast_shapes(*shapes)
        for f in dataclasses.fields(self):
            object.__setattr__(self, f.name, self._bcast(getattr(self, f.name), batch))
        object.__setattr__(self, "_shape", tuple(batch))

    @staticmethod
    def _bcast(v, batch):
        if isinstance(v, Tensor):
            return v.broadcast_to((*batch, v.shape[-1]))
        if isinstance(v, TensorDataclass):
            return v.broadcast_to(batch)
        if isinstance(v, dict):
            return {k: TensorDataclass._bcast(t, batch) for k, t in v.items()}
        return v

    def _apply(self, fn_tensor: Callable, fn_dc: Callable):
        def go(v):
            if isinstance(v, Tensor):
                return fn_tensor(v)
            if isinstance(v, TensorDataclass):
                return fn_dc(v)
            if isinstance(v, dict):
                return {k: go(t) for k, t in v.items()}
            return v

        return dataclasses.replace(self, **{f.name: go(getattr(self, f.name)) for f in dataclasses.fields(self)})

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    @property
    def size(self) -> int:
        n = 1
        for s in self._shape:
            n *= s
        return n

    @property
    def ndim(self) -> int:
        return len(self._shape)

    def __len__(self) -> int:
        if not self._shape:
            raise TypeError("len() of a 0-d tensor")
        return self._shape[0]

    def __bool__(self) -> bool:
        return True

    def __getitem__(self, indices):
        if isinstance(indices, (Tensor, int, slice, type(Ellipsis))):
            indices = (indices,)
        elif isinstance(indices, list):
            indices = (torch.as_tensor(indices),)
        full = tuple(indices) + (slice(None),)
        return self._apply(lambda t: t[full], lambda d: d[tuple(indices)])

    def reshape(self, shape):
        if isinstance(shape, int):
            shape = (shape,)
        return self._apply(lambda t: t.reshape((*shape, t.shape[-1])), lambda d: d.reshape(shape))

    def flatten(self):
        return self.reshape((-1,))

    def broadcast_to(self, shape):
        return self._apply(lambda t: t.broadcast_to((*shape, t.shape[-1])), lambda d: d.broadcast_to(shape))

    def to(self, device):
        return self._apply(lambda t: t.to(device), lambda d: d.to(device))


@dataclass
class Frustums(TensorDataclass):
    """Regions of space along rays (cameras/rays.py:33-139)."""

    origins: Tensor
    directions: Tensor
    starts: Tensor
    ends: Tensor
    pixel_area: Tensor
    offsets: Optional[Tensor] = None

    def get_positions(self) -> Tensor:
        """Centre of each frustum (cameras/rays.py:49-58)."""
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def get_start_positions(self) -> Tensor:
        return self.origins + self.directions * self.starts

    def _per_ray(self) -> Tuple[F.RayData, F.SampleIntervals]:
        """Un-broadcast view of these frustums for the kernels: [N rays, S samples]."""
        return frustums_per_ray(self)

    def get_fast_isotropic_gaussian(self, num_multisamples: int = 1, contraction_scale: Optional[float] = None) -> GaussiansStd:
        """Gaussian approximation of each frustum (cameras/rays.py:109-124), one multisample.

        With `contraction_scale` the ScaledSceneContraction(order=inf) of spatial_distortions.py:103-136 is applied
        in the same kernel and the result lives in [0,1]^3; without it the world-space gaussian is returned."""
        if self.offsets is not None:
            raise NotImplementedError()
        if num_multisamples != 1:
            raise NotImplementedError("the B200 path implements the single-multisample configuration NeuRadar uses")
        if contraction_scale is None:
            t = self.starts + (self.ends - self.starts) / 2
            mean = self.origins.unsqueeze(-2) + self.directions.unsqueeze(-2) * t.unsqueeze(-1)
            area = self.pixel_area.unsqueeze(-2) * t.unsqueeze(-1).pow(2)
            std = (area * ((self.ends - self.starts) / 2).unsqueeze(-2)).pow(1 / 3)
            return GaussiansStd(mean=mean, std=std)
        rays, iv = self._per_ray()
        x, std = F.frustum_gaussians(rays, iv, contraction_scale)
        return GaussiansStd(mean=x.view(*self.shape, 1, 3), std=std.view(*self.shape, 1, 1))

    @classmethod
    def get_mock_frustum(cls, device="cpu") -> "Frustums":
        one = torch.ones((1, 1), device=device)
        return Frustums(origins=torch.ones((1, 3), device=device), directions=torch.ones((1, 3), device=device),
                        starts=one, ends=one.clone(), pixel_area=one.clone())


@dataclass
class RaySamples(TensorDataclass):
    """Samples along rays (cameras/rays.py:142-248)."""

    frustums: Frustums
    camera_indices: Optional[Tensor] = None
    deltas: Optional[Tensor] = None
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None

    # attached by this package's samplers (not dataclass fields: they are per-ray, not per-sample)
    ray_data = None  # F.RayData
    euclidean_bins = None  # [N, S+1]
    spacing_bins = None  # [N, S+1]
    spacing = None  # (lambda, scaling) of the power transform

    def intervals(self) -> F.SampleIntervals:
        return F.SampleIntervals(self.frustums.starts, self.frustums.ends)

    def per_ray(self) -> Tuple[F.RayData, F.SampleIntervals]:
        return per_ray_of(self)

    def get_weights(self, densities: Tensor) -> Tensor:
        """Weights from densities [*, S, 1] (cameras/rays.py:188-210), one warp-scan kernel."""
        shape = densities.shape
        S = shape[-2]
        dens = densities.reshape(-1, S)
        if self.deltas is not None and self.euclidean_bins is None:
            # generic RaySamples: deltas are authoritative (they may have been edited in place)
            starts = torch.zeros_like(dens)
            iv = F.SampleIntervals(starts, self.deltas.reshape(-1, S).float() + starts)
        else:
            iv = F.SampleIntervals(self.frustums.starts.reshape(-1, S), self.frustums.ends.reshape(-1, S)) \
                if len(self.shape) != 2 else self.intervals()
        return F.density_weights(dens, iv).view(shape)

    @staticmethod
    def get_weights_and_transmittance_from_alphas(alphas: Tensor, weights_only: bool = False):
        """cameras/rays.py:226-248: T = cumprod([1, 1 - alpha + 1e-7]), w = alpha * T[:-1]; alphas [N,S,1]."""
        N, S = alphas.shape[0], alphas.shape[1]
        a = alphas.reshape(N, S)
        zeros = torch.zeros((N, S + 1), device=a.device, dtype=torch.float32)
        w, _, _, _ = F.alpha_composite(a, None, F.SampleIntervals.from_bins(zeros), trans_eps=1e-7, sky_sample=False)
        w = w.view(N, S, 1)
        if weights_only:
            return w
        # transmittance [N, S+1, 1]: T_0 = 1, T_{i+1} = T_i (1 - alpha_i + 1e-7) = T_i - w_i + 1e-7 T_i
        om = 1.0 - a + 1e-7
        trans = torch.cumprod(torch.cat([torch.ones((N, 1), device=a.device), om], dim=1), dim=1)
        return w, trans.view(N, S + 1, 1)


@dataclass
class RayBundle(TensorDataclass):
    """A bundle of rays (cameras/rays.py:251-357)."""

    origins: Tensor
    directions: Tensor
    pixel_area: Tensor
    camera_indices: Optional[Tensor] = None
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None
    termination_distances: Optional[Tensor] = None

    def __len__(self) -> int:
        return torch.numel(self.origins) // self.origins.shape[-1]

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()[start_idx:end_idx]

    def ray_data(self) -> F.RayData:
        return ray_data_of(self)

    def get_ray_samples(
        self,
        bin_starts: Tensor,
        bin_ends: Tensor,
        spacing_starts: Optional[Tensor] = None,
        spacing_ends: Optional[Tensor] = None,
        spacing_to_euclidean_fn: Optional[Callable] = None,
    ) -> RaySamples:
        """cameras/rays.py:313-357: per-ray fields become [..., 1, k] views broadcast over the samples."""
        deltas = bin_ends - bin_starts
        camera_indices = self.camera_indices[..., None] if self.camera_indices is not None else None
        shaped = self[..., None]
        frustums = Frustums(origins=shaped.origins, directions=shaped.directions, starts=bin_starts, ends=bin_ends,
                            pixel_area=shaped.pixel_area)
        return RaySamples(
            frustums=frustums,
            camera_indices=camera_indices,
            deltas=deltas,
            spacing_starts=spacing_starts,
            spacing_ends=spacing_ends,
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
            metadata=shaped.metadata,
            times=None if self.times is None else self.times[..., None],
        )
