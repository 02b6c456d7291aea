"""Radar ray generation on the GPU: `Radars._generate_rays_from_fov`
(nerfstudio/cameras/radars.py:268-357) as one kernel instead of a python loop over scans.

`generate_rays_from_fov(radars, scan_indices)` takes the reference's own `Radars` object (or anything with its buffers:
`radar_to_worlds [R,3,4]`, `min/max_azimuth`, `min/max_elevation`, `radar_azimuth/elevation_ray_divergence [R,1]`, `times`,
`metadata`) and returns a RayBundle with the fields the reference fills: origins, directions, pixel_area, camera_indices,
times, fars = 1e6 and metadata {directions_norm, did_return, directions_spher, *indexed sensor metadata}.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import f32c, ptr, stream_ptr
from .rays import RayBundle


def fov_grid_sizes(radars) -> Dict[str, Tensor]:
    """Rays per pose along azimuth and elevation: len(torch.arange(min, max, step)) = ceil((max - min) / step), evaluated
    in double like torch does.  The field of view is a static property of the sensor: computed once on the host."""
    cached = getattr(radars, "_nrb_fov_grid", None)
    if cached is not None:
        return cached

    def count(lo: Tensor, hi: Tensor, step: Tensor) -> Tensor:
        lo, hi, step = (t.detach().reshape(-1).double().cpu() for t in (lo, hi, step))
        return torch.ceil((hi - lo) / step).clamp_min(0).to(torch.int64)

    grid = {"n_az": count(radars.min_azimuth, radars.max_azimuth, radars.radar_azimuth_ray_divergence),
            "n_el": count(radars.min_elevation, radars.max_elevation, radars.radar_elevation_ray_divergence)}
    try:
        object.__setattr__(radars, "_nrb_fov_grid", grid)
    except Exception:  # noqa: BLE001 - a frozen container: recompute next time
        pass
    return grid


def generate_rays_from_fov(radars, scan_indices: Tensor, bundle_cls=RayBundle) -> RayBundle:
    dev = radars.radar_to_worlds.device
    if dev.type != "cuda":
        raise _lib.NeuradarB200Error("neuradar_b200 ops need CUDA tensors; there is no CPU path")
    grid = fov_grid_sizes(radars)
    scans_host = scan_indices.detach().reshape(-1).to("cpu", torch.int64)
    n_az, n_el = grid["n_az"][scans_host], grid["n_el"][scans_host]
    offsets = torch.zeros((scans_host.numel() + 1,), dtype=torch.int64)
    offsets[1:] = torch.cumsum(n_az * n_el, 0)
    total = int(offsets[-1])
    scans = scans_host.to(dev, non_blocking=True)
    offs = offsets.to(dev, non_blocking=True)
    nel = n_el.to(torch.int32).to(dev, non_blocking=True)
    r2w = f32c(radars.radar_to_worlds.reshape(-1, 3, 4))
    mn_az, st_az = f32c(radars.min_azimuth.reshape(-1)), f32c(radars.radar_azimuth_ray_divergence.reshape(-1))
    mn_el, st_el = f32c(radars.min_elevation.reshape(-1)), f32c(radars.radar_elevation_ray_divergence.reshape(-1))
    origins = torch.empty((total, 3), device=dev, dtype=torch.float32)
    directions = torch.empty_like(origins)
    pixel_area = torch.empty((total, 1), device=dev, dtype=torch.float32)
    spher = torch.empty((total, 2), device=dev, dtype=torch.float32)
    norm = torch.empty((total, 1), device=dev, dtype=torch.float32)
    ray_scan = torch.empty((total,), device=dev, dtype=torch.int64)
    _lib.call("nrb_radar_rays", ptr(r2w), ptr(mn_az), ptr(st_az), ptr(mn_el), ptr(st_el), ptr(scans), ptr(offs), ptr(nel),
              scans_host.numel(), total, ptr(origins), ptr(directions), ptr(pixel_area), ptr(spher), ptr(norm), ptr(ray_scan),
              stream_ptr())
    metadata: Dict[str, Tensor] = {}
    if getattr(radars, "metadata", None):
        metadata = {k: v.reshape(-1, *v.shape[radars.radar_to_worlds.dim() - 2:])[ray_scan] for k, v in radars.metadata.items()}
    metadata["directions_norm"] = norm
    metadata["did_return"] = torch.ones((total, 1), dtype=torch.bool, device=dev)
    metadata["directions_spher"] = spher
    times: Optional[Tensor] = None
    if getattr(radars, "times", None) is not None:
        times = radars.times.reshape(-1, 1)[ray_scan]
    return bundle_cls(origins=origins, directions=directions, pixel_area=pixel_area, camera_indices=ray_scan.unsqueeze(-1),
                      times=times, metadata=metadata, fars=torch.full_like(pixel_area, 1_000_000.0))
