"""Renderers of the hot path (dense, un-packed samples) with the reference's signatures.

FeatureRenderer       <- nerfstudio/model_components/renderers.py:59-90
AccumulationRenderer  <- nerfstudio/model_components/renderers.py:322-350
DepthRenderer         <- nerfstudio/model_components/renderers.py:353-418
render_depth_simple   <- nerfstudio/models/neurad.py:721-728
The weights are given, so the weighted sums go through `nerfacc_compat.accumulate_along_rays`, i.e. `nrb_accumulate_fwd/bwd`
(`accumulate_fwd_kernel`, csrc/compositing.cu: one warp per ray).  On the model's own path the field, the weights and these
sums are one fused node instead (`NeuRADField.render`, `functional.field_render`).
"""
from __future__ import annotations

from typing import Literal, Optional

import torch
from torch import Tensor, nn

from . import nerfacc_compat
from .rays import RaySamples


class FeatureRenderer(nn.Module):
    def forward(self, features: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None) -> Tensor:
        if ray_indices is not None and num_rays is not None:
            raise NotImplementedError("packed samples are not produced on the NeuRadar path")
        return nerfacc_compat.accumulate_along_rays(weights[..., 0], values=features)


class AccumulationRenderer(nn.Module):
    @classmethod
    def forward(cls, weights: Tensor, ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None) -> Tensor:
        if ray_indices is not None and num_rays is not None:
            raise NotImplementedError("packed samples are not produced on the NeuRadar path")
        return nerfacc_compat.accumulate_along_rays(weights[..., 0], values=None)


def render_depth_simple(weights: Tensor, ray_samples: RaySamples, ray_indices=None, num_rays=None) -> Tensor:
    steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
    return nerfacc_compat.accumulate_along_rays(weights[..., 0], values=steps)


class DepthRenderer(nn.Module):
    def __init__(self, method: Literal["median", "expected"] = "median") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights: Tensor, ray_samples: RaySamples, ray_indices=None, num_rays=None) -> Tensor:
        if ray_indices is not None and num_rays is not None:
            raise NotImplementedError("packed samples are not produced on the NeuRadar path")
        steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
        if self.method == "median":
            cumulative = torch.cumsum(weights[..., 0], dim=-1)
            split = torch.ones((*weights.shape[:-2], 1), device=weights.device) * 0.5
            idx = torch.clamp(torch.searchsorted(cumulative, split, side="left"), 0, steps.shape[-2] - 1)
            return torch.gather(steps[..., 0], dim=-1, index=idx)
        if self.method == "expected":
            depth = nerfacc_compat.accumulate_along_rays(weights[..., 0], values=steps)
            acc = nerfacc_compat.accumulate_along_rays(weights[..., 0], values=None)
            depth = depth / (acc + 1e-10)
            return torch.clip(depth, steps.min(), steps.max())
        raise NotImplementedError(f"Method {self.method} not implemented")
