"""Drop-in for the three nerfacc entry points the NeuRadar model calls (nerfacc==0.5.2, dense inputs only):

  nerfacc.render_weight_from_alpha    <- nerfstudio/models/neuradar.py:1016, models/neurad.py:711
  nerfacc.render_weight_from_density  <- nerfstudio/models/neuradar.py:1018-1022
  nerfacc.accumulate_along_rays       <- nerfstudio/models/neurad.py:728, model_components/renderers.py:88,345,404

`install()` registers this module as `nerfacc` in sys.modules when the real package is absent, which is what lets
the unmodified reference model import on a box without nerfacc (secondary doors).
"""
from __future__ import annotations

import sys
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import functional as F


def _dense_only(packed_info, ray_indices):
    if packed_info is not None or ray_indices is not None:
        raise NotImplementedError("packed samples are not produced on the NeuRadar path; dense [N,S] inputs only")


def render_weight_from_alpha(alphas: Tensor, packed_info=None, ray_indices=None, n_rays=None,
                             prefix_trans=None) -> Tuple[Tensor, Tensor]:
    """weights = alpha * exclusive_cumprod(1 - alpha) along the last dim; returns (weights, transmittance)."""
    _dense_only(packed_info, ray_indices)
    if prefix_trans is not None:
        raise NotImplementedError("prefix_trans is not used on the NeuRadar path")
    shape = alphas.shape
    w, t = F.alpha_weights(alphas.reshape(-1, shape[-1]), trans_eps=0.0)
    return w.view(shape), t.view(shape)


def render_weight_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info=None, ray_indices=None,
                               n_rays=None, prefix_trans=None) -> Tuple[Tensor, Tensor, Tensor]:
    """weights = (1 - exp(-sigma dt)) * exp(-exclusive_cumsum(sigma dt)); returns (weights, transmittance, alphas)."""
    _dense_only(packed_info, ray_indices)
    shape = sigmas.shape
    S = shape[-1]
    iv = F.SampleIntervals(t_starts.reshape(-1, S), t_ends.reshape(-1, S))
    w = F.density_weights(sigmas.reshape(-1, S), iv).view(shape)
    alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
    trans = torch.where(alphas > 0, w / alphas.clamp_min(1e-30), torch.ones_like(w))
    return w, trans, alphas


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor] = None, ray_indices=None,
                          n_rays: Optional[int] = None) -> Tensor:
    """sum over the sample dim of weights[..., S] * values[..., S, C] (or of the weights alone): [..., C] / [..., 1]."""
    if ray_indices is not None:
        raise NotImplementedError("packed samples are not produced on the NeuRadar path; dense [N,S] inputs only")
    lead, S = weights.shape[:-1], weights.shape[-1]
    w = weights.reshape(-1, S)
    v = None if values is None else values.reshape(-1, S, values.shape[-1])
    out = F.accumulate(w, v)
    return out.view(*lead, out.shape[-1])


class OccGridEstimator:  # imported by name at module scope in the reference (ray_samplers.py:25)
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("occupancy-grid sampling is not on the NeuRadar path")


def install(force: bool = False) -> bool:
    """Expose this module as `nerfacc` if the real one is not importable."""
    if "nerfacc" in sys.modules and not force:
        return False
    if not force:
        try:
            import nerfacc  # noqa: F401

            return False
        except ImportError:
            pass
    sys.modules["nerfacc"] = sys.modules[__name__]
    return True
