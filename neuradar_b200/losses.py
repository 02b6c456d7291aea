"""Per-ray losses on the hot path's outputs, with the reference's signatures (model_components/losses.py).

`distortion_loss(weights_list, ray_samples_list)` (losses.py:151-156) and
`zipnerf_interlevel_loss(weights_list, ray_samples_list)` (losses.py:671-705) are called every training step by
`NeuRadarModel.get_loss_dict` (models/neurad.py:524-545) on the samplers' `weights_list` / `ray_samples_list`.
Each is one warp-per-ray CUDA kernel (csrc/losses.cu) that also emits d loss_ray / d weights, so the backward pass
is a broadcast multiply.  No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor
from torch.amp import custom_bwd, custom_fwd

from . import _lib
from .functional import f32c, ptr, stream_ptr

PULSE_WIDTHS = (0.03, 0.003)  # losses.py:677


def ray_samples_to_sdist(ray_samples) -> Tensor:
    """Spacing-domain bin edges [N, S+1] of a RaySamples (losses.py:107-112)."""
    starts, ends = ray_samples.spacing_starts, ray_samples.spacing_ends
    return torch.cat([starts[..., 0], ends[..., -1:, 0]], dim=-1)


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("neuradar_b200.losses: CUDA tensors required (there is no CPU path)")


class _Distortion(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sdist, w):
        _require_cuda(sdist, w)
        sdist, w = f32c(sdist), f32c(w)
        N, S = w.shape
        loss = torch.empty((N,), device=w.device, dtype=torch.float32)
        need = ctx.needs_input_grad[1]
        gfac = torch.empty_like(w) if need else None
        _lib.call("nrb_distortion_loss", ptr(sdist), sdist.shape[1], ptr(w), N, S, ptr(loss), ptr(gfac), stream_ptr())
        ctx.save_for_backward(gfac)
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("distortion loss: gradients flow to the weights only (bins come from no_grad samplers)")
        return loss

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dloss):
        (gfac,) = ctx.saved_tensors
        return None, (gfac * dloss[:, None]) if gfac is not None else None


def lossfun_distortion(t: Tensor, w: Tensor) -> Tensor:
    """Per-ray distortion loss (losses.py:137-148): t [N, S+1] bin edges, w [N, S] weights -> [N]."""
    S = w.shape[-1]
    out = _Distortion.apply(t.reshape(-1, S + 1), w.reshape(-1, S))
    return out.view(w.shape[:-1])


def distortion_loss(weights_list: Sequence[Tensor], ray_samples_list: Sequence) -> Tensor:
    """MipNeRF-360 distortion loss of the final level, mean over rays (losses.py:151-156)."""
    c = ray_samples_to_sdist(ray_samples_list[-1])
    w = weights_list[-1][..., 0]
    return torch.mean(lossfun_distortion(c, w))


class _Interlevel(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, c, w, cp, wp, pulse_width):
        _require_cuda(c, w, cp, wp)
        c, w, cp, wp = f32c(c), f32c(w), f32c(cp), f32c(wp)
        N, Sc = w.shape
        Sp = wp.shape[1]
        loss = torch.empty((N,), device=w.device, dtype=torch.float32)
        gfac = torch.empty_like(wp) if ctx.needs_input_grad[3] else None
        _lib.call("nrb_interlevel_loss", ptr(c), c.shape[1], ptr(w), Sc, ptr(cp), cp.shape[1], ptr(wp), Sp,
                  float(pulse_width), N, ptr(loss), ptr(gfac), stream_ptr())
        ctx.save_for_backward(gfac)
        return loss

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dloss):
        (gfac,) = ctx.saved_tensors
        return None, None, None, (gfac * dloss[:, None]) if gfac is not None else None, None


def interlevel_per_ray(c: Tensor, w: Tensor, cp: Tensor, wp: Tensor, pulse_width: float) -> Tensor:
    """One proposal round's anti-aliased interlevel loss per ray [N]; c, w (final level) and cp are constants."""
    return _Interlevel.apply(c.detach(), w.detach(), cp.detach(), wp, pulse_width)


def zipnerf_interlevel_loss(weights_list: Sequence[Tensor], ray_samples_list: Sequence) -> Tensor:
    """Anti-aliased interlevel loss of ZipNeRF, mean over rays and summed over proposal rounds (losses.py:671-705)."""
    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = weights_list[-1][..., 0].detach()
    loss = 0
    for i, (ray_samples, weights) in enumerate(zip(ray_samples_list[:-1], weights_list[:-1])):
        cp = ray_samples_to_sdist(ray_samples)
        wp = weights[..., 0]
        loss = loss + interlevel_per_ray(c, w, cp, wp, PULSE_WIDTHS[i]).mean()
    return loss
