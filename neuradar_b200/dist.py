"""Data parallelism of the hot path: rays are independent, so each rank processes its own contiguous slice of the
global batch with replicated parameters, and the only exchange is ONE all-reduce over a flat gradient arena
(hash tables + MLP weights) per step.  Replaces the reference's DistributedDataParallel wrap with
find_unused_parameters=True and 25 MiB buckets (nerfstudio/pipelines/base_pipeline.py:305-307): same result
(sum / world_size), one NCCL call that NVSwitch can reduce in-network (NVLS).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn


def shard_bounds(num_rays: int, world_size: int, rank: int, granule: int = 1) -> Tuple[int, int]:
    """Contiguous slice [start, end) of a global ray batch owned by `rank`.

    `granule` keeps camera patches (1024 rays) or radar scans (256 rays) whole: the decoders downstream consume
    them as units.  Granules are dealt as evenly as possible, earlier ranks get the remainder."""
    if num_rays % granule != 0:
        raise ValueError(f"num_rays={num_rays} is not a multiple of granule={granule}")
    units = num_rays // granule
    base, rem = divmod(units, world_size)
    start = rank * base + min(rank, rem)
    end = start + base + (1 if rank < rem else 0)
    return start * granule, end * granule


class GradArena:
    """One contiguous fp32 buffer that holds the gradient of every trainable parameter as a view.

    `param.grad` is pointed into the arena, so autograd accumulates straight into it, `zero()` is one memset and
    `all_reduce()` is one collective.  Parameters that receive no gradient in a step (e.g. the never-evaluated
    first proposal field) simply contribute zeros, which is what DDP's find_unused_parameters does.
    `direct_scatter=True` additionally lets the hash-grid backward kernels add straight into the arena
    (functional.grad_sink_of) instead of returning a table-sized temporary to autograd."""

    def __init__(self, params: Iterable[nn.Parameter], skip_unused: Optional[List[nn.Parameter]] = None,
                 direct_scatter: bool = False):
        skip = {id(p) for p in (skip_unused or [])}
        self.params = [p for p in params if p.requires_grad and id(p) not in skip]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, total = self.params[0].device, 0
        offsets = []
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("GradArena expects fp32 parameters on one device")
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every view 16-byte aligned for the vector atomics
        self.flat = torch.zeros((total,), device=dev, dtype=torch.float32)
        for p, off in zip(self.params, offsets):
            p.grad = self.flat[off : off + p.numel()].view_as(p)
            if direct_scatter and p.dim() == 2 and p.numel() >= (1 << 16):
                p._nrb_grad_sink = p.grad  # hash tables: the scatter kernels add straight into the arena

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, group=None, average: bool = True) -> None:
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size(group))
