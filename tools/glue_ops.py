"""Which torch operators launch the non-nrb kernels of a step, split into path (model forward/backward) and the synthetic
upstream loss: aten op counts from torch.profiler around each phase.  Usage: python tools/glue_ops.py"""
import collections
import sys

import torch

sys.path.insert(0, ".")
import neuradar_b200 as nb  # noqa: E402
from neuradar_b200 import dist as nbdist  # noqa: E402
from neuradar_b200.synthetic import WORKLOADS, build_workload, scaled_pixel_area, synthetic_rays  # noqa: E402

w = WORKLOADS[2]
model = build_workload(w, device="cuda")
model.train()
r = synthetic_rays(w.rays, device="cuda")
used = [p for name, p in model.named_parameters() if not name.startswith("proposal_fields.0")]
arena = nbdist.GradArena(used, direct_scatter=True, early=[model.field.hashgrid.static_grid.hash_table])


def bundle():
    return nb.RayBundle(origins=r["origins"], directions=r["directions"], pixel_area=r["pixel_area"].clone(), nears=r["nears"].clone(),
                        fars=r["fars"].clone(), times=r["times"], metadata={"is_lidar": r["is_lidar"], "is_radar": r["is_radar"]})


def phase(fn):
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        out = fn()
        torch.cuda.synchronize()
    kernels = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ops = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and len(e.kernels) > 0:
            ops[e.name] += len(e.kernels)
    return out, len(kernels), ops


for _ in range(2):
    arena.zero(); nb.bench_loss(model(bundle())).backward()
arena.zero()
out, n_fwd, ops_fwd = phase(lambda: model(bundle()))
loss, n_loss, ops_loss = phase(lambda: nb.bench_loss(out))
_, n_bwd, ops_bwd = phase(lambda: loss.backward())
print("forward:", n_fwd, dict(ops_fwd))
print("loss   :", n_loss, dict(ops_loss))
print("backward:", n_bwd, dict(ops_bwd))
