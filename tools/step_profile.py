"""Every GPU kernel of one eager step of a BASELINE configuration, by total time (torch.profiler / CUPTI): shows what the
step spends outside the nrb_* kernels (torch elementwise glue, memsets, copies) and, under torchrun, when the NCCL
all-reduce kernels run relative to the backward kernels (--timeline).
Usage: python tools/step_profile.py --config 4
       python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_profile.py --timeline"""
import argparse
import os
import sys

import torch

sys.path.insert(0, ".")
import neuradar_b200 as nb  # noqa: E402
from neuradar_b200 import dist as nbdist  # noqa: E402
from neuradar_b200.synthetic import WORKLOADS, build_workload, scaled_pixel_area, synthetic_rays  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--rays", type=int, default=0)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--timeline", action="store_true", help="print the kernels of the last step in start order (>= 20 us)")
args = ap.parse_args()
w = WORKLOADS[args.config]
n = args.rays or (w.chunk or w.rays)
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist

    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
dev = f"cuda:{torch.cuda.current_device()}"
model = build_workload(w, device=dev)
model.train(w.train)
r = synthetic_rays(n, mix=w.mix, device=dev)
r["pixel_area"] = scaled_pixel_area(r)
used = [p for name, p in model.named_parameters() if not name.startswith("proposal_fields.0")]
arena = nbdist.GradArena(used, direct_scatter=True, early=[model.field.hashgrid.static_grid.hash_table]) if w.train else None


def step():
    rb = nb.RayBundle(origins=r["origins"], directions=r["directions"], pixel_area=r["pixel_area"], nears=r["nears"],
                      fars=r["fars"], times=r["times"], metadata={"is_lidar": r["is_lidar"], "is_radar": r["is_radar"]})
    if not w.train:
        with torch.no_grad():
            return model(rb)["depth"].sum()
    arena.zero()
    loss = nb.bench_loss(model(rb))
    loss.backward()
    arena.all_reduce()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3e3, e.count / 3) for e in prof.key_averages() if e.device_time_total > 0
        and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda t: -t[1])
total = sum(t[1] for t in rows)
print(f"config {args.config}, {n} rays: {total:.3f} ms of GPU kernels per step, {sum(t[2] for t in rows):.0f} launches")
for k, ms, c in rows[: args.top]:
    print(f"{ms:8.3f} ms  x{c:5.1f}  {k[:150]}")
mine = sum(ms for k, ms, c in rows if "nrb" in k)
print(f"nrb kernels: {mine:.3f} ms; everything else: {total - mine:.3f} ms")
if args.timeline and rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
    evs.sort(key=lambda e: e.time_range.start)
    t_end = max(e.time_range.end for e in evs)
    last = [e for e in evs if e.time_range.start > t_end - 1.05 * total * 1e3]  # ~ the last step
    t0 = last[0].time_range.start
    print("start_us  dur_us  kernel (last step, rank 0)")
    for e in last:
        if e.device_time >= 20 or "nccl" in e.name.lower():
            print(f"{e.time_range.start - t0:9.0f} {e.device_time:7.0f}  {e.name[:110]}")
if world > 1:
    dist.destroy_process_group()
