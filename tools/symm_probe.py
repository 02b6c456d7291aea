"""Does torch's symmetric memory (peer-mapped buffers + signal pads) work on this box, and how fast are its all-reduces
next to NCCL's for the gradient pieces of config 2?
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/symm_probe.py"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm

torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
group = dist.group.WORLD


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for mib in (24, 64):
    n = mib << 18
    t = sm.empty(n, dtype=torch.float32, device="cuda")
    hdl = sm.rendezvous(t, group)
    if rank == 0 and mib == 24:
        print("multicast:", hdl.has_multicast_support, "world", hdl.world_size, "signal pad", hdl.signal_pad_size, flush=True)
    ref = torch.ones((n,), device="cuda")
    t.fill_(float(rank + 1))
    torch.cuda.synchronize(); dist.barrier()
    torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", group.group_name)
    torch.cuda.synchronize()
    ok = bool((t == world * (world + 1) / 2).all())
    res = {"nccl": timed(lambda: dist.all_reduce(ref)),
           "two_shot": timed(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", group.group_name))}
    if hdl.has_multicast_support:
        try:
            res["multimem"] = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", group.group_name))
        except Exception as e:  # noqa: BLE001
            res["multimem"] = str(e)[:80]
    if rank == 0:
        print(f"{mib} MiB correct={ok} " + " ".join(f"{k}={v:.0f}us" if isinstance(v, float) else f"{k}={v}" for k, v in res.items()), flush=True)
dist.destroy_process_group()
