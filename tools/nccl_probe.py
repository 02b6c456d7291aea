"""All-reduce duration of the two gradient pieces of config 2 (64 MiB main table, 24 MiB proposal table + MLPs) under the
current NCCL_* environment, alone and with `max_ctas`-limited communicators.
Usage: python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/nccl_probe.py"""
import os

import torch
import torch.distributed as dist

torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
rank = dist.get_rank()


def timed(t, group=None, reps=20):
    for _ in range(5):
        dist.all_reduce(t, group=group)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(t, group=group)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


groups = {"default": None}
for ctas in (2, 4, 8, 16):
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = ctas
    opts.config.min_ctas = 1
    groups[f"max_ctas={ctas}"] = dist.new_group(backend="nccl", pg_options=opts)
env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
for mib in (24, 64, 88):
    t = torch.ones((mib << 18,), device="cuda")
    for name, g in groups.items():
        us = timed(t, g)
        if rank == 0:
            print(f"{env} {mib} MiB {name}: {us:.0f} us  busbw {2 * (dist.get_world_size() - 1) / dist.get_world_size() * mib * 2**20 / us / 1e3:.0f} GB/s", flush=True)
dist.destroy_process_group()
