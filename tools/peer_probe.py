"""Correctness and speed of nrb_peer_all_reduce (csrc/peer_reduce.cu) next to NCCL for the gradient pieces of config 2.
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from neuradar_b200.dist import PeerMemory  # noqa: E402

torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
n = 88 << 18  # 88 MiB of floats
pm = PeerMemory(n, "cuda")
if rank == 0:
    print("multicast pointer:", hex(pm.multicast_ptr), flush=True)
g = torch.Generator(device="cuda").manual_seed(rank)
src = torch.randn((n,), device="cuda", generator=g)
ref = src.clone()
dist.all_reduce(ref)
ref /= world
n_early = 64 << 18
for trial in range(3):
    pm.flat.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    pm.all_reduce(0, 0, n_early, 1.0 / world, 16)
    pm.all_reduce(1, n_early, n - n_early, 1.0 / world)
    torch.cuda.synchronize()
    err = float((pm.flat - ref).abs().max())
    if rank == 0:
        print(f"trial {trial}: max |peer - nccl| = {err:.3e}", flush=True)


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


x = torch.ones((n,), device="cuda")
for name, off, cnt in (("24 MiB", n_early, n - n_early), ("64 MiB", 0, n_early), ("88 MiB", 0, n)):
    t_nccl = timed(lambda: dist.all_reduce(x[off: off + cnt]))
    res = {ctas: timed(lambda: pm.all_reduce(2, off, cnt, 1.0, ctas)) for ctas in (0, 64, 32, 16, 8)}
    if rank == 0:
        print(f"{name}: nccl {t_nccl:.0f} us | peer kernel " + " ".join(f"ctas={k or 'all'}: {v:.0f} us" for k, v in res.items()), flush=True)
dist.destroy_process_group()
