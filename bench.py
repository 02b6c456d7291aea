#!/usr/bin/env python
"""Benchmark of the NeuRadar per-ray hot path (BASELINE.json: train-step rays/sec at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A step = one fwd+bwd pass of the hot path (proposal sampling -> hash encode -> MLPs -> compositing, loss of
SURVEY.md 8d, gradients of every hash table and MLP) over one batch of synthetic rays; for N > 1 each rank owns
its own rays (weak scaling, 65536 rays per GPU) and the step ends with ONE all-reduce of the flat gradient arena.
Workload at N=1 = BASELINE.json configs[1]: 65536 mixed camera/lidar/radar rays, 16-level 2^19 main grid, 48
samples per ray, proposal rounds of 64 and 48 samples on the 6-level 2^20 grid.  Prints ONE JSON line.

  --optimizer      adds the fused Adam / AdamW update of SURVEY.md 8f next-2 to the step (both arms)
  --regularisers   adds the interlevel + distortion losses of SURVEY.md 8f next-1 to the step (both arms)

`roofline` = the kernel with the largest share of the step (`rooflines` has every major kernel): algorithmic bytes or
FLOPs per launch / CUDA-event duration, against MEASURED_PEAKS.json; `traffic` = DRAM bytes per
launch from the committed ncu capture (profiles/r1_traffic.json).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train_step_rays_per_sec"
UNIT = "rays/s"
RAYS_PER_GPU = 65536
PROP_SAMPLES = (64, 48)
NERF_SAMPLES = 48
WORKLOAD = ("config2: 65536 mixed camera/lidar/radar rays per GPU, main grid L16/F2/T2^19 (res 16..1024), "
            "proposals (64,48) on L6/F1/T2^20, 48 samples/ray, fwd+bwd")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(num_rays: int, steps: int, warmup: int, seed: int = 42, regularisers: bool = False,
                      optimizer: bool = False):
    """fwd+bwd of the reference algorithm (CPU oracle port, fp32 torch ops) on `num_rays` rays of the workload."""
    from oracle import neuradar_oracle as O
    from tests.parity_utils import scaled_pixel_area, synthetic_rays

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    main = O.GridParams(((torch.rand((16 << 19, 2)) * 2 - 1) * 1e-3).requires_grad_(True), O.level_scalings(16, 16, 1024), 19)
    lin = torch.nn.Linear

    def wb(i, o):
        layer = lin(i, o)
        return layer.weight.detach().requires_grad_(True), layer.bias.detach().requires_grad_(True)

    g0, g1 = wb(32, 32), wb(32, 33)
    f0, f1, f2 = wb(48, 32), wb(32, 32), wb(32, 32)
    fld = O.FieldParams(main, [g0[0], g1[0]], [g0[1], g1[1]], [f0[0], f1[0], f2[0]], [f0[1], f1[1], f2[1]],
                        torch.full((1,), 20.0, requires_grad=True))
    prop = O.ProposalParams(
        O.GridParams(((torch.rand((6 << 20, 1)) * 2 - 1) * 1e-3).requires_grad_(True), O.level_scalings(6, 128, 4096), 20),
        lin(6, 1, bias=False).weight.detach().requires_grad_(True),
    )
    rays = synthetic_rays(num_rays, seed=seed)
    pa = scaled_pixel_area(rays)
    cfg = O.PathConfig(num_proposal_samples=PROP_SAMPLES, num_nerf_samples=NERF_SAMPLES)
    leaves = [main.table, *fld.geo_w, *fld.geo_b, *fld.feat_w, *fld.feat_b, fld.beta, prop.grid.table, prop.decoder_w]
    opts = []
    if optimizer:  # the reference's parameter groups (configs/method_configs.py:393-400)
        opts = [torch.optim.Adam([main.table, prop.grid.table], lr=1e-2, eps=1e-15),
                torch.optim.AdamW([t for t in leaves if t is not main.table and t is not prop.grid.table], lr=1e-2, eps=1e-15,
                                  weight_decay=1e-7)]

    def step():
        for t in leaves:
            t.grad = None
        jit = [torch.rand((num_rays, PROP_SAMPLES[0] + 1)), torch.rand((num_rays, 1)), torch.rand((num_rays, 1))]
        out = O.nff_forward(fld, [prop, prop], rays["origins"], rays["directions"], pa, rays["nears"], rays["fars"], cfg, jit)
        loss = O.bench_loss(out)
        if regularisers:
            loss = loss + O.training_losses(out)
        loss.backward()
        for o in opts:
            o.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return num_rays / dt, dt * 1e3, cores


def run_reference(args, rank):
    if rank != 0:
        return
    sample = 4096
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    value, ms, cores = cpu_reference_run(sample, steps, warmup, regularisers=args.regularisers, optimizer=args.optimizer)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference algorithm on the host CPU; the reference is pure Python and "
                   "cannot travel to the GPU box, so this is the oracle port of its torch path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of the 65536 rays per step, {steps} timed fwd+bwd steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
ALGO = {
    # algorithmic bytes per sample of one launch (8 corners x L x F x 4 B, gathers/scatters counted once)
    "nrb_hash_fwd:L16F2T19": ("hbm", 8 * 16 * 2 * 4), "nrb_hash_bwd:L16F2T19": ("hbm", 8 * 16 * 2 * 4),
    "nrb_proposal_fwd": ("hbm", 8 * 6 * 1 * 4), "nrb_proposal_bwd": ("hbm", 8 * 6 * 1 * 4),
}
MLP_FLOP_FWD = {"geo": 2 * (32 * 32 + 32 * 33), "feature": 2 * (48 * 32 + 32 * 32 + 32 * 32)}


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist

    import neuradar_b200 as nb
    from neuradar_b200 import _lib
    from neuradar_b200.dist import GradArena
    from tests.parity_utils import build_hot_path, synthetic_rays

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    n = args.rays
    model = build_hot_path(num_proposal_samples=PROP_SAMPLES, num_nerf_samples=NERF_SAMPLES, seed=42, device=dev)
    model.train()
    used = [(name, p) for name, p in model.named_parameters() if not name.startswith("proposal_fields.0")]
    opt = None
    if args.optimizer:
        # the reference's "hashgrids" (Adam) and "fields" (AdamW) groups, each one flat buffer and one fused kernel
        from neuradar_b200.optim import FusedAdam, FusedAdamW

        tables = [p for name, p in used if name.endswith("hash_table")]
        others = [p for name, p in used if not name.endswith("hash_table")]
        opts = [FusedAdam(tables, lr=1e-2, eps=1e-15, direct_scatter=True),
                FusedAdamW(others, lr=1e-2, eps=1e-15, weight_decay=1e-7)]
        arena_bytes = sum(g.numel() * 4 for o in opts for g in o.flat_grads())
    else:
        arena = GradArena([p for _, p in used], direct_scatter=True)
        arena_bytes = arena.nbytes
    rays = synthetic_rays(n, seed=42 + rank)  # each rank draws its own rays (train.py:104 seeds seed+rank)
    keys = ["origins", "directions", "pixel_area", "nears", "fars", "times", "is_lidar", "is_radar"]
    host = {k: rays[k].pin_memory() for k in keys}
    resident = {k: rays[k].to(dev) for k in keys}
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in keys)

    def bundle(src):
        return nb.RayBundle(origins=src["origins"].clone(), directions=src["directions"].clone(),
                            pixel_area=src["pixel_area"].clone(), nears=src["nears"].clone(), fars=src["fars"].clone(),
                            times=src["times"], metadata={"is_lidar": src["is_lidar"], "is_radar": src["is_radar"]})

    def step(src):
        if not args.optimizer:
            arena.zero()
        out = model(bundle(src))
        loss = nb.bench_loss(out)
        if args.regularisers:
            loss = loss + nb.training_losses(out)
        loss.backward()
        if args.optimizer:
            for o in opts:  # sum over ranks; the average and zero_grad ride inside the fused update
                o.step(grad_mult=o.all_reduce_grads(), zero_grad=True)
        else:
            arena.all_reduce()
        return loss

    def step_e2e():
        dev_rays = {k: host[k].to(dev, non_blocking=True) for k in keys}
        return step(dev_rays).item()  # device -> host read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    for _ in range(args.warmup):
        step(resident)
    launches0 = _lib.launch_count()
    with ClockSampler(local_rank) as clocks:
        ms = timed(lambda: step(resident), args.steps)
    launches = _lib.launch_count() - launches0
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel durations from CUDA events on the launching stream (separate pass: the events add launch gaps)
    _lib.TIMER = _lib.KernelTimer()
    for _ in range(max(3, min(args.steps, 10))):
        step(resident)
    kernels = _lib.TIMER.summary()
    _lib.TIMER = None

    peaks = load_peaks()
    total_rays = n * world
    reps = max(3, min(args.steps, 10))
    per_step = {name: {"launches_per_step": count / reps, "mean_ms": mean_ms} for name, (count, mean_ms) in kernels.items()}
    # Roofline of every major kernel: algorithmic work per launch / CUDA-event duration.
    #   hash / proposal kernels: 8 corners x L x F x 4 B per sample, gathers (or scatters) counted once  -> HBM GB/s
    #   compositor: the [N,S,32] feature tensor streamed once (+ once written in the backward)          -> HBM GB/s
    #   field MLP: 2 * sum(in*out) FLOP per sample forward, twice that backward                          -> tensor TFLOP/s
    n_main = n * NERF_SAMPLES
    mlp_flop = 2 * (32 * 32 + 32 * 33 + 48 * 32 + 32 * 32 + 32 * 32)
    work = {
        "nrb_hash_fwd:L16F2T19": ("hbm", n_main * 1024.0),
        "nrb_hash_bwd:L16F2T19": ("hbm", n_main * 1024.0),
        "nrb_proposal_fwd": ("hbm", n * (PROP_SAMPLES[0] + PROP_SAMPLES[1]) / 2 * 192.0),
        "nrb_proposal_bwd": ("hbm", n * (PROP_SAMPLES[0] + PROP_SAMPLES[1]) / 2 * 192.0),
        "nrb_alpha_composite_fwd": ("hbm", n_main * 32 * 4.0),
        "nrb_alpha_composite_bwd": ("hbm", n_main * 32 * 4.0 * 2),
        "nrb_field_mlp_fwd": ("tensor", n_main * float(mlp_flop)),
        "nrb_field_mlp_bwd": ("tensor", n_main * float(mlp_flop) * 2),
    }
    if args.optimizer:  # read p, g, m, v + write p, m, v, g = 32 B per parameter, averaged over the two groups' launches
        work["nrb_adam_step"] = ("hbm", arena_bytes / 4 * 32.0 / 2)
    # DRAM bytes per launch from the committed ncu --set full capture (None when a kernel was not captured)
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")) as fh:
            captured = json.load(fh)["kernels"] if n == RAYS_PER_GPU else {}
    except (OSError, ValueError, KeyError):
        captured = {}
    rooflines = {}
    for name, (bound, amount) in work.items():
        if name not in kernels:
            continue
        mean_ms = kernels[name][1]
        if bound == "hbm":
            achieved, peak, unit = amount / (mean_ms * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s"
        else:
            achieved, peak, unit = amount / (mean_ms * 1e-3) / 1e12, peaks["tflops"], "TFLOP/s"
        rooflines[name] = {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                           "frac": achieved / peak, "traffic": captured.get(name, {}).get("dram_bytes_per_launch"),
                           "traffic_unit": "B (ncu dram read+write per launch, profiles/r1_traffic.json)",
                           "peak_source": peaks["source"],
                           "algorithmic_per_launch": amount, "mean_launch_ms": mean_ms,
                           "ms_per_step": mean_ms * per_step[name]["launches_per_step"]}
    # the headline entry: the kernel with the largest share of the step
    roofline = max(rooflines.values(), key=lambda r: r["ms_per_step"]) if rooflines else None

    line = {
        "metric": METRIC, "value": total_rays / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": n, "global_rays": total_rays,
                   "parallelism": f"ray-sharded dp{world}, one all-reduce of a {arena_bytes / 2**20:.0f} MiB gradient arena",
                   "l2": "per-step working set (saved activations + gradient arena, > 1 GB) exceeds the 126 MB L2; "
                         "no explicit flush", "optimizer": "fused Adam (hash grids) + AdamW (MLPs) inside the step" if args.optimizer
                   else "not part of the path (SURVEY.md 8f next-2; --optimizer adds it)",
                   "regularisers": bool(args.regularisers)},
        "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline,
        "rooflines": rooflines,
        "kernels": per_step,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            v, cms, cores = cpu_reference_run(4096, 2, 1, regularisers=args.regularisers, optimizer=args.optimizer)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "4096 of the 65536 rays per step, 2 timed fwd+bwd steps of the oracle"}
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", action="store_true",
                    help="add the fused Adam / AdamW update (SURVEY.md 8f next-2) to the step of both arms")
    ap.add_argument("--regularisers", action="store_true",
                    help="add the interlevel + distortion losses (SURVEY.md 8f next-1) to the step of both arms")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: neuradar_b200 has no CPU path (use --impl reference for the CPU oracle)")
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; stdout carries ONE JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
