"""Scratch: clock64 phase trace of field_fused_bwd_kernel.  Builds a debug copy of the library with -DNRB_FUSED_TRACE."""
import os, sys, subprocess, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1:] == ["build"]:
    from neuradar_b200 import build as B
    os.makedirs(os.path.join(ROOT, "tools", "bin"), exist_ok=True)
    objs = []
    for src in B.sources():
        obj = os.path.join(ROOT, "tools", "bin", os.path.basename(src)[:-3] + ".trace.o")
        subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, "-DNRB_FUSED_TRACE", "-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o",
                           os.path.join(ROOT, "tools", "bin", "libtrace.so"), *objs])
    sys.exit(0)
import torch
from neuradar_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libtrace.so")
from neuradar_b200 import functional as Fn
from tests.test_gpu_tensorcore import _field_inputs
from neuradar_b200.field_components import HashEncoding
DEV = "cuda"
N, S = 65536, 48
_, sh, ws, bs, beta = _field_inputs(256, S, seed=3)
grid = HashEncoding().to(DEV)
M = N * S
o = torch.rand((N, 1, 3), device=DEV) * 0.2 + 0.4
d = torch.randn((N, 1, 3), device=DEV); d = d / d.norm(dim=-1, keepdim=True)
tt = (torch.arange(S, device=DEV).float() / S * 0.05).view(1, S, 1)
x3 = (o + d * tt).reshape(M, 3).clamp(0, 1).contiguous()
std = torch.full((M,), 1e-3, device=DEV)
shd = sh.to(DEV).repeat(N // 256, 1).contiguous()
gf = torch.randn((M, 32), device=DEV); ga = torch.randn((M,), device=DEV)
wd = [w.to(DEV).requires_grad_(True) for w in ws]; bd = [b.to(DEV).requires_grad_(True) for b in bs]
betad = beta.to(DEV).requires_grad_(True)
for _ in range(2):
    grid.hash_table.grad = None
    f, s_, a = Fn.field_fused(grid.hash_table, None, x3, std, shd, S, grid.spec, wd, bd, betad, 1e-4)
    torch.autograd.backward([f, a], [gf, ga])
lib = _lib.load()
buf = (C.c_longlong * 4096)()
lib.nrb_debug_fused_trace.argtypes = [C.c_void_p, C.c_int32]
print("rc", lib.nrb_debug_fused_trace(buf, 4096))
ev = [(v >> 48, v & 0xFFFFFFFFFFFF) for v in buf[:2000] if v]
iev = [(v >> 48, v & 0xFFFFFFFFFFFF) for v in buf[2000:4000] if v]
# per-tile table for the worker: deltas between consecutive events, for tiles 3..6
tiles = []
cur = []
for slot, c in ev:
    if slot == 0 and cur:
        tiles.append(cur); cur = []
    cur.append((slot, c))
if cur: tiles.append(cur)
print("worker tiles traced:", len(tiles))
for tl in tiles[3:7]:
    t0 = tl[0][1]
    print(" ".join(f"{s}:{c - t0}" for s, c in tl), "| total", tl[-1][1] - t0)
import collections
acc = collections.defaultdict(list)
for tl in tiles[2:]:
    for (s0, c0), (s1, c1) in zip(tl, tl[1:]):
        acc[(s0, s1)].append(c1 - c0)
print("mean cycles between consecutive trace points (from, to): mean")
for k, v in acc.items():
    print(f"  {k}: {sum(v) / len(v):.0f}")
tile_tot = [tl[-1][1] - tl[0][1] for tl in tiles[2:-1]]
starts = [tl[0][1] for tl in tiles]
print("mean tile body", sum(tile_tot) / max(len(tile_tot), 1), "mean tile period", (starts[-1] - starts[2]) / max(len(starts) - 3, 1))
# issuer (layer 4): issue durations
dur = [c1 - c0 for (s0, c0), (s1, c1) in zip(iev, iev[1:]) if s0 in (100, 101) and s1 in (110, 111)]
print("layer-4 dW issue: n", len(dur), "mean cycles", sum(dur) / max(len(dur), 1))
