"""Scratch: per-phase clock64 trace of the column-split field backward kernel (CTA 0, thread 0)."""
import ctypes, os, sys
os.environ["NRB_FIELD_BWD_DEBUG"] = str(4 | int(os.environ.get("EXTRA_DBG", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuradar_b200 import functional as Fn, _lib
from tests.test_gpu_tensorcore import _field_inputs
DEV = "cuda"; N, S = 65536, 48
x, sh, ws, bs, beta = _field_inputs(256, S, seed=1)
M = N * S
g = torch.Generator(device=DEV).manual_seed(0)
xd = (torch.randn((M, 32), device=DEV, generator=g) * 0.5).requires_grad_(True)
shd = sh.to(DEV).repeat(N // 256, 1).contiguous()
wd = [w.to(DEV).requires_grad_(True) for w in ws]; bd = [b.to(DEV).requires_grad_(True) for b in bs]
betad = beta.to(DEV).requires_grad_(True)
gf = torch.randn((M, 32), device=DEV, generator=g); ga = torch.randn((M,), device=DEV, generator=g)
for _ in range(3):
    f, s_, a = Fn.field_mlp(xd, shd, S, wd, bd, betad, 1e-4)
    torch.autograd.backward([f, a], [gf, ga])
torch.cuda.synchronize()
lib = _lib.load() if hasattr(_lib, "load") else _lib.LIB
buf = (ctypes.c_longlong * 2048)()
lib.nrb_debug_bwd_trace.argtypes = [ctypes.c_void_p, ctypes.c_int32]
print("rc", lib.nrb_debug_bwd_trace(buf, 2048))
raw = np.array(buf[:], dtype=np.int64)
tr = raw[:1024]
itr = raw[1024:]
per = 18
tiles = len(tr) // per
tr = tr[: tiles * per].reshape(tiles, per)
d = np.diff(tr, axis=1)
sel = d[5:60]
print("tile cycles (start->dx stored) median", np.median(tr[5:60, -1] - tr[5:60, 0]), " tile-to-tile", np.median(np.diff(tr[5:60, 0])))
names = ["0>1 bounce df + issue loads", "1>2 wait_dw l4", "2>3 stage l4", "3>4 signal+wait dIn l4", "4>5 ld+mask+wait_dw l3", "5>6 stage l3",
         "6>7 signal+wait dIn l3", "7>8 ld+mask+wait_dw l2", "8>9 stage l2 (+SH)", "9>10 signal+wait dIn l2", "10>11 ld+sdf+wait_dw l1",
         "11>12 stage l1", "12>13 signal+bounce x+wait dIn l1", "13>14 ld+mask+wait_dw l0", "14>15 stage l0", "15>16 signal+prefetch+wait dIn l0",
         "16>17 ld+store dx"]
for n, v in zip(names, np.median(sel, axis=0)):
    print(f"{n:34s} {v:8.0f}")
# issuer: 3 stamps per layer, 15 per tile
n_it = (len(itr) // 15) * 15
I = itr[:n_it].reshape(-1, 5, 3)[5:50]
W = tr[5:50]
print("issuer per layer (l4..l0): arrive->dIn issued, dIn issued->dW issued")
print(np.median(I[:, :, 1] - I[:, :, 0], axis=0), np.median(I[:, :, 2] - I[:, :, 1], axis=0))
# worker signal stamps are 3, 6, 9, 12, 15; dIn-done stamps 4, 7, 10, 13, 16
sig = W[:, [3, 6, 9, 12, 15]]; done = W[:, [4, 7, 10, 13, 16]]
print("thread0 signal -> all arrived:", np.median(I[:, :, 0] - sig, axis=0))
print("dIn issued -> worker sees dIn done:", np.median(done - I[:, :, 1], axis=0))
