"""Scratch: clock64 trace of the two-tiles-in-flight field backward kernel (group 0 thread 0 + issuer)."""
import ctypes, os, sys
os.environ["NRB_FIELD_BWD_DEBUG"] = "4"; os.environ["NRB_FIELD_BWD_SPLIT"] = "22"
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
lib = _lib.load()
buf = (ctypes.c_longlong * 2048)()
lib.nrb_debug_bwd_trace.argtypes = [ctypes.c_void_p, ctypes.c_int32]
lib.nrb_debug_bwd_trace(buf, 2048)
raw = np.array(buf[:], dtype=np.int64)
w, it = raw[:1024], raw[1024:]
# worker: per layer 4 stamps: before wait_dw, after wait_dw, staged, dIn done -> 20 per tile
per = 20
nt = len(w) // per
W = w[: nt * per].reshape(nt, 5, 4)[3:40]
print("tile period (group 0):", np.median(np.diff(W[:, 0, 0])))
print("per layer l4..l0:  wait_dw | stage | signal->dIn done | dIn done->next wait_dw")
nxt = np.concatenate([W[:, 1:, 0], np.full((W.shape[0], 1), np.nan)], axis=1)
for name, v in (("wait_dw", W[:, :, 1] - W[:, :, 0]), ("stage", W[:, :, 2] - W[:, :, 1]), ("sig->dIn", W[:, :, 3] - W[:, :, 2]), ("dIn->next", nxt - W[:, :, 3])):
    print(f"{name:10s}", np.nanmedian(v, axis=0))
# data issuer of group 0: 2 stamps per layer (all arrived, chain issued): 10 per tile
I = it[: (len(it) // 10) * 10].reshape(-1, 5, 2)[3:40]
print("data issuer g0: dIn issue per layer", np.median(I[..., 1] - I[..., 0], axis=0).tolist())
sig = W[:, :, 2]
print("thread0 staged -> all of the group arrived:", np.median(I[: len(sig), :, 0] - sig[: len(I)], axis=0).tolist())
print("chain issued -> thread0 sees dIn done:", np.median(W[: len(I), :, 3] - I[: len(W), :, 1], axis=0).tolist())
