"""Scratch: time the field MLP fwd/bwd kernels alone at config-2 size (NRB_FIELD_BWD_DEBUG bits disable stages)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuradar_b200 import functional as Fn
from tests.test_gpu_tensorcore import _field_inputs
DEV = "cuda"
N, S = 65536, 48
x, sh, ws, bs, beta = _field_inputs(256, S, seed=1)
M = N * S
g = torch.Generator(device=DEV).manual_seed(0)
xd = (torch.randn((M, 32), device=DEV, generator=g) * 0.5).requires_grad_(True)
shd = sh.to(DEV).repeat(N // 256, 1).contiguous()
wd = [w.to(DEV).requires_grad_(True) for w in ws]
bd = [b.to(DEV).requires_grad_(True) for b in bs]
betad = beta.to(DEV).requires_grad_(True)
gf = torch.randn((M, 32), device=DEV, generator=g); ga = torch.randn((M,), device=DEV, generator=g)
def run():
    f, s_, a = Fn.field_mlp(xd, shd, S, wd, bd, betad, 1e-4)
    torch.autograd.backward([f, a], [gf, ga])
from neuradar_b200 import _lib
for _ in range(2): run()
_lib.TIMER = _lib.KernelTimer()
for _ in range(5): run()
print(os.environ.get("NRB_FIELD_BWD_DEBUG", "0"), {k: round(v[1], 3) for k, v in _lib.TIMER.summary().items()})
