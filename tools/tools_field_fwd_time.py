import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuradar_b200 import functional as Fn, _lib
from tests.test_gpu_tensorcore import _field_inputs
DEV = "cuda"
N, S = 65536, 48
x, sh, ws, bs, beta = _field_inputs(256, S, seed=1)
M = N * S
xd = torch.randn((M, 32), device=DEV) * 0.5
shd = sh.to(DEV).repeat(N // 256, 1).contiguous()
wd = [w.to(DEV) for w in ws]; bd = [b.to(DEV) for b in bs]; betad = beta.to(DEV)
for save in (False, True):
    for _ in range(2): Fn.field_mlp_forward(xd, shd, S, wd, bd, betad, 1e-4, save=save)
    _lib.TIMER = _lib.KernelTimer()
    for _ in range(5): Fn.field_mlp_forward(xd, shd, S, wd, bd, betad, 1e-4, save=save)
    ms = _lib.TIMER.summary()["nrb_field_mlp_fwd"][1]
    print(os.environ.get("NRB_FIELD_FWD_DEBUG", "0"), "save" if save else "infer", f"{ms:.3f} ms -> {ms*1e-3/(M/128/296)*1.9e9/5:.0f} cycles per layer per CTA")
