import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuradar_b200 import functional as Fn, _lib
DEV = "cuda"
M = 65536 * 48
x = torch.randn((M, 32), device=DEV); w = torch.randn((32, 32), device=DEV); b = torch.randn((32,), device=DEV)
for _ in range(2): Fn.tc_linear(x, w, b)
_lib.TIMER = _lib.KernelTimer()
for _ in range(5): Fn.tc_linear(x, w, b)
ms = _lib.TIMER.summary()["nrb_tc_linear"][1]
tiles_per_cta = M / 128 / 296
print(os.environ.get("NRB_TC_LINEAR_DEBUG", "0"), f"{ms:.3f} ms  -> {ms*1e-3/tiles_per_cta*1.9e9:.0f} cycles per tile-layer per CTA (2 CTAs/SM)")
