import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuradar_b200 import functional as Fn
from tests.test_gpu_tensorcore import _field_inputs, _field_ref
DEV = "cuda"
N, S = 600, 48
x, sh, ws, bs, beta = _field_inputs(N, S, seed=7 * N + S)
g = torch.Generator().manual_seed(1)
M = N * S
gf, gs, ga = torch.randn((M, 32), generator=g), torch.randn((M,), generator=g), torch.randn((M,), generator=g)
xr = x.clone().requires_grad_(True)
wr = [w.clone().requires_grad_(True) for w in ws]
br = [b.clone().requires_grad_(True) for b in bs]
betar = beta.clone().requires_grad_(True)
rf, rs, ra = _field_ref(xr, sh, S, wr, br, betar)
((rf * gf).sum() + (rs * gs).sum() + (ra * ga).sum()).backward()
xd = x.to(DEV).requires_grad_(True)
wd = [w.to(DEV).requires_grad_(True) for w in ws]
bd = [b.to(DEV).requires_grad_(True) for b in bs]
betad = beta.to(DEV).requires_grad_(True)
f, s_, a = Fn.field_mlp(xd, sh.to(DEV), S, wd, bd, betad, 1e-4)
((f * gf.to(DEV)).sum() + (s_ * gs.to(DEV)).sum() + (a * ga.to(DEV)).sum()).backward()
err = (xd.grad.cpu() - xr.grad).abs().max(dim=1).values
scale = xr.grad.abs().max()
bad = (err > 1e-4 * scale).nonzero().flatten()
print("bad rows:", bad.numel(), "of", M, "first", bad[:10].tolist(), "last", bad[-5:].tolist())
if bad.numel():
    tiles = torch.unique(bad // 128)
    print("bad tiles:", tiles.tolist()[:40], "count", tiles.numel())
    r = int(bad[0]); print("row", r, "err", float(err[r]), "ref max", float(xr.grad[r].abs().max()), "alpha", float(ra[r]), "sdf", float(rs[r]))
for k in range(5):
    print("dW", k, float((wd[k].grad.cpu() - wr[k].grad).abs().max() / wr[k].grad.abs().max()), "db", float((bd[k].grad.cpu() - br[k].grad).abs().max() / br[k].grad.abs().max()))
print("dbeta", float(betad.grad), float(betar.grad))
