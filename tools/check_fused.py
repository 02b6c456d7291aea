"""Scratch: numerics + timing of the fused field kernels (csrc/field_fused.cu) against the oracle and the round-1 chain."""
import os, sys, time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuradar_b200 import functional as Fn, _lib
from tests.test_gpu_tensorcore import _field_inputs, _field_ref
from tests.parity_utils import rel_err

DEV = "cuda"
torch.manual_seed(0)


def check_x_mode(N, S):
    x, sh, ws, bs, beta = _field_inputs(N, S, seed=7 * N + S)
    M = N * S
    g = torch.Generator().manual_seed(1)
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
    f, s_, a = Fn.field_fused(None, xd, None, None, sh.to(DEV), S, None, wd, bd, betad, 1e-4)
    torch.cuda.synchronize()
    out = {"feat": rel_err(f, rf), "sdf": rel_err(s_, rs), "alpha": float((a.cpu() - ra).abs().max())}
    ((f * gf.to(DEV)).sum() + (s_ * gs.to(DEV)).sum() + (a * ga.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    out["dx"] = rel_err(xd.grad, xr.grad)
    for k in range(5):
        out[f"dW{k}"] = rel_err(wd[k].grad, wr[k].grad)
        out[f"db{k}"] = rel_err(bd[k].grad, br[k].grad)
    out["dbeta"] = rel_err(betad.grad, betar.grad)
    print(f"x-mode N={N} S={S}: " + " ".join(f"{k}={v:.1e}" for k, v in out.items()), flush=True)


def check_gather_mode(N, S, log2=14):
    """fused gather + MLP against hash_encode -> field_mlp (round-1 chain)."""
    from neuradar_b200.field_components import HashEncoding
    _, sh, ws, bs, beta = _field_inputs(N, S, seed=3)
    grid = HashEncoding(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=log2, features_per_level=2).to(DEV)
    with torch.no_grad():
        grid.hash_table.mul_(300.0)
    M = N * S
    x3 = torch.rand((M, 3), device=DEV)
    std = torch.rand((M,), device=DEV) * 0.01
    g = torch.Generator().manual_seed(1)
    gf, gs, ga = [t.to(DEV) for t in (torch.randn((M, 32), generator=g), torch.randn((M,), generator=g), torch.randn((M,), generator=g))]
    res = []
    for fused in (False, True):
        grid.hash_table.grad = None
        wd = [w.to(DEV).requires_grad_(True) for w in ws]
        bd = [b.to(DEV).requires_grad_(True) for b in bs]
        betad = beta.to(DEV).requires_grad_(True)
        if fused:
            f, s_, a = Fn.field_fused(grid.hash_table, None, x3, std, sh.to(DEV), S, grid.spec, wd, bd, betad, 1e-4)
        else:
            feats = Fn.hash_encode(x3, grid.hash_table, grid.spec, std, samples_per_ray=S)
            f, s_, a = Fn.field_mlp(feats, sh.to(DEV), S, wd, bd, betad, 1e-4)
        ((f * gf).sum() + (s_ * gs).sum() + (a * ga).sum()).backward()
        torch.cuda.synchronize()
        res.append((f.detach(), s_.detach(), a.detach(), grid.hash_table.grad.clone(), [w.grad.clone() for w in wd], betad.grad.clone()))
    o, n = res
    print(f"gather-mode N={N} S={S}: feat={rel_err(n[0], o[0]):.1e} sdf={rel_err(n[1], o[1]):.1e} alpha={float((n[2]-o[2]).abs().max()):.1e} "
          f"dtable={rel_err(n[3], o[3]):.1e} dW0={rel_err(n[4][0], o[4][0]):.1e} dW4={rel_err(n[4][4], o[4][4]):.1e} dbeta={rel_err(n[5], o[5]):.1e}", flush=True)


def timing(N=65536, S=48):
    from neuradar_b200.field_components import HashEncoding
    _, sh, ws, bs, beta = _field_inputs(256, S, seed=3)
    grid = HashEncoding().to(DEV)
    M = N * S
    # ray-ordered samples: points along rays so that neighbouring samples share cells like on the real path
    o = torch.rand((N, 1, 3), device=DEV) * 0.2 + 0.4
    d = torch.randn((N, 1, 3), device=DEV)
    d = d / d.norm(dim=-1, keepdim=True)
    tt = (torch.arange(S, device=DEV).float() / S * 0.05).view(1, S, 1)
    x3 = (o + d * tt).reshape(M, 3).clamp(0, 1).contiguous()
    std = torch.full((M,), 1e-3, device=DEV)
    shd = sh.to(DEV).repeat(N // 256, 1).contiguous()
    gf = torch.randn((M, 32), device=DEV)
    ga = torch.randn((M,), device=DEV)
    for fused in (False, True):
        wd = [w.to(DEV).requires_grad_(True) for w in ws]
        bd = [b.to(DEV).requires_grad_(True) for b in bs]
        betad = beta.to(DEV).requires_grad_(True)

        def step():
            grid.hash_table.grad = None
            if fused:
                f, s_, a = Fn.field_fused(grid.hash_table, None, x3, std, shd, S, grid.spec, wd, bd, betad, 1e-4)
            else:
                feats = Fn.hash_encode(x3, grid.hash_table, grid.spec, std, samples_per_ray=S)
                f, s_, a = Fn.field_mlp(feats, shd, S, wd, bd, betad, 1e-4)
            torch.autograd.backward([f, a], [gf, ga])

        for _ in range(3):
            step()
        _lib.TIMER = _lib.KernelTimer()
        for _ in range(5):
            step()
        summ = _lib.TIMER.summary()
        _lib.TIMER = None
        print("fused" if fused else "round-1 chain", {k: round(v[1], 3) for k, v in summ.items()}, "total",
              round(sum(v[1] for v in summ.values()), 3), "ms", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "x"):
        for N, S in [(8, 48), (3, 128), (600, 48), (129, 33)]:
            check_x_mode(N, S)
    if what in ("all", "gather"):
        for N, S in [(8, 48), (700, 48), (129, 33)]:
            check_gather_mode(N, S)
    if what in ("all", "time"):
        timing()
