"""GPU tests of the tcgen05 / TMEM kernels: one linear layer (descriptor and layout conventions), then the fused
NeuRAD field kernel (csrc/field_fused.cu) in row mode - hash features given - forward and backward against the fp32
oracle.  Tolerances are far below the 1e-3 parity bar: the forward evaluates every product as 3xTF32, the backward as
bf16 hi/mid pairs (16 mantissa bits per operand), both with fp32 accumulation."""
import pytest
import torch

from oracle import neuradar_oracle as O
from tests.parity_utils import outside_bar, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,n_out", [(32, 32), (32, 33), (48, 32), (48, 48), (32, 2), (48, 16)])
@pytest.mark.parametrize("M", [128, 1000, 70001])
def test_tc_linear(K, n_out, M):
    from neuradar_b200 import functional as Fn

    g = torch.Generator().manual_seed(K * 100 + n_out)
    x = torch.randn((M, K), generator=g) * torch.exp(torch.randn((M, 1), generator=g) * 2)
    w = torch.randn((n_out, K), generator=g) / K**0.5
    b = torch.randn((n_out,), generator=g)
    for relu in (False, True):
        y = Fn.tc_linear(x.to(DEV), w.to(DEV), b.to(DEV), relu=relu)
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        ref = torch.relu(ref) if relu else ref
        err = (y.cpu().double() - ref).abs()
        scale = (x.double().abs() @ w.double().abs().T) + b.double().abs()  # magnitude of the terms summed
        assert float((err / scale).max()) <= 2e-6, (K, n_out, M, relu)
    y0 = Fn.tc_linear(x.to(DEV), w.to(DEV), None)
    assert rel_err(y0, torch.nn.functional.linear(x, w)) <= 1e-5


def _field_inputs(N, S, seed):
    g = torch.Generator().manual_seed(seed)
    lin = torch.nn.Linear
    torch.manual_seed(seed)
    layers = [lin(32, 32), lin(32, 33), lin(48, 32), lin(32, 32), lin(32, 32)]
    ws = [l.weight.detach().clone() for l in layers]
    bs = [l.bias.detach().clone() for l in layers]
    x = torch.randn((N * S, 32), generator=g) * 0.5
    d = torch.randn((N, 3), generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    sh = O.sh16((d + 1.0) / 2.0)
    beta = torch.tensor([20.0])
    return x, sh, ws, bs, beta


def _field_ref(x, sh, S, ws, bs, beta):
    geo = O.mlp(x, ws[:2], bs[:2])
    sdf, emb = geo[:, :1], geo[:, 1:]
    she = sh[:, None, :].expand(-1, S, -1).reshape(-1, 16)
    feat = emb + O.mlp(torch.cat([emb, she], -1), ws[2:], bs[2:])
    alpha = torch.sigmoid(-sdf * (beta.abs() + 1e-4))
    return feat, sdf[:, 0], alpha[:, 0]


@pytest.mark.parametrize("N,S", [(8, 48), (3, 128), (1000, 48), (257, 33)])
def test_field_fused_forward_rows(N, S):
    from neuradar_b200 import functional as Fn

    x, sh, ws, bs, beta = _field_inputs(N, S, seed=N + S)
    with torch.no_grad():
        feat, sdf, alpha = Fn.field_fused(None, x.to(DEV), None, None, sh.to(DEV), S, None, [w.to(DEV) for w in ws],
                                          [b.to(DEV) for b in bs], beta.to(DEV), 1e-4)
    rf, rs, ra = _field_ref(x, sh, S, ws, bs, beta)
    assert rel_err(feat, rf) <= 1e-5
    assert rel_err(sdf, rs) <= 1e-5
    assert float((alpha.cpu() - ra).abs().max()) <= 2e-5
    assert outside_bar(feat, rf) == 0 and outside_bar(sdf, rs) == 0 and outside_bar(alpha, ra) == 0


@pytest.mark.parametrize("N,S", [(8, 48), (600, 48), (129, 33)])
def test_field_fused_backward_rows(N, S):
    from neuradar_b200 import functional as Fn

    x, sh, ws, bs, beta = _field_inputs(N, S, seed=7 * N + S)
    g = torch.Generator().manual_seed(1)
    M = N * S
    gf, gs, ga = torch.randn((M, 32), generator=g), torch.randn((M,), generator=g), torch.randn((M,), generator=g)
    # a ReLU whose pre-activation is within rounding of zero may legitimately switch between two fp32 evaluation
    # orders; rows with such a unit get no upstream gradient so that the comparison is about arithmetic, not kinks
    with torch.no_grad():
        xd64 = x.double()
        p0 = torch.nn.functional.linear(xd64, ws[0].double(), bs[0].double())
        geo = torch.nn.functional.linear(torch.relu(p0), ws[1].double(), bs[1].double())
        she = sh[:, None, :].expand(-1, S, -1).reshape(-1, 16).double()
        p2 = torch.nn.functional.linear(torch.cat([geo[:, 1:], she], -1), ws[2].double(), bs[2].double())
        p3 = torch.nn.functional.linear(torch.relu(p2), ws[3].double(), bs[3].double())
        near_kink = (torch.cat([p0, p2, p3], -1).abs().min(dim=-1).values < 1e-5)
    gf[near_kink] = 0
    gs[near_kink] = 0
    ga[near_kink] = 0
    # reference
    xr = x.clone().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs]
    betar = beta.clone().requires_grad_(True)
    rf, rs, ra = _field_ref(xr, sh, S, wr, br, betar)
    ((rf * gf).sum() + (rs * gs).sum() + (ra * ga).sum()).backward()
    # kernel
    xd = x.to(DEV).requires_grad_(True)
    wd = [w.to(DEV).requires_grad_(True) for w in ws]
    bd = [b.to(DEV).requires_grad_(True) for b in bs]
    betad = beta.to(DEV).requires_grad_(True)
    f, s_, a = Fn.field_fused(None, xd, None, None, sh.to(DEV), S, None, wd, bd, betad, 1e-4)
    assert rel_err(f, rf) <= 1e-5
    ((f * gf.to(DEV)).sum() + (s_ * gs.to(DEV)).sum() + (a * ga.to(DEV)).sum()).backward()
    tol = 1e-4  # bf16 hi/mid operands: ~2^-17 per product, far inside the 1e-3 bar
    assert rel_err(xd.grad, xr.grad) <= tol
    assert outside_bar(xd.grad, xr.grad) == 0
    for k in range(5):
        assert rel_err(wd[k].grad, wr[k].grad) <= tol, f"dW{k}"
        assert rel_err(bd[k].grad, br[k].grad) <= tol, f"db{k}"
        assert outside_bar(wd[k].grad, wr[k].grad) == 0 and outside_bar(bd[k].grad, br[k].grad) == 0, k
    assert rel_err(betad.grad, betar.grad) <= tol
