"""Scratch: explore tcgen05 descriptor conventions (majors, M=64 layout, LBO/SBO roles) on the GPU."""
import itertools, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuradar_b200 import functional as Fn
torch.manual_seed(0)
dev = "cuda"
P = torch.randint(-3, 4, (128, 32)).float()
Q = torch.randint(-3, 4, (128, 32)).float()
Pd, Qd = P.to(dev), Q.to(dev)

def show(tag, cfg, expect_fn):
    d = Fn.tc_probe(Pd, Qd, cfg).cpu()
    res = expect_fn(d)
    print(f"{tag:40s} cfg={cfg} nonzero={int((d!=0).sum())} -> {res}")
    return d

# 1. baseline K-major/K-major M=128 N=32: D = P[:, :8k] Q[:, :8k]^T  (rows of Q = N)
def kk(d, ks=4):
    ref = P[:, :8*ks] @ Q[:32, :8*ks].T
    return "OK" if torch.equal(d, ref) else f"maxdiff {float((d-ref).abs().max())}"
show("KK M128", [0,0,128,32, 128,1024,256, 128,1024,256, 4], kk)

# 2. A K-major, B MN-major, M=128: D[s, n] = sum_{r<8*ks} P[s, r] * Q[r, n]   (Q rows = reduction)
def kmn(d, ks=4):
    ref = P[:, :8*ks] @ Q[:8*ks, :32]
    return "OK" if torch.equal(d, ref) else f"maxdiff {float((d-ref).abs().max())}"
for (lbo, sbo) in [(1024,128),(128,1024),(16,128),(128,16),(1024,16),(16,1024)]:
    show("K/MN M128 b(lbo,sbo)=%s"%((lbo,sbo),), [0,1,128,32, 128,1024,256, lbo,sbo,1024, 4], kmn)

# 3. both MN-major M=64: D[j, n] = sum_{s<8*ks} P[s, j] Q[s, n]
def tn(d, ks=16):
    ref = P[:8*ks].T @ Q[:8*ks]          # [32, 32]
    # find where rows of ref appear in d
    hits = []
    for lane in range(128):
        for j in range(32):
            if torch.equal(d[lane], ref[j]): hits.append((lane, j))
    return f"{len(hits)} rows matched; first {hits[:6]}"
for (lbo, sbo) in [(1024,128),(128,1024)]:
    show("MN/MN M64 (lbo,sbo)=%s"%((lbo,sbo),), [1,1,64,32, lbo,sbo,1024, lbo,sbo,1024, 16], tn)
    show("MN/MN M128 (lbo,sbo)=%s"%((lbo,sbo),), [1,1,128,32, lbo,sbo,1024, lbo,sbo,1024, 16], tn)
# 4. M=64 K-major both: D[r, n] = P[r,:] Q[n,:]^T for r < 64
def kk64(d, ks=4):
    ref = P[:64, :8*ks] @ Q[:32, :8*ks].T
    hits = []
    for lane in range(128):
        for j in range(64):
            if torch.equal(d[lane], ref[j]): hits.append((lane, j))
    return f"{len(hits)} rows matched; first {hits[:5]} last {hits[-3:]}"
show("KK M64", [0,0,64,32, 128,1024,256, 128,1024,256, 4], kk64)
