import numpy as np, torch
from oracle import neuradar_oracle as O
from neuradar_b200 import losses as L
g = {k: torch.from_numpy(v) for k, v in np.load("tests/golden/losses.npz").items()}
c, w = g["sbins2"], g["w2"][..., 0]
print("c range", float(c.min()), float(c.max()), "min width", float((c[:,1:]-c[:,:-1]).min()))
for i, r in enumerate((0.03, 0.003)):
    cp, wp = g[f"sbins{i}"], g[f"w{i}"][..., 0]
    ref = torch.stack([O.zipnerf_interlevel_loss(c[n:n+1], w[n:n+1], [(cp[n:n+1], wp[n:n+1])], pulse_widths=(r,)) for n in range(c.shape[0])])
    out = L.interlevel_per_ray(c.cuda(), w.cuda(), cp.cuda(), wp.cuda(), r).cpu()
    d = (out - ref).abs()
    idx = torch.argsort(d, descending=True)[:5]
    print("round", i, "sum ref", float(ref.mean()), "out", float(out.mean()))
    for n in idx.tolist():
        wd = c[n,1:]-c[n,:-1]
        print("  ray", n, "ref", float(ref[n]), "out", float(out[n]), "minwidth", float(wd.min()), "cmax", float(c[n].max()), "cpmax", float(cp[n].max()), "cpmin", float(cp[n].min()), "wsum", float(w[n].sum()))
