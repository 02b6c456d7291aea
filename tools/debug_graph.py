import os, sys, traceback, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuradar_b200 as nb
from neuradar_b200.dist import GradArena
from neuradar_b200.synthetic import WORKLOADS, build_workload, synthetic_rays
dev = "cuda:0"
w = WORKLOADS[2]
n = 8192
model = build_workload(w, device=dev); model.train()
used = [p for name, p in model.named_parameters() if not name.startswith("proposal_fields.0")]
arena = GradArena(used, direct_scatter=True, early=[model.field.hashgrid.static_grid.hash_table])
rays = synthetic_rays(n, seed=1)
cur = {k: v.to(dev) for k, v in rays.items()}
res = {}
def step():
    arena.zero()
    rb = nb.RayBundle(origins=cur["origins"].clone(), directions=cur["directions"].clone(), pixel_area=cur["pixel_area"].clone(),
                      nears=cur["nears"].clone(), fars=cur["fars"].clone(), times=cur["times"],
                      metadata={"is_lidar": cur["is_lidar"], "is_radar": cur["is_radar"]})
    out = model(rb)
    loss = nb.bench_loss(out)
    loss.backward()
    arena.all_reduce()
    res["loss"] = loss.detach()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
import warnings
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as wl:
    warnings.simplefilter("always")
    step()
    torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("default")
print("SYNC WARNINGS:", len(wl))
for wmsg in wl:
    print("  ", wmsg.filename, wmsg.lineno, str(wmsg.message)[:100])
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=side):
        step()
    g.replay(); torch.cuda.synchronize()
    print("captured ok, loss", float(res["loss"]))
    a = float(res["loss"]); g.replay(); torch.cuda.synchronize(); print("replay 2", float(res["loss"]))
except Exception:
    tb = traceback.format_exc()
    print(tb[:6000])
