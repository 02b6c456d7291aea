"""Scratch: time nrb_hash_bwd for the main grid under the current NRB_BWD_* environment."""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuradar_b200 as nb
from neuradar_b200 import functional as F, _lib
from tests.parity_utils import synthetic_rays, build_hot_path, make_ray_bundle
dev = "cuda"
model = build_hot_path(device=dev); model.eval()
rays = synthetic_rays(65536, seed=42)
with torch.no_grad():
    rs = model._get_ray_samples(make_ray_bundle(rays, dev))[0]
rd, iv = rs.per_ray()
x, std = F.frustum_gaussians(rd, iv, 100.0)
M = x.shape[0]
enc = model.field.hashgrid.static_grid
table = enc.hash_table.detach(); g = enc.spec.struct(table)
dy = torch.randn((M, 32), device=dev); dtable = torch.zeros_like(table)
nbytes = int(_lib.load().nrb_hash_bwd_workspace_bytes(C.byref(g), M))
ws = torch.empty((max(nbytes, 16),), device=dev, dtype=torch.uint8)
def bwd():
    _lib.call("nrb_hash_bwd", C.byref(g), x.data_ptr(), std.data_ptr(), dy.data_ptr(), dtable.data_ptr(), None, M, ws.data_ptr(), nbytes, _lib.stream_ptr())
for _ in range(2): bwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): bwd()
e1.record(); torch.cuda.synchronize()
print({k: v for k, v in os.environ.items() if k.startswith("NRB_")}, f"workspace {nbytes/2**20:.1f} MiB  bwd {e0.elapsed_time(e1)/5:.3f} ms")
