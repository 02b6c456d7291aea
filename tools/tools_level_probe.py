"""GPU probe: per-level cost of hash bwd in the production mapping (16 same-res levels)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import neuradar_b200 as nb
from neuradar_b200 import functional as F, _lib
from tests.parity_utils import synthetic_rays, build_hot_path, make_ray_bundle

dev = "cuda"
model = build_hot_path(device=dev)
model.eval()
rays = synthetic_rays(65536, seed=42)
with torch.no_grad():
    out_rs = model._get_ray_samples(make_ray_bundle(rays, dev))[0]
rd, iv = out_rs.per_ray()
x, std = F.frustum_gaussians(rd, iv, 100.0)
M = x.shape[0]

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

Fdim = 2
for res in (16, 21, 27, 36, 48, 64, 84, 111, 147, 194, 256, 445, 1024):
    enc = nb.HashEncoding(num_levels=16, min_res=res, max_res=res, log2_hashmap_size=19, features_per_level=Fdim).to(dev)
    spec = F.GridSpec(16, Fdim, 19, tuple([float(res)] * 16))
    table = enc.hash_table.detach()
    dy = torch.randn((M, 16 * Fdim), device=dev)
    dtable = torch.zeros_like(table)
    g = spec.struct(table)
    out = torch.empty((M, 16 * Fdim), device=dev)
    def fwd():
        _lib.call("nrb_hash_fwd", C.byref(g), x.data_ptr(), std.data_ptr(), out.data_ptr(), M, _lib.stream_ptr())
    def bwd():
        _lib.call("nrb_hash_bwd", C.byref(g), x.data_ptr(), std.data_ptr(), dy.data_ptr(), dtable.data_ptr(), None, M, _lib.stream_ptr())
    print(f"16 levels @ res={res:5d}: fwd {timeit(fwd):7.3f} ms  bwd {timeit(bwd):7.3f} ms  -> per level fwd {timeit(fwd)/16*1e3:6.1f} us bwd {timeit(bwd)/16*1e3:7.1f} us")
