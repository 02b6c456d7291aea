"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, argument errors are
reported without touching a GPU, the product refuses to run without CUDA, and the host containers behave like the
reference's TensorDataclass (tests/utils/test_tensor_dataclass.py, tests/cameras/test_rays.py)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "neuradar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from neuradar_b200 import _lib

    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/neuradar_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures and header disagree"
    assert lib.nrb_version() == 100
    assert lib.nrb_launch_count() == 0


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "neuradar_b200.h"\nint main(void){ nrb_grid_t g; (void)g; return NRB_VERSION == 100 ? 0 : 1; }\n')
    import subprocess

    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                    str(tmp_path / "t")], check=True)
    subprocess.run([str(tmp_path / "t")], check=True)


def test_argument_errors_are_reported_without_launching():
    from neuradar_b200 import _lib

    lib = _lib.load()
    g = _lib.Grid()
    rc = lib.nrb_hash_fwd(ctypes.byref(g), None, None, None, 10, None)
    assert rc == -1 and b"null" in lib.nrb_last_error_string()
    buf = (ctypes.c_float * 64)()
    g.table = ctypes.addressof(buf)
    g.num_levels, g.features_per_level, g.log2_hashmap_size = 16, 3, 19
    assert lib.nrb_hash_fwd(ctypes.byref(g), ctypes.addressof(buf), None, ctypes.addressof(buf), 1, None) == -3
    g.features_per_level, g.num_levels = 2, 17
    assert lib.nrb_hash_fwd(ctypes.byref(g), ctypes.addressof(buf), None, ctypes.addressof(buf), 1, None) == -1
    iv = _lib.Intervals()
    iv.starts = iv.ends = ctypes.addressof(buf)
    iv.row_stride, iv.num_samples = 4, 300
    assert lib.nrb_density_weights_fwd(ctypes.addressof(buf), ctypes.byref(iv), 1, ctypes.addressof(buf), None) == -1
    assert b"num_samples" in lib.nrb_last_error_string()
    m = _lib.Mlp()
    m.num_layers = 5
    assert lib.nrb_mlp_fwd(ctypes.byref(m), ctypes.addressof(buf), ctypes.addressof(buf), None, 1, None) == -1
    assert lib.nrb_launch_count() == 0
    # M == 0 is a valid no-op
    g.num_levels = 16
    assert lib.nrb_hash_fwd(ctypes.byref(g), ctypes.addressof(buf), None, ctypes.addressof(buf), 0, None) == 0


def test_no_cpu_path():
    import neuradar_b200 as nb
    from neuradar_b200._lib import NeuradarB200Error

    enc = nb.HashEncoding(num_levels=2, log2_hashmap_size=4)
    with pytest.raises(NeuradarB200Error):
        enc(torch.rand((4, 3)))
    mlp = nb.MLP(in_dim=4, num_layers=2, layer_width=8, out_dim=2)
    with pytest.raises(NeuradarB200Error):
        mlp(torch.rand((4, 4)))


def test_losses_and_optimizer_have_no_cpu_path():
    """SURVEY 8f next-1 / next-2 entry points: CPU tensors are refused, nothing falls back to torch."""
    from types import SimpleNamespace

    from neuradar_b200 import losses
    from neuradar_b200.optim import FusedAdam, FusedAdamW

    sb = torch.linspace(0, 1, 9).repeat(3, 1)
    rs = SimpleNamespace(spacing_starts=sb[:, :-1, None], spacing_ends=sb[:, 1:, None])
    w = torch.full((3, 8, 1), 0.1, requires_grad=True)
    with pytest.raises(RuntimeError):
        losses.distortion_loss([w], [rs])
    with pytest.raises(RuntimeError):
        losses.zipnerf_interlevel_loss([w, w], [rs, rs])
    for cls in (FusedAdam, FusedAdamW):
        with pytest.raises(ValueError):
            cls([torch.nn.Parameter(torch.zeros(8))], lr=1e-2)
    # the reference's sdist helper (losses.py:107-112) is plain indexing and works anywhere
    assert torch.equal(losses.ray_samples_to_sdist(rs), sb)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neuradar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text or "import oracle" not in text and "from oracle" not in text, f


def test_frustum_positions_known_answer():
    """reference tests/cameras/test_rays.py:11-30"""
    import neuradar_b200 as nb

    fr = nb.Frustums(origins=torch.ones((5, 3)), directions=torch.tensor([[0.0, 1.0, 0.0]]).expand(5, 3).contiguous(),
                     starts=torch.ones((5, 1)) * 2, ends=torch.ones((5, 1)) * 3, pixel_area=torch.ones((5, 1)))
    assert torch.allclose(fr.get_positions(), torch.tensor([1.0, 3.5, 1.0]).expand(5, 3))
    assert fr.shape == (5,)


def test_tensor_dataclass_semantics():
    import neuradar_b200 as nb

    N, S = 7, 5
    rb = nb.RayBundle(origins=torch.rand((N, 3)), directions=torch.rand((N, 3)), pixel_area=torch.rand((N, 1)),
                      nears=torch.zeros((N, 1)), fars=torch.ones((N, 1)), times=torch.rand((N, 1)),
                      metadata={"is_lidar": torch.zeros((N, 1), dtype=torch.bool)})
    assert len(rb) == N and rb.shape == (N,)
    assert rb[2:5].shape == (3,) and rb[2:5].metadata["is_lidar"].shape == (3, 1)
    bins = torch.linspace(0, 1, S + 1).expand(N, S + 1).contiguous()
    rs = rb.get_ray_samples(bin_starts=bins[:, :-1, None], bin_ends=bins[:, 1:, None],
                            spacing_starts=bins[:, :-1, None], spacing_ends=bins[:, 1:, None])
    assert rs.shape == (N, S)
    # broadcast views, not copies
    assert rs.frustums.origins.shape == (N, S, 3) and rs.frustums.origins.stride() == (3, 0, 1)
    assert rs.frustums.pixel_area.stride() == (1, 0, 1)
    assert rs.frustums.starts.stride() == (S + 1, 1, 1)
    assert rs.times.shape == (N, S, 1) and rs.metadata["is_lidar"].shape == (N, S, 1)
    assert torch.equal(rs.deltas, rs.frustums.ends - rs.frustums.starts)
    cut = rs[..., :-1]
    assert cut.shape == (N, S - 1) and cut.frustums.ends.shape == (N, S - 1, 1)
    assert rs[3].shape == (S,)
    assert rs.flatten().shape == (N * S,) and rs.reshape((S, N)).shape == (S, N)
    pos = rs.frustums.get_positions()
    assert pos.shape == (N, S, 3)


def test_shard_bounds_cover_the_batch():
    from neuradar_b200.dist import shard_bounds

    for world in (1, 2, 4, 8):
        cuts = [shard_bounds(262144, world, r, granule=256) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == 262144
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all((e - s) % 256 == 0 for s, e in cuts)
    assert shard_bounds(10 * 1024, 4, 0, granule=1024) == (0, 3072)
    assert shard_bounds(10 * 1024, 4, 3, granule=1024) == (8192, 10240)
    with pytest.raises(ValueError):
        shard_bounds(1000, 2, 0, granule=256)


def test_radar_fov_grid_sizes_match_torch_arange():
    """Host arithmetic of the radar ray generator: rays per scan = len(torch.arange(min, max, step)) along each axis, also
    where (max - min) / step lands within rounding of an integer in fp32."""
    import types

    from neuradar_b200.radars import fov_grid_sizes

    g = torch.Generator().manual_seed(0)
    n = 64
    lo = -torch.rand((n, 1), generator=g)
    step = torch.rand((n, 1), generator=g) * 0.1 + 0.01
    k = torch.randint(3, 40, (n, 1), generator=g).float()
    hi = lo + step * k + (torch.rand((n, 1), generator=g) - 0.5) * 1e-7  # straddles exact multiples
    sensor = types.SimpleNamespace(min_azimuth=lo, max_azimuth=hi, radar_azimuth_ray_divergence=step,
                                   min_elevation=lo * 0.5, max_elevation=hi * 0.5, radar_elevation_ray_divergence=step)
    grid = fov_grid_sizes(sensor)
    for i in range(n):
        assert int(grid["n_az"][i]) == torch.arange(float(lo[i]), float(hi[i]), float(step[i])).numel(), i
        assert int(grid["n_el"][i]) == torch.arange(float(lo[i] * 0.5), float(hi[i] * 0.5), float(step[i])).numel(), i


def test_workload_roofline_arithmetic_matches_the_survey():
    """bytes / FLOPs per ray that bench.py divides by."""
    from neuradar_b200.synthetic import WORKLOADS

    assert abs(WORKLOADS[2].bytes_per_ray() - 142_780) <= 300          # 2 x 70 656 gather/scatter + ray I/O
    assert WORKLOADS[1].bytes_per_ray() == WORKLOADS[2].bytes_per_ray() == WORKLOADS[3].bytes_per_ray()
    assert abs(WORKLOADS[5].bytes_per_ray() - 70_832) <= 10            # inference: gathers + 176 B of ray I/O
    assert abs(WORKLOADS[4].bytes_per_ray() - 356_000) <= 2_000        # 2 x 176 947 + I/O
    assert WORKLOADS[2].mlp_flop_per_ray() == 3 * 11_328 * 48          # 1.63 MFLOP per ray, fwd + bwd


def test_peer_memory_is_only_used_for_single_box_nccl_jobs():
    from neuradar_b200.dist import order_early_first, peer_memory_or_none

    assert peer_memory_or_none(1024, "cpu") is None                    # no process group, not a CUDA device
    a, b, c = (torch.nn.Parameter(torch.zeros(2)) for _ in range(3))
    ordered, n_first = order_early_first([a, b, c], [c])
    assert ordered[0] is c and n_first == 1 and ordered[1:] == [a, b]
