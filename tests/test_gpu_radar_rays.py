"""GPU tests of the radar ray generator: `neuradar_b200.radars.generate_rays_from_fov` against the
outputs of the reference's `Radars._generate_rays_from_fov` (tests/golden/radar_rays.npz, produced on the CPU), against the
oracle at size, and - where oracle/_ref is materialised - against the reference's method running on the same GPU."""
import math
import types

import pytest
import torch

from oracle import neuradar_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def nb():
    import neuradar_b200

    return neuradar_b200


def _sensor(g, device):
    keys = ("radar_azimuth_ray_divergence", "radar_elevation_ray_divergence", "min_azimuth", "max_azimuth", "min_elevation",
            "max_elevation")
    s = types.SimpleNamespace(radar_to_worlds=g["radar_to_worlds"].to(device), times=g["times_in"].to(device), metadata=None,
                              **{k: g[k].to(device) for k in keys})
    return s


def test_radar_rays_golden(nb, golden):
    from neuradar_b200 import radars as R

    g = golden("radar_rays")
    rb = R.generate_rays_from_fov(_sensor(g, DEV), g["scan_indices"])
    assert rb.origins.shape == g["origins"].shape
    assert torch.equal(rb.camera_indices.cpu(), g["camera_indices"])          # which scan every ray belongs to
    assert torch.equal(rb.origins.cpu(), g["origins"])
    # (x / 5 is x * 0.2f on the device and a true division on the host: one ulp)
    torch.testing.assert_close(rb.pixel_area.cpu(), g["pixel_area"], rtol=2.4e-7, atol=0)
    assert torch.equal(rb.times.cpu(), g["times"])
    assert torch.equal(rb.fars.cpu(), g["fars"])
    assert torch.equal(rb.metadata["did_return"].cpu(), g["did_return"])
    # torch.arange: fp32 fused multiply-add on the device, double on the host - equal for dyadic steps, an ulp otherwise
    assert float((rb.metadata["directions_spher"].cpu() - g["directions_spher"]).abs().max()) <= 1.2e-7
    # the reference forms R d + t and subtracts t again: the result carries a rounding of ulp(|t|) = 7.6e-6 at |t| ~ 100,
    # and which way it falls depends on the last bit of R d (matmul order), so that is the floor of this comparison
    assert float((rb.directions.cpu() - g["directions"]).abs().max()) <= 1e-5
    assert float((rb.metadata["directions_norm"].cpu() - g["directions_norm"]).abs().max()) <= 1e-5
    assert float((rb.directions.norm(dim=-1) - 1).abs().max()) <= 1e-6  # unit length regardless


def _random_sensor(n, seed):
    g = torch.Generator().manual_seed(seed)
    yaw = torch.rand(n, generator=g) * 2 * math.pi
    c, s, z, o = torch.cos(yaw), torch.sin(yaw), torch.zeros(n), torch.ones(n)
    R = torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)
    t = (torch.rand((n, 3, 1), generator=g) - 0.5) * 100
    d = dict(radar_azimuth_ray_divergence=torch.full((n, 1), 0.0625), radar_elevation_ray_divergence=torch.full((n, 1), 0.0625),
             min_azimuth=torch.full((n, 1), -0.5), max_azimuth=torch.full((n, 1), 0.5),
             min_elevation=torch.full((n, 1), -0.5), max_elevation=torch.full((n, 1), 0.5))
    return torch.cat([R, t], -1), torch.arange(n).float() * 0.05, d


def test_radar_rays_one_million_rays_vs_oracle(nb):
    """BASELINE config 5's sweep: 4096 scans x 256 rays in one launch (the reference loops over the scans in python)."""
    from neuradar_b200 import radars as R

    n = 4096
    r2w, times, d = _random_sensor(n, 3)
    sensor = types.SimpleNamespace(radar_to_worlds=r2w.to(DEV), times=times.to(DEV), metadata={"sensor_idxs": torch.full((n, 1), 2).to(DEV)},
                                   **{k: v.to(DEV) for k, v in d.items()})
    scans = torch.randperm(n, generator=torch.Generator().manual_seed(1))
    rb = R.generate_rays_from_fov(sensor, scans)
    assert rb.origins.shape == (n * 256, 3)
    ref = O.radar_rays(r2w, d["min_azimuth"].reshape(-1), d["max_azimuth"].reshape(-1), d["radar_azimuth_ray_divergence"].reshape(-1),
                       d["min_elevation"].reshape(-1), d["max_elevation"].reshape(-1), d["radar_elevation_ray_divergence"].reshape(-1), scans)
    assert torch.equal(rb.camera_indices[:, 0].cpu(), ref["ray_scan"])
    assert torch.equal(rb.origins.cpu(), ref["origins"])
    assert torch.equal(rb.metadata["directions_spher"].cpu(), ref["directions_spher"])   # dyadic grid: bit-exact
    assert float((rb.directions.cpu() - ref["directions"]).abs().max()) <= 1e-5
    assert torch.equal(rb.metadata["sensor_idxs"].cpu(), torch.full((n * 256, 1), 2))
    assert torch.equal(rb.times.cpu()[:, 0], times[ref["ray_scan"]])


@pytest.mark.skipif(not ref_shim.available(), reason="reference modules not materialised (oracle/build_ref.py)")
def test_radar_rays_vs_reference_on_the_same_gpu(nb, golden):
    from neuradar_b200 import radars as R

    ref_shim.install()
    from nerfstudio.cameras.radars import Radars

    g = golden("radar_rays")
    keys = ("radar_azimuth_ray_divergence", "radar_elevation_ray_divergence", "min_azimuth", "max_azimuth", "min_elevation",
            "max_elevation")
    radars = Radars(radar_to_worlds=g["radar_to_worlds"], times=g["times_in"][:, 0], **{k: g[k] for k in keys}).to(DEV)
    ref = radars._generate_rays_from_fov(g["scan_indices"])
    got = R.generate_rays_from_fov(radars, g["scan_indices"])        # the reference's own Radars object as input
    assert torch.equal(got.camera_indices, ref.camera_indices)
    assert torch.equal(got.origins, ref.origins) and torch.equal(got.pixel_area, ref.pixel_area)
    assert torch.equal(got.metadata["directions_spher"], ref.metadata["directions_spher"])  # same device arithmetic: bit-exact
    assert float((got.directions - ref.directions).abs().max()) <= 1e-5
    assert torch.equal(got.times, ref.times)
