// Radar ray generation: Radars._generate_rays_from_fov (nerfstudio/cameras/radars.py:268-357) for
// a list of scans in one launch instead of a python loop of arange / meshgrid / cat per scan.
//   scan s, ray (ia, ie) on its azimuth x elevation grid (azimuth-major, torch.meshgrid(indexing="ij")):
//     az = min_azimuth + ia * azimuth_step,  el = min_elevation + ie * elevation_step          (torch.arange)
//     d_local = (cos el cos az, cos el sin az, sin el)
//     p = R d_local + t;  v = p - t;  norm = max(|v|, eps);  direction = v / norm            (:312-319, camera_utils.py:596)
//     origin = t;  pixel_area = (azimuth_step / 5) * (elevation_step / 5)                      (:322-328)
// The + t - t round trip of the reference is kept (it costs an ulp of |t| in the direction).
#include "common.cuh"

namespace nrb {

struct RadarScansDev {
  const float* r2w;       // [R,3,4]
  const float* min_az;    // [R]
  const float* az_step;   // [R]
  const float* min_el;    // [R]
  const float* el_step;   // [R]
};

__global__ void __launch_bounds__(256) radar_rays_kernel(const __grid_constant__ RadarScansDev sc,
                                                         const int64_t* __restrict__ scan_indices,
                                                         const int64_t* __restrict__ ray_offsets,  // [n_scans + 1]
                                                         const int32_t* __restrict__ n_elevations,  // [n_scans]
                                                         int n_scans, float* __restrict__ origins,
                                                         float* __restrict__ directions, float* __restrict__ pixel_area,
                                                         float* __restrict__ spher, float* __restrict__ norm_out,
                                                         int64_t* __restrict__ ray_scan, int64_t total) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= total) return;
  int lo = 0, hi = n_scans - 1;  // last scan whose first ray is <= r
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ray_offsets[mid] <= r) lo = mid; else hi = mid - 1;
  }
  const int64_t s = scan_indices[lo];
  const int local = static_cast<int>(r - ray_offsets[lo]);
  const int ne = n_elevations[lo];
  const int ia = local / ne, ie = local - ia * ne;
  // torch.arange on the device evaluates start + step * i in fp32 (one fused multiply-add)
  const float az = fmaf(sc.az_step[s], static_cast<float>(ia), sc.min_az[s]);
  const float el = fmaf(sc.el_step[s], static_cast<float>(ie), sc.min_el[s]);
  const float ce = cosf(el);
  const float lx = mul(ce, cosf(az)), ly = mul(ce, sinf(az)), lz = sinf(el);
  const float* m = sc.r2w + s * 12;
  const float tx = m[3], ty = m[7], tz = m[11];
  const float px = add(fmaf(m[2], lz, fmaf(m[1], ly, mul(m[0], lx))), tx);
  const float py = add(fmaf(m[6], lz, fmaf(m[5], ly, mul(m[4], lx))), ty);
  const float pz = add(fmaf(m[10], lz, fmaf(m[9], ly, mul(m[8], lx))), tz);
  const float vx = sub(px, tx), vy = sub(py, ty), vz = sub(pz, tz);
  const float nrm = fmaxf(sqrtf(fmaf(vz, vz, fmaf(vy, vy, mul(vx, vx)))), 1.1920928955078125e-7f);  // _EPS = finfo(float32).eps
  origins[3 * r] = tx, origins[3 * r + 1] = ty, origins[3 * r + 2] = tz;
  directions[3 * r] = __fdiv_rn(vx, nrm), directions[3 * r + 1] = __fdiv_rn(vy, nrm), directions[3 * r + 2] = __fdiv_rn(vz, nrm);
  // (torch divides a CUDA tensor by a python scalar as a multiplication with its reciprocal: x / 5 == x * 0.2f there)
  pixel_area[r] = mul(mul(sc.az_step[s], 0.2f), mul(sc.el_step[s], 0.2f));
  spher[2 * r] = az, spher[2 * r + 1] = el;
  norm_out[r] = nrm;
  ray_scan[r] = s;
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_radar_rays(const float* radar_to_worlds, const float* min_azimuth, const float* azimuth_step,
                              const float* min_elevation, const float* elevation_step, const int64_t* scan_indices,
                              const int64_t* ray_offsets, const int32_t* n_elevations, int32_t n_scans, int64_t total_rays,
                              float* origins, float* directions, float* pixel_area, float* directions_spher,
                              float* directions_norm, int64_t* ray_scan, nrb_stream_t stream) {
  NRB_REQUIRE(radar_to_worlds && min_azimuth && azimuth_step && min_elevation && elevation_step && scan_indices &&
                  ray_offsets && n_elevations && origins && directions && pixel_area && directions_spher &&
                  directions_norm && ray_scan,
              NRB_ERR_BAD_ARG, "nrb_radar_rays: null pointer");
  NRB_REQUIRE(n_scans >= 0 && total_rays >= 0, NRB_ERR_BAD_ARG, "nrb_radar_rays: negative size");
  if (n_scans == 0 || total_rays == 0) return NRB_OK;
  const RadarScansDev sc{radar_to_worlds, min_azimuth, azimuth_step, min_elevation, elevation_step};
  radar_rays_kernel<<<blocks_for(total_rays, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sc, scan_indices, ray_offsets, n_elevations, n_scans, origins, directions, pixel_area, directions_spher,
      directions_norm, ray_scan, total_rays);
  return finish_launch("nrb_radar_rays");
}
