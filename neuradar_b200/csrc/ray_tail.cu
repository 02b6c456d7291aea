// Per-ray tail of the hot path: the point heads and the lidar-carving terms that NeuRadarModel.get_nff_outputs emits in
// training.
//   * point heads: a rendered depth becomes a 3-D point per ray.  Lidar (and camera) rays: p = o + d * depth, optionally
//     taken to the sensor frame (nerfstudio/models/ad_model.py:103-108); radar rays: p = depth * (cos phi cos theta,
//     sin phi cos theta, sin theta) from the ray's spherical direction (phi, theta) = directions_spher
//     (nerfstudio/models/neuradar.py:463-473,1025-1029) - the input of the radar transformer's positional embedding.
//   * carving: is_close_to_lidar per sample (nerfstudio/models/neuradar.py:971-994) and the proposal carving loss
//     sum((w * (is_lidar & ~is_close_to_lidar))^2) (:527-531) with its gradient, one thread per sample.
#include "common.cuh"

namespace nrb {

__global__ void __launch_bounds__(256) point_heads_fwd_kernel(const float* __restrict__ origins,
                                                              const float* __restrict__ directions,
                                                              const float* __restrict__ depth,
                                                              const uint8_t* __restrict__ is_radar,
                                                              const float* __restrict__ spher,  // [N,2] (phi, theta)
                                                              const float* __restrict__ world2sensor,  // [3,4] or null
                                                              float* __restrict__ points, int64_t N) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float t = depth[n];
  float x, y, z;
  if (is_radar != nullptr && is_radar[n]) {
    const float phi = spher[2 * n], theta = spher[2 * n + 1];
    const float ct = cosf(theta);
    x = t * cosf(phi) * ct;
    y = t * sinf(phi) * ct;
    z = t * sinf(theta);
  } else {
    x = origins[3 * n] + directions[3 * n] * t;
    y = origins[3 * n + 1] + directions[3 * n + 1] * t;
    z = origins[3 * n + 2] + directions[3 * n + 2] * t;
    if (world2sensor != nullptr) {
      const float* m = world2sensor;
      const float px = m[0] * x + m[1] * y + m[2] * z + m[3];
      const float py = m[4] * x + m[5] * y + m[6] * z + m[7];
      const float pz = m[8] * x + m[9] * y + m[10] * z + m[11];
      x = px, y = py, z = pz;
    }
  }
  points[3 * n] = x;
  points[3 * n + 1] = y;
  points[3 * n + 2] = z;
}

__global__ void __launch_bounds__(256) point_heads_bwd_kernel(const float* __restrict__ directions,
                                                              const uint8_t* __restrict__ is_radar,
                                                              const float* __restrict__ spher,
                                                              const float* __restrict__ world2sensor,
                                                              const float* __restrict__ dpoints, float* __restrict__ ddepth,
                                                              int64_t N) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float ux, uy, uz;  // d point / d depth
  if (is_radar != nullptr && is_radar[n]) {
    const float phi = spher[2 * n], theta = spher[2 * n + 1];
    const float ct = cosf(theta);
    ux = cosf(phi) * ct, uy = sinf(phi) * ct, uz = sinf(theta);
  } else {
    ux = directions[3 * n], uy = directions[3 * n + 1], uz = directions[3 * n + 2];
    if (world2sensor != nullptr) {
      const float* m = world2sensor;
      const float rx = m[0] * ux + m[1] * uy + m[2] * uz, ry = m[4] * ux + m[5] * uy + m[6] * uz;
      const float rz = m[8] * ux + m[9] * uy + m[10] * uz;
      ux = rx, uy = ry, uz = rz;
    }
  }
  ddepth[n] = dpoints[3 * n] * ux + dpoints[3 * n + 1] * uy + dpoints[3 * n + 2] * uz;
}

// One thread per sample.  mode 0: write is_close [N,S] and, when weights are given, loss_per_sample = (w * mask)^2 with
// mask = is_lidar & ~is_close; mode 1 (backward): dweights = dloss * 2 * w * mask.
__global__ void __launch_bounds__(256) carving_kernel(const float* __restrict__ starts, const float* __restrict__ ends,
                                                      int64_t row_stride, int S, const uint8_t* __restrict__ is_lidar,
                                                      const float* __restrict__ dir_norm,
                                                      const uint8_t* __restrict__ did_return, float carving_eps,
                                                      float non_return_dist, const float* __restrict__ weights,
                                                      const float* __restrict__ dloss, int backward,
                                                      uint8_t* __restrict__ is_close, float* __restrict__ out, int64_t total) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int64_t n = gid / S;
  const int s = static_cast<int>(gid - n * S);
  bool close = false;
  const bool lidar = is_lidar[n] != 0;
  if (lidar) {
    const float mid = (starts[n * row_stride + s] + ends[n * row_stride + s]) * 0.5f;
    const bool hit = fabsf(dir_norm[n] - mid) < carving_eps;
    if (did_return != nullptr) {
      const bool ret = did_return[n] != 0;
      close = (ret && hit) || (!ret && mid < non_return_dist);
    } else {
      close = hit;
    }
  }
  const bool mask = lidar && !close;
  if (!backward) {
    if (is_close != nullptr) is_close[gid] = close ? 1 : 0;
    if (out != nullptr) {
      const float w = mask ? weights[gid] : 0.0f;
      out[gid] = w * w;
    }
  } else {
    out[gid] = mask ? 2.0f * weights[gid] * dloss[0] : 0.0f;
  }
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_point_heads_fwd(const float* origins, const float* directions, const float* depth,
                                   const uint8_t* is_radar, const float* directions_spher, const float* world2sensor,
                                   float* points, int64_t N, nrb_stream_t stream) {
  NRB_REQUIRE(origins && directions && depth && points && N >= 0, NRB_ERR_BAD_ARG, "nrb_point_heads_fwd: null pointer");
  NRB_REQUIRE(is_radar == nullptr || directions_spher != nullptr, NRB_ERR_BAD_ARG,
              "nrb_point_heads_fwd: radar rays need directions_spher");
  if (N == 0) return NRB_OK;
  point_heads_fwd_kernel<<<blocks_for(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      origins, directions, depth, is_radar, directions_spher, world2sensor, points, N);
  return finish_launch("nrb_point_heads_fwd");
}

extern "C" int nrb_point_heads_bwd(const float* directions, const uint8_t* is_radar, const float* directions_spher,
                                   const float* world2sensor, const float* dpoints, float* ddepth, int64_t N,
                                   nrb_stream_t stream) {
  NRB_REQUIRE(directions && dpoints && ddepth && N >= 0, NRB_ERR_BAD_ARG, "nrb_point_heads_bwd: null pointer");
  if (N == 0) return NRB_OK;
  point_heads_bwd_kernel<<<blocks_for(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      directions, is_radar, directions_spher, world2sensor, dpoints, ddepth, N);
  return finish_launch("nrb_point_heads_bwd");
}

extern "C" int nrb_lidar_carving(const nrb_intervals_t* iv, int64_t N, const uint8_t* is_lidar, const float* directions_norm,
                                 const uint8_t* did_return, float carving_epsilon, float non_return_lidar_distance,
                                 const float* weights, const float* dloss, uint8_t* is_close, float* out,
                                 nrb_stream_t stream) {
  if (int rc = check_intervals("nrb_lidar_carving", iv)) return rc;
  NRB_REQUIRE(is_lidar && directions_norm && N >= 0, NRB_ERR_BAD_ARG, "nrb_lidar_carving: null pointer");
  NRB_REQUIRE(dloss == nullptr || (weights != nullptr && out != nullptr), NRB_ERR_BAD_ARG,
              "nrb_lidar_carving: the backward needs weights and an output");
  NRB_REQUIRE(out == nullptr || weights != nullptr, NRB_ERR_BAD_ARG, "nrb_lidar_carving: a loss output needs weights");
  if (N == 0) return NRB_OK;
  const int64_t total = N * iv->num_samples;
  carving_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      iv->starts, iv->ends, iv->row_stride, iv->num_samples, is_lidar, directions_norm, did_return, carving_epsilon,
      non_return_lidar_distance, weights, dloss, dloss != nullptr ? 1 : 0, is_close, out, total);
  return finish_launch("nrb_lidar_carving");
}
