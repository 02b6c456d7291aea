// Multiresolution hash grid: forward gather, index dump, backward scatter, and the sample-gaussian producer.
// Semantics: HashEncoding.pytorch_fwd / hash_fn (nerfstudio/field_components/encodings.py:406-466) and
// NeuRADHashEncoding._rescale_grid_features (nerfstudio/field_components/neurad_encoding.py:309-316).
//
// Work decomposition: one thread per (point, level) pair with the level index fastest.  The 2^19..2^22-row tables
// are hit at random, so there is nothing to coalesce on the gather side; what can be coalesced is the [M, L*F]
// feature matrix, and with the level fastest a warp writes (reads, in the backward pass) one contiguous
// 32*F*4-byte span of it.  Each thread issues its 8 independent F-wide vector gathers before touching any of them.
#include <algorithm>

#include "common.cuh"

namespace nrb {

struct GridDev {
  const float* table;
  float scalings[NRB_MAX_LEVELS];
  int num_levels;
  int log2_size;
};

static GridDev to_dev(const nrb_grid_t* g) {
  GridDev d;
  d.table = g->table;
  for (int i = 0; i < NRB_MAX_LEVELS; ++i) d.scalings[i] = g->scalings[i];
  d.num_levels = g->num_levels;
  d.log2_size = g->log2_hashmap_size;
  return d;
}

template <int F>
__global__ void __launch_bounds__(256) hash_fwd_kernel(const __grid_constant__ GridDev g, const float* __restrict__ x,
                                                       const float* __restrict__ std, float* __restrict__ out,
                                                       int64_t total) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int64_t m = gid / g.num_levels;
  const int l = static_cast<int>(gid - m * g.num_levels);
  const float scal = g.scalings[l];
  const Cell c = locate_cell(__ldg(x + 3 * m), __ldg(x + 3 * m + 1), __ldg(x + 3 * m + 2), scal,
                             (1u << g.log2_size) - 1u);
  const float* base = g.table + (static_cast<size_t>(l) << g.log2_size) * F;
  float v[F];
  interpolate<F>(base, c, v);
  if (std != nullptr) {
    const float w = level_weight(scal, __ldg(std + m));
#pragma unroll
    for (int j = 0; j < F; ++j) v[j] *= w;
  }
  using V = typename Feat<F>::type;
  V o;
  if constexpr (F == 1) {
    o = v[0];
  } else if constexpr (F == 2) {
    o = make_float2(v[0], v[1]);
  } else {
    o = make_float4(v[0], v[1], v[2], v[3]);
  }
  reinterpret_cast<V*>(out)[gid] = o;
}

__global__ void __launch_bounds__(256) hash_indices_kernel(const __grid_constant__ GridDev g, const float* __restrict__ x,
                                                           int64_t* __restrict__ idx, int64_t total) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int64_t m = gid / g.num_levels;
  const int l = static_cast<int>(gid - m * g.num_levels);
  const Cell c = locate_cell(x[3 * m], x[3 * m + 1], x[3 * m + 2], g.scalings[l], (1u << g.log2_size) - 1u);
  const int64_t off = static_cast<int64_t>(l) << g.log2_size;
#pragma unroll
  for (int k = 0; k < 8; ++k) idx[gid * 8 + k] = off + c.row[k];
}

template <int F>
__device__ __forceinline__ void scatter_row(float* __restrict__ base, uint32_t row, const float g[F], float w) {
  float* p = base + static_cast<size_t>(row) * F;
  if constexpr (F == 1) {
    atomicAdd(p, g[0] * w);
  } else if constexpr (F == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(g[0] * w, g[1] * w));
  } else {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(g[0] * w, g[1] * w, g[2] * w, g[3] * w));
  }
}

// Levels [level0, L) with plain vector reductions into the table (levels below level0 are handled by the
// shared-memory privatised kernel further down).
template <int F, bool kNeedDx>
__global__ void __launch_bounds__(256) hash_bwd_kernel(const __grid_constant__ GridDev g, const float* __restrict__ x,
                                                       const float* __restrict__ std, const float* __restrict__ dy,
                                                       float* __restrict__ dtable, float* __restrict__ dx,
                                                       int64_t total, int level0) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int nl = g.num_levels - level0;
  const int64_t m = gid / nl;
  const int l = level0 + static_cast<int>(gid - m * nl);
  const float scal = g.scalings[l];
  const Cell c = locate_cell(__ldg(x + 3 * m), __ldg(x + 3 * m + 1), __ldg(x + 3 * m + 2), scal,
                             (1u << g.log2_size) - 1u);
  using V = typename Feat<F>::type;
  const V gv = __ldg(reinterpret_cast<const V*>(dy) + m * g.num_levels + l);
  float gr[F];
  if constexpr (F == 1) {
    gr[0] = gv;
  } else if constexpr (F == 2) {
    gr[0] = gv.x;
    gr[1] = gv.y;
  } else {
    gr[0] = gv.x;
    gr[1] = gv.y;
    gr[2] = gv.z;
    gr[3] = gv.w;
  }
  if (std != nullptr) {
    const float lw = level_weight(scal, __ldg(std + m));
#pragma unroll
    for (int j = 0; j < F; ++j) gr[j] *= lw;
  }
  const size_t level_off = (static_cast<size_t>(l) << g.log2_size) * F;
  if constexpr (kNeedDx) {
    // d out / d offset needs the corner features themselves.
    float f[8][F];
#pragma unroll
    for (int k = 0; k < 8; ++k) load_row<F>(g.table + level_off, c.row[k], f[k]);
    const float ax = c.ox, bx = 1.0f - c.ox, ay = c.oy, by = 1.0f - c.oy, az = c.oz, bz = 1.0f - c.oz;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const float f03 = f[0][j] * ax + f[3][j] * bx, f12 = f[1][j] * ax + f[2][j] * bx;
      const float f56 = f[5][j] * ax + f[6][j] * bx, f47 = f[4][j] * ax + f[7][j] * bx;
      const float f0312 = f03 * ay + f12 * by, f4756 = f47 * ay + f56 * by;
      gz += gr[j] * (f0312 - f4756);
      gy += gr[j] * ((f03 - f12) * az + (f47 - f56) * bz);
      gx += gr[j] * (((f[0][j] - f[3][j]) * ay + (f[1][j] - f[2][j]) * by) * az +
                     ((f[4][j] - f[7][j]) * ay + (f[5][j] - f[6][j]) * by) * bz);
    }
    atomicAdd(dx + 3 * m + 0, gx * scal);
    atomicAdd(dx + 3 * m + 1, gy * scal);
    atomicAdd(dx + 3 * m + 2, gz * scal);
  }
  float w[8];
  corner_weights(c, w);
#pragma unroll
  for (int k = 0; k < 8; ++k) scatter_row<F>(dtable + level_off, c.row[k], gr, w[k]);
}

// Coarse levels: a level whose dense vertex lattice (res+1)^3 x F floats fits in shared memory is accumulated there
// first.  Millions of samples fall on a few thousand table rows at these levels, and same-row float reductions
// serialise in L2 (measured: 25 M reductions on the 4913 rows of a res-16 level take 7 ms, on a res-1024 level
// 0.16 ms).  Each CTA privatises the lattice, consumes a contiguous chunk of samples (walked ray-fastest so that the
// lanes of a warp sit in different cells) and flushes each touched vertex with ONE reduction to its hashed row.
constexpr int kDenseThreads = 512;
constexpr int kDenseMaxBytes = 200 * 1024;

template <int F>
__global__ void __launch_bounds__(kDenseThreads) hash_bwd_dense_kernel(
    const __grid_constant__ GridDev g, int level, const float* __restrict__ x, const float* __restrict__ std,
    const float* __restrict__ dy, float* __restrict__ dtable, int64_t M, int group, int64_t chunk) {
  extern __shared__ __align__(16) float acc[];
  const float scal = g.scalings[level];
  const int R1 = static_cast<int>(scal) + 1;
  const int V = R1 * R1 * R1;
  for (int i = threadIdx.x; i < V * F; i += kDenseThreads) acc[i] = 0.0f;
  __syncthreads();
  const uint32_t mask = (1u << g.log2_size) - 1u;
  float* level_base = dtable + (static_cast<size_t>(level) << g.log2_size) * F;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * chunk;
  const int count = static_cast<int>(min(chunk, M - m0));
  const int nr = (count % group == 0) ? count / group : count;  // rays in this chunk (whole rays when possible)
  const int ns = (count % group == 0) ? group : 1;
  for (int idx = threadIdx.x; idx < count; idx += kDenseThreads) {
    const int s = idx / nr, r = idx - s * nr;
    const int64_t m = m0 + static_cast<int64_t>(r) * ns + s;
    const float px = __ldg(x + 3 * m), py = __ldg(x + 3 * m + 1), pz = __ldg(x + 3 * m + 2);
    using Vt = typename Feat<F>::type;
    const Vt gv = __ldg(reinterpret_cast<const Vt*>(dy) + m * g.num_levels + level);
    float gr[F];
    if constexpr (F == 1) {
      gr[0] = gv;
    } else if constexpr (F == 2) {
      gr[0] = gv.x;
      gr[1] = gv.y;
    } else {
      gr[0] = gv.x;
      gr[1] = gv.y;
      gr[2] = gv.z;
      gr[3] = gv.w;
    }
    if (std != nullptr) {
      const float lw = level_weight(scal, __ldg(std + m));
#pragma unroll
      for (int j = 0; j < F; ++j) gr[j] *= lw;
    }
    const float sx = mul(px, scal), sy = mul(py, scal), sz = mul(pz, scal);
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    const int xf = static_cast<int>(fx), yf = static_cast<int>(fy), zf = static_cast<int>(fz);
    const int xc = static_cast<int>(ceilf(sx)), yc = static_cast<int>(ceilf(sy)), zc = static_cast<int>(ceilf(sz));
    Cell c;
    c.ox = sub(sx, fx);
    c.oy = sub(sy, fy);
    c.oz = sub(sz, fz);
    float w[8];
    corner_weights(c, w);
    const bool inside = xf >= 0 && yf >= 0 && zf >= 0 && xc < R1 && yc < R1 && zc < R1;
    const int cx[8] = {xc, xc, xf, xf, xc, xc, xf, xf};
    const int cy[8] = {yc, yf, yf, yc, yc, yf, yf, yc};
    const int cz[8] = {zc, zc, zc, zc, zf, zf, zf, zf};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (inside) {
        float* p = acc + ((cz[k] * R1 + cy[k]) * R1 + cx[k]) * F;
#pragma unroll
        for (int j = 0; j < F; ++j) atomicAdd(p + j, gr[j] * w[k]);
      } else {  // points outside the unit cube: straight to the hashed row
        const uint32_t row = (static_cast<uint32_t>(cx[k]) ^ (static_cast<uint32_t>(cy[k]) * kPrimeY) ^
                              (static_cast<uint32_t>(cz[k]) * kPrimeZ)) & mask;
        scatter_row<F>(level_base, row, gr, w[k]);
      }
    }
  }
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += kDenseThreads) {
    float val[F];
    bool any = false;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      val[j] = acc[v * F + j];
      any |= (val[j] != 0.0f);
    }
    if (!any) continue;
    const int ix = v % R1, iy = (v / R1) % R1, iz = v / (R1 * R1);
    const uint32_t row = (static_cast<uint32_t>(ix) ^ (static_cast<uint32_t>(iy) * kPrimeY) ^
                          (static_cast<uint32_t>(iz) * kPrimeZ)) & mask;
    scatter_row<F>(level_base, row, val, 1.0f);
  }
}

__global__ void __launch_bounds__(256) frustum_gaussians_kernel(const float* __restrict__ origins,
                                                                const float* __restrict__ directions,
                                                                const float* __restrict__ pixel_area,
                                                                nrb_intervals_t iv, float scale,
                                                                float* __restrict__ x, float* __restrict__ std,
                                                                int64_t total) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int S = iv.num_samples;
  const int64_t n = gid / S;
  const int s = static_cast<int>(gid - n * S);
  const Gaussian g = sample_gaussian(origins[3 * n], origins[3 * n + 1], origins[3 * n + 2], directions[3 * n],
                                     directions[3 * n + 1], directions[3 * n + 2], pixel_area[n],
                                     iv.starts[n * iv.row_stride + s], iv.ends[n * iv.row_stride + s], scale);
  x[3 * gid + 0] = g.x;
  x[3 * gid + 1] = g.y;
  x[3 * gid + 2] = g.z;
  std[gid] = g.std;
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_hash_fwd(const nrb_grid_t* grid, const float* x, const float* std, float* out, int64_t M,
                            nrb_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  NRB_REQUIRE(x && out && M >= 0, NRB_ERR_BAD_ARG, "nrb_hash_fwd: null pointer or negative M");
  NRB_REQUIRE(aligned16(out), NRB_ERR_ALIGNMENT, "nrb_hash_fwd: out must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  const int64_t total = M * grid->num_levels;
  const GridDev g = to_dev(grid);
  auto s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = blocks_for(total, 256);
  switch (grid->features_per_level) {
    case 1: hash_fwd_kernel<1><<<blocks, 256, 0, s>>>(g, x, std, out, total); break;
    case 2: hash_fwd_kernel<2><<<blocks, 256, 0, s>>>(g, x, std, out, total); break;
    default: hash_fwd_kernel<4><<<blocks, 256, 0, s>>>(g, x, std, out, total); break;
  }
  return finish_launch("nrb_hash_fwd");
}

extern "C" int nrb_hash_indices(const nrb_grid_t* grid, const float* x, int64_t* idx, int64_t M,
                                nrb_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  NRB_REQUIRE(x && idx && M >= 0, NRB_ERR_BAD_ARG, "nrb_hash_indices: null pointer or negative M");
  if (M == 0) return NRB_OK;
  const int64_t total = M * grid->num_levels;
  hash_indices_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(to_dev(grid), x, idx,
                                                                                             total);
  return finish_launch("nrb_hash_indices");
}

template <int F>
static int launch_hash_bwd(const nrb_grid_t* grid, const GridDev& g, const float* x, const float* std, const float* dy,
                           float* dtable, float* dx, int64_t M, int group, cudaStream_t s) {
  // dense-privatised coarse levels (not used when the input gradient is requested: that path needs the features)
  int level0 = 0;
  if (dx == nullptr) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(hash_bwd_dense_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kDenseMaxBytes);
      NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_hash_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_set = true;
    }
    while (level0 < grid->num_levels) {
      const int64_t r1 = static_cast<int64_t>(grid->scalings[level0]) + 1;
      const int64_t bytes = r1 * r1 * r1 * F * 4;
      if (bytes > kDenseMaxBytes || M < 65536) break;
      // enough CTAs to fill the machine at this level's shared-memory footprint, whole rays per chunk
      const int per_sm = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(4, (220 * 1024) / bytes)));
      int64_t ctas = static_cast<int64_t>(sm_count()) * per_sm;
      int64_t chunk = (M + ctas - 1) / ctas;
      chunk = (chunk + group - 1) / group * group;
      ctas = (M + chunk - 1) / chunk;
      hash_bwd_dense_kernel<F><<<static_cast<unsigned>(ctas), kDenseThreads, static_cast<size_t>(bytes), s>>>(
          g, level0, x, std, dy, dtable, M, group, chunk);
      count_launch();
      ++level0;
    }
  }
  if (level0 < grid->num_levels) {
    const int64_t total = M * (grid->num_levels - level0);
    const unsigned blocks = blocks_for(total, 256);
    if (dx != nullptr) {
      hash_bwd_kernel<F, true><<<blocks, 256, 0, s>>>(g, x, std, dy, dtable, dx, total, level0);
    } else {
      hash_bwd_kernel<F, false><<<blocks, 256, 0, s>>>(g, x, std, dy, dtable, dx, total, level0);
    }
  }
  return finish_launch("nrb_hash_bwd");
}

extern "C" int nrb_hash_bwd(const nrb_grid_t* grid, const float* x, const float* std, const float* dy, float* dtable,
                            float* dx, int64_t M, int32_t samples_per_ray, nrb_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  NRB_REQUIRE(x && dy && dtable && M >= 0, NRB_ERR_BAD_ARG, "nrb_hash_bwd: null pointer or negative M");
  NRB_REQUIRE(aligned16(dy) && aligned16(dtable), NRB_ERR_ALIGNMENT, "nrb_hash_bwd: dy/dtable must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  const GridDev g = to_dev(grid);
  auto s = static_cast<cudaStream_t>(stream);
  const int group = (samples_per_ray > 0 && M % samples_per_ray == 0) ? samples_per_ray : 1;
  if (dx != nullptr) {
    cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * 3 * M, s);
    NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_hash_bwd: memset failed: %s", cudaGetErrorString(e));
  }
  switch (grid->features_per_level) {
    case 1: return launch_hash_bwd<1>(grid, g, x, std, dy, dtable, dx, M, group, s);
    case 2: return launch_hash_bwd<2>(grid, g, x, std, dy, dtable, dx, M, group, s);
    default: return launch_hash_bwd<4>(grid, g, x, std, dy, dtable, dx, M, group, s);
  }
}

extern "C" int nrb_frustum_gaussians(const nrb_rays_t* rays, const nrb_intervals_t* iv, float scale, float* x,
                                     float* std, nrb_stream_t stream) {
  if (int rc = check_rays(rays)) return rc;
  if (int rc = check_intervals("nrb_frustum_gaussians", iv)) return rc;
  NRB_REQUIRE(x && std, NRB_ERR_BAD_ARG, "nrb_frustum_gaussians: null output pointer");
  NRB_REQUIRE(scale > 0.f, NRB_ERR_BAD_ARG, "nrb_frustum_gaussians: scale must be positive");
  if (rays->num_rays == 0) return NRB_OK;
  const int64_t total = rays->num_rays * iv->num_samples;
  frustum_gaussians_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rays->origins, rays->directions, rays->pixel_area, *iv, scale, x, std, total);
  return finish_launch("nrb_frustum_gaussians");
}
