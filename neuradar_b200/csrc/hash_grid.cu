// Multiresolution hash grid: forward gather, index dump, backward scatter, and the sample-gaussian producer.
// Semantics: HashEncoding.pytorch_fwd / hash_fn (nerfstudio/field_components/encodings.py:406-466) and
// NeuRADHashEncoding._rescale_grid_features (nerfstudio/field_components/neurad_encoding.py:309-316).
//
// Work decomposition: one thread per (point, level) pair with the level index fastest.  The 2^19..2^22-row tables
// are hit at random, so there is nothing to coalesce on the gather side; what can be coalesced is the [M, L*F]
// feature matrix, and with the level fastest a warp writes (reads, in the backward pass) one contiguous
// 32*F*4-byte span of it.  Each thread issues its 8 independent F-wide vector gathers before touching any of them.
#include <algorithm>
#include <cstdlib>

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

// Backward scatter.  Same-row float reductions serialise in L2: 25 M reductions take 0.16 ms when they spread over
// the 2^19 rows of a res-1024 level and 7 ms when they pile onto the 4913 rows of a res-16 level (measured, B200).
// Coarse levels are therefore accumulated into `copies[l]` replicas of the level's dense vertex lattice
// ((res+1)^3 x F floats, a few MB in total, L2 resident); warps are dealt round-robin over the replicas, which
// divides the per-address contention by the replica count, and a small fold kernel adds the replicas into the
// hashed rows.  Levels that already fill the table but whose samples are still spatially coherent (res 84..256 at
// T = 2^19) get a few replicas of the hashed level table itself.  Replica counts were swept on B200 (tools_bwd_sweep.py,
// config 2): 8.5 ms without replicas, 3.9 ms with lattice replicas only, 2.7-2.8 ms with the defaults below.  Shared-memory privatisation was measured and rejected: fp32 shared atomics are CAS loops
// (ATOMS.CAST.SPIN) on sm_100 and cost ~1 ms per level.
struct BwdPlan {
  float* scratch;                   // replicated lattices, zeroed by the launcher
  int64_t offset[NRB_MAX_LEVELS];   // float offset of level l's first replica
  int copies[NRB_MAX_LEVELS];       // 0/1: scatter straight into the table
  int r1[NRB_MAX_LEVELS];           // res + 1
};

template <int F, bool kNeedDx>
__global__ void __launch_bounds__(256) hash_bwd_kernel(const __grid_constant__ GridDev g,
                                                       const __grid_constant__ BwdPlan plan,
                                                       const float* __restrict__ x, const float* __restrict__ std,
                                                       const float* __restrict__ dy, float* __restrict__ dtable,
                                                       float* __restrict__ dx, int64_t total) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int64_t m = gid / g.num_levels;
  const int l = static_cast<int>(gid - m * g.num_levels);
  const float scal = g.scalings[l];
  const float px = __ldg(x + 3 * m), py = __ldg(x + 3 * m + 1), pz = __ldg(x + 3 * m + 2);
  const Cell c = locate_cell(px, py, pz, scal, (1u << g.log2_size) - 1u);
  using V = typename Feat<F>::type;
  const V gv = __ldg(reinterpret_cast<const V*>(dy) + gid);
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
  const int copies = plan.copies[l];
  if (copies > 1 && plan.r1[l] == 0) {  // replicas of the hashed level table
    const unsigned warp_global = static_cast<unsigned>(gid >> 5);
    float* rep = plan.scratch + plan.offset[l] +
                 (static_cast<size_t>(warp_global % static_cast<unsigned>(copies)) << g.log2_size) * F;
#pragma unroll
    for (int k = 0; k < 8; ++k) scatter_row<F>(rep, c.row[k], gr, w[k]);
    return;
  }
  if (copies > 1) {
    const int R1 = plan.r1[l];
    const float sx = mul(px, scal), sy = mul(py, scal), sz = mul(pz, scal);
    const int xf = static_cast<int>(floorf(sx)), yf = static_cast<int>(floorf(sy)), zf = static_cast<int>(floorf(sz));
    const int xc = static_cast<int>(ceilf(sx)), yc = static_cast<int>(ceilf(sy)), zc = static_cast<int>(ceilf(sz));
    if (xf >= 0 && yf >= 0 && zf >= 0 && xc < R1 && yc < R1 && zc < R1) {
      const unsigned warp_global = static_cast<unsigned>(gid >> 5);
      float* rep = plan.scratch + plan.offset[l] +
                   static_cast<size_t>(warp_global % static_cast<unsigned>(copies)) * (static_cast<size_t>(R1) * R1 * R1 * F);
      const int cx[8] = {xc, xc, xf, xf, xc, xc, xf, xf};
      const int cy[8] = {yc, yf, yf, yc, yc, yf, yf, yc};
      const int cz[8] = {zc, zc, zc, zc, zf, zf, zf, zf};
#pragma unroll
      for (int k = 0; k < 8; ++k) scatter_row<F>(rep, static_cast<uint32_t>((cz[k] * R1 + cy[k]) * R1 + cx[k]), gr, w[k]);
      return;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) scatter_row<F>(dtable + level_off, c.row[k], gr, w[k]);
}

// Adds the replicas of every replicated level into the table: one thread per lattice vertex.
template <int F>
__global__ void __launch_bounds__(256) hash_bwd_fold_kernel(const __grid_constant__ GridDev g,
                                                            const __grid_constant__ BwdPlan plan,
                                                            float* __restrict__ dtable, int64_t total_vertices) {
  int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= total_vertices) return;
  int l = 0;
  for (; l < g.num_levels; ++l) {
    if (plan.copies[l] <= 1) continue;
    const int64_t nvl = plan.r1[l] > 0 ? static_cast<int64_t>(plan.r1[l]) * plan.r1[l] * plan.r1[l]
                                       : (int64_t{1} << g.log2_size);
    if (v < nvl) break;
    v -= nvl;
  }
  if (l >= g.num_levels) return;
  const int R1 = plan.r1[l];
  const int64_t nv = R1 > 0 ? static_cast<int64_t>(R1) * R1 * R1 : (int64_t{1} << g.log2_size);
  const float* rep = plan.scratch + plan.offset[l] + v * F;
  float sum[F];
#pragma unroll
  for (int j = 0; j < F; ++j) sum[j] = 0.0f;
  for (int cpy = 0; cpy < plan.copies[l]; ++cpy) {
    float t[F];
    load_row<F>(rep + static_cast<size_t>(cpy) * nv * F, 0, t);
#pragma unroll
    for (int j = 0; j < F; ++j) sum[j] += t[j];
  }
  bool any = false;
#pragma unroll
  for (int j = 0; j < F; ++j) any |= (sum[j] != 0.0f);
  if (!any) return;
  uint32_t row;
  if (R1 > 0) {
    const int ix = static_cast<int>(v % R1), iy = static_cast<int>((v / R1) % R1), iz = static_cast<int>(v / (R1 * R1));
    row = (static_cast<uint32_t>(ix) ^ (static_cast<uint32_t>(iy) * kPrimeY) ^ (static_cast<uint32_t>(iz) * kPrimeZ)) &
          ((1u << g.log2_size) - 1u);
  } else {
    row = static_cast<uint32_t>(v);
  }
  scatter_row<F>(dtable + (static_cast<size_t>(l) << g.log2_size) * F, row, sum, 1.0f);
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

// Replica plan: as many copies of a coarse level's lattice as it takes to bring its reductions per address down to
// those of a level that fills the whole table, within the caller's workspace.
static double env_or(const char* name, double dflt) {
  const char* v = std::getenv(name);
  return v != nullptr ? std::atof(v) : dflt;
}

static int64_t plan_hash_bwd(const nrb_grid_t* grid, int64_t M, BwdPlan* plan) {
  // tunables (defaults chosen on B200 with tools/sweep): replicas ~ scale * table_rows / vertices,
  // at most cap_mb per level; levels that already fill the table get `hashed` replicas of the hashed table itself
  static const double scale = env_or("NRB_BWD_SCALE", 4.0), cap_mb = env_or("NRB_BWD_CAP_MB", 32.0);
  static const int hashed = static_cast<int>(env_or("NRB_BWD_HASHED_COPIES", 4.0));
  static const int hashed_levels = static_cast<int>(env_or("NRB_BWD_HASHED_LEVELS", 6.0));
  const int F = grid->features_per_level;
  const int64_t T = int64_t{1} << grid->log2_hashmap_size;
  const double table_rows = static_cast<double>(T);
  int64_t floats = 0;
  int hashed_used = 0;
  for (int l = 0; l < NRB_MAX_LEVELS; ++l) {
    plan->copies[l] = 0;
    plan->offset[l] = 0;
    plan->r1[l] = 0;
    if (l >= grid->num_levels || M < (int64_t{1} << 16)) continue;
    const int64_t r1 = static_cast<int64_t>(grid->scalings[l]) + 1;
    const double verts = static_cast<double>(r1) * r1 * r1;
    int copies;
    int64_t per_copy;
    if (verts * 1.5 > table_rows || r1 > 1024) {  // the level already spreads over (most of) the table
      if (hashed <= 1 || hashed_used >= hashed_levels) continue;
      ++hashed_used;
      copies = hashed;
      per_copy = T * F;
      plan->r1[l] = 0;  // replicas of the hashed level table
    } else {
      copies = static_cast<int>(scale * table_rows / verts + 0.5);
      copies = std::min(copies, 1024);
      per_copy = r1 * r1 * r1 * F;
      plan->r1[l] = static_cast<int>(r1);
    }
    while (copies > 1 && static_cast<double>(copies) * per_copy * 4 > cap_mb * 1024 * 1024) --copies;
    if (copies <= 1) {
      plan->r1[l] = 0;
      continue;
    }
    plan->copies[l] = copies;
    plan->offset[l] = floats;
    floats += static_cast<int64_t>(copies) * per_copy;
    floats = (floats + 3) & ~int64_t{3};
  }
  return floats * 4;
}

extern "C" int64_t nrb_hash_bwd_workspace_bytes(const nrb_grid_t* grid, int64_t M) {
  if (grid == nullptr || check_grid(grid) != NRB_OK) return -1;
  BwdPlan plan;
  return plan_hash_bwd(grid, M, &plan);
}

template <int F>
static int launch_hash_bwd(const nrb_grid_t* grid, const GridDev& g, const float* x, const float* std, const float* dy,
                           float* dtable, float* dx, int64_t M, void* workspace, int64_t workspace_bytes,
                           cudaStream_t s) {
  BwdPlan plan;
  const int64_t need = plan_hash_bwd(grid, M, &plan);
  int64_t vertices = 0;
  if (need > 0 && workspace != nullptr && workspace_bytes >= need && aligned16(workspace)) {
    plan.scratch = static_cast<float*>(workspace);
    cudaError_t e = cudaMemsetAsync(workspace, 0, static_cast<size_t>(need), s);
    NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_hash_bwd: memset failed: %s", cudaGetErrorString(e));
    for (int l = 0; l < grid->num_levels; ++l)
      if (plan.copies[l] > 1)
        vertices += plan.r1[l] > 0 ? static_cast<int64_t>(plan.r1[l]) * plan.r1[l] * plan.r1[l]
                                   : (int64_t{1} << grid->log2_hashmap_size);
  } else {  // no (or too small a) workspace: every level scatters straight into the table
    plan.scratch = nullptr;
    for (int l = 0; l < NRB_MAX_LEVELS; ++l) plan.copies[l] = 0;
  }
  const int64_t total = M * grid->num_levels;
  const unsigned blocks = blocks_for(total, 256);
  if (dx != nullptr) {
    hash_bwd_kernel<F, true><<<blocks, 256, 0, s>>>(g, plan, x, std, dy, dtable, dx, total);
  } else {
    hash_bwd_kernel<F, false><<<blocks, 256, 0, s>>>(g, plan, x, std, dy, dtable, dx, total);
  }
  if (vertices > 0) {
    count_launch();
    hash_bwd_fold_kernel<F><<<blocks_for(vertices, 256), 256, 0, s>>>(g, plan, dtable, vertices);
  }
  return finish_launch("nrb_hash_bwd");
}

extern "C" int nrb_hash_bwd(const nrb_grid_t* grid, const float* x, const float* std, const float* dy, float* dtable,
                            float* dx, int64_t M, void* workspace, int64_t workspace_bytes, nrb_stream_t stream) {
  if (int rc = check_grid(grid)) return rc;
  NRB_REQUIRE(x && dy && dtable && M >= 0, NRB_ERR_BAD_ARG, "nrb_hash_bwd: null pointer or negative M");
  NRB_REQUIRE(aligned16(dy) && aligned16(dtable), NRB_ERR_ALIGNMENT, "nrb_hash_bwd: dy/dtable must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  const GridDev g = to_dev(grid);
  auto s = static_cast<cudaStream_t>(stream);
  if (dx != nullptr) {
    cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * 3 * M, s);
    NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_hash_bwd: memset failed: %s", cudaGetErrorString(e));
  }
  switch (grid->features_per_level) {
    case 1: return launch_hash_bwd<1>(grid, g, x, std, dy, dtable, dx, M, workspace, workspace_bytes, s);
    case 2: return launch_hash_bwd<2>(grid, g, x, std, dy, dtable, dx, M, workspace, workspace_bytes, s);
    default: return launch_hash_bwd<4>(grid, g, x, std, dy, dtable, dx, M, workspace, workspace_bytes, s);
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
