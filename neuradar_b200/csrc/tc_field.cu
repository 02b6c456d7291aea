// Fused NeuRAD field MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Semantics: NeuRADField.forward after the hash grid (nerfstudio/fields/neurad_field.py:132-152):
//   geo   = mlp_geo(x)                      32 -> 32 (ReLU) -> 33         (field_components/mlp.py:159-178)
//   sdf, emb = split(geo, [1, 32])
//   feat  = emb + mlp_feature([emb, sh])    48 -> 32 (ReLU) -> 32 (ReLU) -> 32
//   alpha = sigmoid(-sdf * (|beta| + 1e-4))                               (model_components/utils.py:30-41)
// One CTA = 128 threads = one 128-sample tile at a time (persistent over tiles); thread t owns row t: it stages its
// row of the next layer's A operand (hi/lo tf32 halves) in shared memory, one elected thread issues the layer's
// tcgen05.mma chain, and every thread reads its accumulator row back with tcgen05.ld for bias + ReLU.  Weights are
// staged once per CTA.  Activations needed by the backward pass are written feature-major ([32][ld], coalesced
// 128-byte warp stores) together with one bit mask per ReLU layer.
#include <algorithm>
#include <cstdlib>

#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace nrb {

using namespace tc;

struct FieldParams {
  const float* w[5];  // mlp_geo.layers.{0,1}.weight, mlp_feature.layers.{0,1,2}.weight   ([out, in] row-major)
  const float* b[5];
  const float* beta;  // sdf_to_density.beta [1]
  float beta_min;
};

struct FieldSaved {  // activations kept for the backward pass, feature-major with leading dimension ld
  float* h1;         // [32][ld] post-ReLU hidden of mlp_geo
  float* emb;        // [32][ld] geometry embedding
  float* g1;         // [32][ld] post-ReLU hidden 1 of mlp_feature
  float* g2;         // [32][ld] post-ReLU hidden 2 of mlp_feature
  uint32_t* masks;   // [3][ld] bit j = (unit j > 0) for h1, g1, g2
  int64_t ld;        // multiple of 128, >= M
};

constexpr int kTmemCols = 256;               // accumulator (48) + the A operand as tf32 hi / lo (2 x 48); 2 CTAs per SM
constexpr int kFwdAHi = 64, kFwdALo = 128;  // TMEM columns of the A operand
// layer geometry: K (input width), N (output width padded to a multiple of 16), rows of the weight matrix
__device__ constexpr int kK[5] = {32, 32, 48, 32, 32};
__device__ constexpr int kN[5] = {32, 48, 32, 32, 32};
__device__ constexpr int kOut[5] = {32, 33, 32, 32, 32};

// byte offsets into dynamic shared memory
__host__ __device__ constexpr int field_w_floats(int l) { return l == 0 ? 32 * 32 : l == 1 ? 48 * 32 : l == 2 ? 32 * 48 : 32 * 32; }
__host__ __device__ constexpr int field_w_hi(int l) {
  int o = 0;
  for (int i = 0; i < l; ++i) o += 2 * field_w_floats(i) * 4;
  return o;
}
__host__ __device__ constexpr int field_w_lo(int l) { return field_w_hi(l) + field_w_floats(l) * 4; }

struct FieldSmem {
  static constexpr int bias = field_w_hi(5);        // 5 x 48 floats
  static constexpr int a_hi = bias + 5 * 48 * 4;    // 128 x 48 floats
  static constexpr int a_lo = a_hi + kRows * 48 * 4;
  static constexpr int mbar = a_lo + kRows * 48 * 4;
  static constexpr int tmem = mbar + 8;
  static constexpr int total = tmem + 8;
};

__device__ __forceinline__ void load_row32(const float* __restrict__ p, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + c);
    v[4 * c] = t.x;
    v[4 * c + 1] = t.y;
    v[4 * c + 2] = t.z;
    v[4 * c + 3] = t.w;
  }
}

__device__ __forceinline__ void store_row32(float* __restrict__ p, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    reinterpret_cast<float4*>(p)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// coalesced feature-major store of one row: element j goes to p[j * ld + row]
__device__ __forceinline__ void store_col32(float* __restrict__ p, int64_t ld, int64_t row, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) p[j * ld + row] = v[j];
}

__device__ __forceinline__ uint32_t relu_bias_mask(float (&v)[32], const float* __restrict__ bias) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] = fmaxf(v[j] + bias[j], 0.0f);
    m |= (v[j] > 0.0f ? 1u : 0u) << j;
  }
  return m;
}

__global__ void __launch_bounds__(kRows, 2) field_mlp_fwd_kernel(const __grid_constant__ FieldParams prm,
                                                                 const __grid_constant__ FieldSaved sv,
                                                                 const float* __restrict__ x,   // [M,32]
                                                                 const float* __restrict__ sh,  // [N_rays,16]
                                                                 int samples_per_ray, int64_t M,
                                                                 float* __restrict__ feature,  // [M,32]
                                                                 float* __restrict__ sdf,      // [M]
                                                                 float* __restrict__ alpha,    // [M]
                                                                 int dbg) {
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5;
  float* s_bias = reinterpret_cast<float*>(smem + FieldSmem::bias);
  char* a_hi = smem + FieldSmem::a_hi;
  char* a_lo = smem + FieldSmem::a_lo;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + FieldSmem::mbar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FieldSmem::tmem);

  for (int l = 0; l < 5; ++l) {
    stage_weight_split(prm.w[l], kOut[l], kN[l], kK[l], smem + field_w_hi(l), smem + field_w_lo(l));
    for (int j = t; j < 48; j += kRows) s_bias[l * 48 + j] = (j < kOut[l] && prm.b[l] != nullptr) ? __ldg(prm.b[l] + j) : 0.0f;
  }
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
  if (t == 0) mbar_init(mbar, 1);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_hi_u = smem_u32(a_hi), a_lo_u = smem_u32(a_lo);
  const float beta = fabsf(__ldg(prm.beta)) + prm.beta_min;
  const bool train = sv.h1 != nullptr;
  const int uwarp = uniform_warp_idx();
  uint32_t phase = 0;

  // one layer: operands are staged; fence, issue, wait for the accumulator
  // The layer input (this thread's row, tf32 hi / lo) goes straight to tensor memory (tcgen05.st) and is read from
  // there by the ".ts" form of tcgen05.mma: no shared-memory staging (2 x 16-24 KB of stores per layer and tile) and
  // no shared-memory operand fetch for A.  dbg bit 4 selects the shared-memory A path (kept for comparison).
  const bool a_smem = (dbg & 16) != 0;
  auto run_layer = [&](int l) {
    if (a_smem) fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (dbg & 1) return;
    if (uwarp == 0) {
      if (elect_one()) {
        fence_after_sync();
        if (a_smem)
          issue_gemm(tmem_base, 128, kN[l], a_hi_u, a_lo_u, kK[l], smem_u32(smem + field_w_hi(l)),
                     smem_u32(smem + field_w_lo(l)), kK[l], kK[l], false);
        else if (!(dbg & 2))
          issue_gemm_ts(tmem_base, kN[l], tmem_base + kFwdAHi, tmem_base + kFwdALo, smem_u32(smem + field_w_hi(l)),
                        smem_u32(smem + field_w_lo(l)), kK[l], kK[l]);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
  };

  auto stage32 = [&](const float (&r)[32]) {
    if (a_smem)
      store_row_split<32>(a_hi, a_lo, t, r);
    else
      tmem_store_row_split<32>(tmem_base, warp, kFwdAHi, kFwdALo, 0, r);
  };
  const int64_t tiles = (M + kRows - 1) / kRows;
  const int lane = t & 31;
  // The operand tiles double as bounce buffers for the coalesced row I/O while no MMA is reading them: a warp's 32
  // rows occupy exactly bytes [4096 w, 4096 (w+1)) of a 32-column tile, so warps never touch each other's part.
  char* bounce_in = a_lo + warp * 4096;
  char* bounce_out = a_hi + warp * 4096;
  float v[32];
  float4 xpf[8];  // this warp's 32 input rows, chunk-major, loaded one tile ahead
  if (static_cast<int64_t>(blockIdx.x) < tiles)
    warp_load_rows_coalesced(x, static_cast<int64_t>(blockIdx.x) * kRows + warp * 32, M, lane, xpf);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row = tile * kRows + t;
    const bool ok = row < M;
    const int64_t rr = ok ? row : (M - 1);  // out-of-range rows compute on a valid row; only saved activations are stored
    // ---- mlp_geo layer 0: 32 -> 32, ReLU
    warp_bounce_to_rows(bounce_in, lane, xpf, v);
    stage32(v);
    const int64_t ntile = tile + gridDim.x;
    if (ntile < tiles) warp_load_rows_coalesced(x, ntile * kRows + warp * 32, M, lane, xpf);
    run_layer(0);
    tmem_load_row<32>(tmem_base, warp, 0, v);
    const uint32_t m_h1 = relu_bias_mask(v, s_bias);
    if (train) store_col32(sv.h1, sv.ld, row, v);
    stage32(v);
    // ---- mlp_geo layer 1: 32 -> 33 (sdf | embedding), no activation
    run_layer(1);
    float geo[48];
    tmem_load_row<48>(tmem_base, warp, 0, geo);
    const float sdf_v = geo[0] + s_bias[48 + 0];
    float emb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) emb[j] = geo[1 + j] + s_bias[48 + 1 + j];
    if (train) store_col32(sv.emb, sv.ld, row, emb);
    // ---- mlp_feature layer 0: [emb | sh] 48 -> 32, ReLU
    {
      float in48[48];
#pragma unroll
      for (int j = 0; j < 32; ++j) in48[j] = emb[j];
      const float4* shp = reinterpret_cast<const float4*>(sh + (rr / samples_per_ray) * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 q = __ldg(shp + c);
        in48[32 + 4 * c] = q.x;
        in48[32 + 4 * c + 1] = q.y;
        in48[32 + 4 * c + 2] = q.z;
        in48[32 + 4 * c + 3] = q.w;
      }
      if (a_smem)
        store_row_split<48>(a_hi, a_lo, t, in48);
      else
        tmem_store_row_split<48>(tmem_base, warp, kFwdAHi, kFwdALo, 0, in48);
    }
    run_layer(2);
    tmem_load_row<32>(tmem_base, warp, 0, v);
    const uint32_t m_g1 = relu_bias_mask(v, s_bias + 2 * 48);
    if (train) store_col32(sv.g1, sv.ld, row, v);
    stage32(v);
    // ---- mlp_feature layer 1: 32 -> 32, ReLU
    run_layer(3);
    tmem_load_row<32>(tmem_base, warp, 0, v);
    const uint32_t m_g2 = relu_bias_mask(v, s_bias + 3 * 48);
    if (train) {
      store_col32(sv.g2, sv.ld, row, v);
      sv.masks[row] = m_h1;
      sv.masks[sv.ld + row] = m_g1;
      sv.masks[2 * sv.ld + row] = m_g2;
    }
    stage32(v);
    // ---- mlp_feature layer 2: 32 -> 32, no activation; residual with the embedding
    run_layer(4);
    tmem_load_row<32>(tmem_base, warp, 0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = emb[j] + (v[j] + s_bias[4 * 48 + j]);
    warp_store_rows_coalesced(feature, tile * kRows + warp * 32, M, bounce_out, lane, v);  // last MMA is complete
    if (ok) {
      sdf[row] = sdf_v;
      alpha[row] = 1.0f / (1.0f + expf(sdf_v * beta));  // sigmoid(-sdf * beta)
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of the fused field MLP.  Per 128-sample tile and per layer (last to first) two GEMM chains are issued on
// the tensor cores:
//     dIn_l  = delta_l * W_l            M = 128, N = in_l, reduction over out_l; A = delta rows (K-major tile),
//                                        B = W_l^T staged once per CTA.  On the critical path: committed first.
//     dW_l (+)= delta_l^T * in_l        M = 64 (rows >= out_l ignored), N = in_l, reduction over the 128 samples;
//                                        A = delta^T, B = in_l^T, both [features x 128 samples] tiles.  delta^T comes
//                                        straight from registers through 4x4 shuffle transposes, in_l^T is a coalesced
//                                        copy of the feature-major activations saved by the forward kernel (prefetched
//                                        one layer ahead).  Accumulated in TMEM across ALL tiles of the CTA, flushed once.
// Bias gradients are column sums of delta_l (31-shuffle transpose-reduce per warp) kept in shared memory.
struct FieldBwdIn {
  const float* x;         // [M,32] hash features (row-major, as produced by the hash kernel)
  const float* h1;        // saved activations, feature-major [32][ld]
  const float* emb;
  const float* g1;
  const float* g2;
  const uint32_t* masks;  // [3][ld]
  int64_t ld;
  const float* sh;        // [N_rays,16]
  const float* sdf;       // [M]
  const float* alpha;     // [M]
  const float* dfeature;  // [M,32]
  const float* dsdf;      // [M] or null
  const float* dalpha;    // [M] or null
};

struct FieldBwdOut {
  float* dx;      // [M,32] or null
  float* dw[5];   // accumulated into
  float* db[5];
  float* dbeta;   // [1]: d loss / d(|beta| + beta_min), accumulated into
};

constexpr int kBwdTmemCols = 256;
__device__ constexpr int kDwCol[5] = {208, 176, 128, 96, 64};  // TMEM column of each layer's dW accumulator

struct FieldBwdSmem {
  // transposed weights (hi/lo) share the forward kernel's per-layer sizes: W_l^T is [in_l rows x out_l(pad) cols]
  static constexpr int d_hi = field_w_hi(5);            // delta, [128 x 48] K-major (A of the data-gradient GEMM)
  static constexpr int d_lo = d_hi + kRows * 48 * 4;
  static constexpr int dt_hi = d_lo + kRows * 48 * 4;   // delta^T, [48(+16 read-only slack) x 128]
  static constexpr int dt_lo = dt_hi + 48 * kRows * 4;
  static constexpr int at_hi = dt_lo + 48 * kRows * 4;  // input^T, [48 x 128]
  static constexpr int at_lo = at_hi + 48 * kRows * 4;
  static constexpr int dbacc = at_lo + 48 * kRows * 4;  // [4 warps][5 layers][48]
  static constexpr int bounce = dbacc + 4 * 5 * 48 * 4; // 4 warps x 4 KB for coalesced row I/O
  static constexpr int mbar = bounce + 4 * 4096;        // two mbarriers
  static constexpr int tmem = mbar + 16;
  static constexpr int total = tmem + 8;
};

// TMEM lane that holds row r of an M = 64 accumulator (cta_group::1): 16 rows per 32-lane quarter (measured with
// tools_tc_probe.py: lane (r / 16) * 32 + r % 16).
__device__ __forceinline__ int m64_row_of_lane(int lane128) {
  const int q = lane128 >> 5, i = lane128 & 31;
  return i < 16 ? q * 16 + i : -1;
}

// Feature-major activations -> transposed operand tile.  The [32 x 128] block of one tile is 1024 16-byte chunks
// (feature j, samples 4c..4c+3); thread t takes chunks e = t, t+128, ...: j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31,
// so that a quarter warp writes 128 contiguous bytes of shared memory.
__device__ __forceinline__ void prefetch_act_chunks(const float* __restrict__ fm, int64_t ld, int64_t row0, int t,
                                                    float4 (&pf)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int e = t + q * kRows;
    const int j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
    pf[q] = __ldg(reinterpret_cast<const float4*>(fm + j * ld + row0 + 4 * c));
  }
}

__device__ __forceinline__ void commit_act_chunks(char* at_hi, char* at_lo, int t, const float4 (&pf)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int e = t + q * kRows;
    const int j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
    const float4 a = pf[q];
    const float4 h = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
    const float4 l = make_float4(a.x - h.x, a.y - h.y, a.z - h.z, a.w - h.w);
    const uint32_t off = tile_offset(j, c, kRows);
    *reinterpret_cast<float4*>(at_hi + off) = h;
    *reinterpret_cast<float4*>(at_lo + off) = l;
  }
}

__global__ void __launch_bounds__(kRows, 1) field_mlp_bwd_kernel(const __grid_constant__ FieldParams prm,
                                                                 const __grid_constant__ FieldBwdIn in,
                                                                 const __grid_constant__ FieldBwdOut out,
                                                                 int samples_per_ray, int64_t M, int dbg) {
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  char* d_hi = smem + FieldBwdSmem::d_hi;
  char* d_lo = smem + FieldBwdSmem::d_lo;
  char* dt_hi = smem + FieldBwdSmem::dt_hi;
  char* dt_lo = smem + FieldBwdSmem::dt_lo;
  char* at_hi = smem + FieldBwdSmem::at_hi;
  char* at_lo = smem + FieldBwdSmem::at_lo;
  float* dbacc = reinterpret_cast<float*>(smem + FieldBwdSmem::dbacc);
  uint64_t* mbar_data = reinterpret_cast<uint64_t*>(smem + FieldBwdSmem::mbar);
  uint64_t* mbar_dw = mbar_data + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FieldBwdSmem::tmem);

  for (int l = 0; l < 5; ++l)
    stage_weight_transposed_split(prm.w[l], kOut[l], kN[l], kK[l], smem + field_w_hi(l), smem + field_w_lo(l));
  for (int i = t; i < 4 * 5 * 48; i += kRows) dbacc[i] = 0.0f;
  if (warp == 0) tmem_alloc<kBwdTmemCols>(tmem_slot);
  if (t == 0) {
    mbar_init(mbar_data, 1);
    mbar_init(mbar_dw, 1);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const float beta = fabsf(__ldg(prm.beta)) + prm.beta_min;
  uint32_t phase_data = 0, phase_dw = 0;
  bool first = true, dw_pending = false;
  const int uwarp = uniform_warp_idx();
  float dbeta_acc = 0.0f;
  float* my_db = dbacc + warp * 5 * 48;
  char* bounce = smem + FieldBwdSmem::bounce + warp * 4096;

  // Tiles are staged (delta rows in d_*, delta^T in dt_*, input^T in at_*).  Issue both chains; the data chain is
  // committed first so that the next layer's delta does not wait for the (4x longer) weight-gradient chain.
  auto issue_layer = [&](int l, int dcols, int acols, int kred) {
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (uwarp == 0) {
      if (elect_one()) {
        fence_after_sync();
        issue_gemm(tmem_base, 128, acols, smem_u32(d_hi), smem_u32(d_lo), dcols, smem_u32(smem + field_w_hi(l)),
                   smem_u32(smem + field_w_lo(l)), kN[l], kred, false);
        mma_commit(mbar_data);
        if (!(dbg & 1))
          issue_gemm(tmem_base + kDwCol[l], 64, acols, smem_u32(dt_hi), smem_u32(dt_lo), kRows, smem_u32(at_hi),
                     smem_u32(at_lo), kRows, kRows, !first);
        mma_commit(mbar_dw);
      }
      __syncwarp();
    }
    dw_pending = true;
  };
  auto wait_data = [&]() {
    mbar_wait(mbar_data, phase_data);
    phase_data ^= 1;
    fence_after_sync();
  };
  auto wait_dw = [&]() {  // the previous weight-gradient chain still reads dt_* / at_*
    if (dw_pending) {
      mbar_wait(mbar_dw, phase_dw);
      phase_dw ^= 1;
      dw_pending = false;
    }
  };

  const int64_t tiles = (M + kRows - 1) / kRows;
  float4 pf[8];
  if (static_cast<int64_t>(blockIdx.x) < tiles) prefetch_act_chunks(in.g2, in.ld, static_cast<int64_t>(blockIdx.x) * kRows, t, pf);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kRows;
    const int64_t row = row0 + t;
    const bool ok = row < M;
    const int64_t rr = ok ? row : (M - 1);
    const uint32_t m_h1 = __ldg(in.masks + row), m_g1 = __ldg(in.masks + in.ld + row), m_g2 = __ldg(in.masks + 2 * in.ld + row);
    float delta[32], tmp[32];
    // ---- layer 4 (mlp_feature.layers.2): delta = d feature, input g2
    {
      float4 cpf[8];
      warp_load_rows_coalesced(in.dfeature, row0 + warp * 32, M, lane, cpf);
      warp_bounce_to_rows(bounce, lane, cpf, delta);
    }
    if (!ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) delta[j] = 0.0f;
    }
    float demb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) demb[j] = delta[j];  // residual branch
    wait_dw();
    store_row_split<32>(d_hi, d_lo, t, delta);
    if (!(dbg & 2)) store_rows_transposed_split(dt_hi, dt_lo, warp, lane, delta);
    if (!(dbg & 8)) commit_act_chunks(at_hi, at_lo, t, pf);
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    if (!(dbg & 4)) my_db[4 * 48 + lane] += warp_column_sums(tmp, lane);
    issue_layer(4, 32, 32, 32);
    prefetch_act_chunks(in.g1, in.ld, row0, t, pf);
    wait_data();
    tmem_load_row<32>(tmem_base, warp, 0, delta);
#pragma unroll
    for (int j = 0; j < 32; ++j) delta[j] = ((m_g2 >> j) & 1u) ? delta[j] : 0.0f;
    // ---- layer 3 (mlp_feature.layers.1): input g1
    wait_dw();
    store_row_split<32>(d_hi, d_lo, t, delta);
    if (!(dbg & 2)) store_rows_transposed_split(dt_hi, dt_lo, warp, lane, delta);
    if (!(dbg & 8)) commit_act_chunks(at_hi, at_lo, t, pf);
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    if (!(dbg & 4)) my_db[3 * 48 + lane] += warp_column_sums(tmp, lane);
    issue_layer(3, 32, 32, 32);
    prefetch_act_chunks(in.emb, in.ld, row0, t, pf);
    wait_data();
    tmem_load_row<32>(tmem_base, warp, 0, delta);
#pragma unroll
    for (int j = 0; j < 32; ++j) delta[j] = ((m_g1 >> j) & 1u) ? delta[j] : 0.0f;
    // ---- layer 2 (mlp_feature.layers.0): input [emb | sh]
    wait_dw();
    store_row_split<32>(d_hi, d_lo, t, delta);
    if (!(dbg & 2)) store_rows_transposed_split(dt_hi, dt_lo, warp, lane, delta);
    if (!(dbg & 8)) commit_act_chunks(at_hi, at_lo, t, pf);
    if (!(dbg & 8)) {  // rows 32..47 of the input^T tile: the ray's SH basis, 16 x 32 chunks, 4 per thread
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = t + q * kRows;
        const int k = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = min(row0 + 4 * c + i, M - 1);
          a[i] = __ldg(in.sh + (r / samples_per_ray) * 16 + k);
        }
        const float4 h = make_float4(tf32_hi(a[0]), tf32_hi(a[1]), tf32_hi(a[2]), tf32_hi(a[3]));
        const float4 l = make_float4(a[0] - h.x, a[1] - h.y, a[2] - h.z, a[3] - h.w);
        const uint32_t off = tile_offset(32 + k, c, kRows);
        *reinterpret_cast<float4*>(at_hi + off) = h;
        *reinterpret_cast<float4*>(at_lo + off) = l;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    if (!(dbg & 4)) my_db[2 * 48 + lane] += warp_column_sums(tmp, lane);
    issue_layer(2, 32, 48, 32);
    prefetch_act_chunks(in.h1, in.ld, row0, t, pf);
    wait_data();
    {
      float din[48];
      tmem_load_row<48>(tmem_base, warp, 0, din);
#pragma unroll
      for (int j = 0; j < 32; ++j) demb[j] += din[j];  // the SH part of the input carries no gradient
    }
    // ---- layer 1 (mlp_geo.layers.1): delta = [d sdf | d emb], padded to 48 columns; input h1
    float dsdf_v = 0.0f;
    if (ok) {
      const float a = __ldg(in.alpha + row), sd = __ldg(in.sdf + row);
      const float da = in.dalpha != nullptr ? __ldg(in.dalpha + row) : 0.0f;
      const float s = a * (1.0f - a);
      dsdf_v = (in.dsdf != nullptr ? __ldg(in.dsdf + row) : 0.0f) - da * beta * s;
      dbeta_acc -= da * sd * s;
    }
    wait_dw();
    {
      float d48[48];
      d48[0] = dsdf_v;
#pragma unroll
      for (int j = 0; j < 32; ++j) d48[1 + j] = demb[j];
#pragma unroll
      for (int j = 33; j < 48; ++j) d48[j] = 0.0f;
      store_row_split<48>(d_hi, d_lo, t, d48);
      // delta^T rows 0..31 = [dsdf, demb[0..30]], rows 32..47 = [demb[31], 0, ...]
      float lo32[32], hi16[32];
      lo32[0] = dsdf_v;
#pragma unroll
      for (int j = 1; j < 32; ++j) lo32[j] = demb[j - 1];
      hi16[0] = demb[31];
#pragma unroll
      for (int j = 1; j < 32; ++j) hi16[j] = 0.0f;
      if (!(dbg & 2)) store_rows_transposed_split(dt_hi, dt_lo, warp, lane, lo32);
      if (!(dbg & 2)) store_rows_transposed_split<4>(dt_hi, dt_lo, warp, lane, hi16, 32);  // rows 32..47 only
    }
    if (!(dbg & 8)) commit_act_chunks(at_hi, at_lo, t, pf);
    if (!(dbg & 4)) {
      const float s0 = warp_sum(dsdf_v);
      if (lane == 0) my_db[1 * 48 + 0] += s0;
#pragma unroll
      for (int j = 0; j < 32; ++j) tmp[j] = demb[j];
      my_db[1 * 48 + 1 + lane] += warp_column_sums(tmp, lane);
    }
    issue_layer(1, 48, 32, 40);
    float xrow[32];
    {
      float4 cpf[8];
      warp_load_rows_coalesced(in.x, row0 + warp * 32, M, lane, cpf);
      warp_bounce_to_rows(bounce, lane, cpf, xrow);
    }
    wait_data();
    tmem_load_row<32>(tmem_base, warp, 0, delta);
#pragma unroll
    for (int j = 0; j < 32; ++j) delta[j] = ((m_h1 >> j) & 1u) ? delta[j] : 0.0f;
    // ---- layer 0 (mlp_geo.layers.0): input x (row-major in memory: transposed through registers)
    wait_dw();
    store_row_split<32>(d_hi, d_lo, t, delta);
    if (!(dbg & 2)) store_rows_transposed_split(dt_hi, dt_lo, warp, lane, delta);
    if (!(dbg & 2)) store_rows_transposed_split(at_hi, at_lo, warp, lane, xrow);
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    if (!(dbg & 4)) my_db[0 * 48 + lane] += warp_column_sums(tmp, lane);
    issue_layer(0, 32, 32, 32);
    {
      const int64_t ntile = tile + gridDim.x;
      if (ntile < tiles) prefetch_act_chunks(in.g2, in.ld, ntile * kRows, t, pf);
    }
    wait_data();
    if (out.dx != nullptr) {
      tmem_load_row<32>(tmem_base, warp, 0, delta);
      warp_store_rows_coalesced(out.dx, row0 + warp * 32, M, bounce, lane, delta);
    }
    first = false;
  }
  // ---- flush: weight gradients from TMEM, bias gradients from shared memory, beta
  wait_dw();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (!first) {
    const int r = m64_row_of_lane(t);
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      float acc48[48];
      tmem_load_row<48>(tmem_base, warp, kDwCol[l], acc48);
      if (r >= 0 && r < kOut[l] && out.dw[l] != nullptr) {
#pragma unroll
        for (int k = 0; k < 48; ++k)
          if (k < kK[l]) atomicAdd(out.dw[l] + r * kK[l] + k, acc48[k]);
      }
    }
    for (int i = t; i < 5 * 48; i += kRows) {
      const int l = i / 48, j = i - l * 48;
      if (j < kOut[l] && out.db[l] != nullptr)
        atomicAdd(out.db[l] + j, dbacc[i] + dbacc[5 * 48 + i] + dbacc[2 * 5 * 48 + i] + dbacc[3 * 5 * 48 + i]);
    }
    const float s = warp_sum(dbeta_acc);
    if (lane == 0 && out.dbeta != nullptr) atomicAdd(out.dbeta, s);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<kBwdTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Column-split backward kernel (the one that runs; field_mlp_bwd_kernel above is kept as NRB_FIELD_BWD_SPLIT=1).
// Same mathematics and the same two GEMM chains per layer, re-organised around what the ncu / clock64 traces showed to
// bound the one-thread-per-row kernel: shared-memory bandwidth (operand staging + the tensor core re-reading both
// operands for every K = 8 instruction) and a fully serial stage -> issue -> wait chain with one warp per scheduler.
//   * a tile is worked on by 128 * S threads (S = 2 or 4): thread (h, r) owns columns [h*W, (h+1)*W), W = 32 / S, of
//     row r; one extra warp only issues tcgen05.mma (workers signal it through named barrier 1, it answers through
//     mbarriers), so the workers never stall behind the issue of the long weight-gradient chain;
//   * the weight-gradient GEMM [dW_l | db_l] += delta^T [input | 1] runs as kind::f16 on bf16 pairs: every value is
//     split into hi = top 16 bits and mid = bf16(x - hi), products hi*hi + hi*mid + mid*hi (relative error ~2^-16 per
//     product, averaged down further by the reduction over all samples - the tolerance is 1e-3).  K = 16 per
//     instruction and 2-byte elements halve both the staging bytes and the operand bytes the tensor core reads, and
//     the smaller tiles leave room to double-buffer delta^T / input^T, so the chain of layer l overlaps the staging
//     and the data-gradient GEMM of layer l-1.  hi and mid tiles of delta^T are adjacent and read as ONE 128-row A
//     operand (two instructions per K step give all four partial products);
//   * the data-gradient chain (the critical path, errors compound through five layers) stays 3xTF32, with its A operand
//     (this thread's delta row, hi / lo) written straight to tensor memory with tcgen05.st and read from there by the
//     ".ts" form of tcgen05.mma: no shared-memory staging or operand fetch for it;
//   * bias gradients come from the tensor cores: the input^T operand carries one extra row of ones (N = 40 / 56), so
//     column 32 (48) of every dW accumulator is sum_s delta = db;
//   * layer 1 (33 outputs = [sdf | emb]) feeds only the 32 emb columns to the data-gradient GEMM; the sdf row of W1 is
//     a rank-1 update on the CUDA cores.  Every delta tile is then 32 columns wide and chunk-aligned;
//   * global loads that are consumed layers later (x rows, SH basis, sdf / alpha terms, next tile's d feature rows) are
//     issued at the top of the tile.
// TMEM map (512 columns): the 48-column data-gradient accumulator, the five weight-gradient accumulators, and the
// data-gradient GEMM's A operand (delta as tf32 hi / lo).
constexpr int kBwd2TmemCols = 512;
__device__ constexpr int kDw2Col[5] = {192, 240, 288, 352, 400};  // TMEM column of each layer's dW (+db) accumulator
__device__ constexpr int kDw2N[5] = {48, 48, 64, 48, 48};         // N of the weight-gradient GEMM (inputs + ones + pad)
constexpr int kBwd2AHi = 448, kBwd2ALo = 480;
constexpr int kDt16Bytes = 64 * kRows * 2;  // delta^T as bf16: 64 rows (33 used) x 128 samples
constexpr int kAt16Bytes = 64 * kRows * 2;  // [input | ones]^T as bf16: 64 rows (<= 49 used) x 128 samples

struct FieldBwd2Smem {
  static constexpr int w1row0 = field_w_hi(5);           // sdf row of W1, 32 floats
  static constexpr int dt = w1row0 + 128;                // [2 buffers][hi, mid] delta^T
  static constexpr int at = dt + 4 * kDt16Bytes;         // [2 buffers][hi, mid] [input | ones]^T
  static constexpr int bounce = at + 4 * kAt16Bytes;     // 16 KB of per-warp bounce buffers
  static constexpr int mbar = bounce + 16384;            // data, dw[0], dw[1]
  static constexpr int tmem = mbar + 32;
  static constexpr int total = tmem + 16;
};

// byte offset of (row, 8-sample chunk) in a canonical K-major bf16 tile with 128 columns
__device__ __forceinline__ uint32_t tile16_offset(int row, int chunk8) {
  return static_cast<uint32_t>((row >> 3) * 2048 + chunk8 * 128 + (row & 7) * 16);
}

// four consecutive K elements -> packed bf16 hi (truncated) and mid (rounded residual)
__device__ __forceinline__ void split_bf16x4(float a0, float a1, float a2, float a3, uint2& hi, uint2& mid) {
  const uint32_t b0 = __float_as_uint(a0), b1 = __float_as_uint(a1), b2 = __float_as_uint(a2), b3 = __float_as_uint(a3);
  hi.x = __byte_perm(b0, b1, 0x7632);
  hi.y = __byte_perm(b2, b3, 0x7632);
  const float r0 = a0 - __uint_as_float(b0 & 0xFFFF0000u), r1 = a1 - __uint_as_float(b1 & 0xFFFF0000u);
  const float r2 = a2 - __uint_as_float(b2 & 0xFFFF0000u), r3 = a3 - __uint_as_float(b3 & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid.x) : "f"(r1), "f"(r0));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid.y) : "f"(r3), "f"(r2));
}

// (row, samples 4*c4 .. 4*c4+3) of a bf16 hi / mid tile pair
__device__ __forceinline__ void store_bf16x4(char* hi_tile, char* mid_tile, int row, int c4, float a0, float a1, float a2,
                                             float a3) {
  uint2 hi, mid;
  split_bf16x4(a0, a1, a2, a3, hi, mid);
  const uint32_t off = tile16_offset(row, c4 >> 1) + (c4 & 1) * 8;
  *reinterpret_cast<uint2*>(hi_tile + off) = hi;
  *reinterpret_cast<uint2*>(mid_tile + off) = mid;
}

// Transposed bf16 staging of W columns held row-per-lane (see store_part_transposed_split).
template <int W>
__device__ __forceinline__ void store_part_transposed_bf16(char* hi_tile, char* mid_tile, int quarter, int lane,
                                                           const float (&v)[W], int row_offset) {
  const int i = lane & 3;
  const int c4 = quarter * 8 + (lane >> 2);
#pragma unroll
  for (int g = 0; g < W / 4; ++g) {
    float a0 = v[4 * g], a1 = v[4 * g + 1], a2 = v[4 * g + 2], a3 = v[4 * g + 3];
    {
      const bool odd = (i & 1) != 0;
      const float s01 = odd ? a0 : a1, s23 = odd ? a2 : a3;
      const float r01 = __shfl_xor_sync(kFull, s01, 1), r23 = __shfl_xor_sync(kFull, s23, 1);
      if (odd) {
        a0 = r01;
        a2 = r23;
      } else {
        a1 = r01;
        a3 = r23;
      }
    }
    {
      const bool up = (i & 2) != 0;
      const float s0 = up ? a0 : a2, s1 = up ? a1 : a3;
      const float r0 = __shfl_xor_sync(kFull, s0, 2), r1 = __shfl_xor_sync(kFull, s1, 2);
      if (up) {
        a0 = r0;
        a1 = r1;
      } else {
        a2 = r0;
        a3 = r1;
      }
    }
    store_bf16x4(hi_tile, mid_tile, row_offset + 4 * g + i, c4, a0, a1, a2, a3);
  }
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}

// D[128, n] (+)= [A_hi; A_mid][128, 128] * B^T over bf16 hi / mid pairs: the hi and mid tiles of A (64 rows each) are
// adjacent in shared memory and form ONE 128-row operand, so two instructions per K step (x B_hi, x B_mid) produce
// all four partial products; rows 0..63 of D hold hi * (hi + mid), rows 64..127 mid * (hi + mid).  One thread.
__device__ __forceinline__ void issue_gemm_bf16_stacked(uint32_t d_tmem, int n, uint32_t a_stacked, uint32_t b_hi,
                                                        uint32_t b_mid, bool accumulate_first) {
  // kind::f16 instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24);
  uint64_t a = make_desc(a_stacked, 128, 2048);
  uint64_t bh = make_desc(b_hi, 128, 2048), bm = make_desc(b_mid, 128, 2048);
  bool acc = accumulate_first;
#pragma unroll
  for (int k = 0; k < kRows / 16; ++k) {  // K = 16 per instruction = two 16-byte chunks = 256 bytes
    mma_bf16(d_tmem, a, bm, idesc, acc);
    mma_bf16(d_tmem, a, bh, idesc, true);
    acc = true;
    a += 16;
    bh += 16;
    bm += 16;
  }
}

// Debug only (NRB_FIELD_BWD_DEBUG bit 2): thread 0 of CTA 0 stamps clock64() at the phase boundaries of its tiles.
__device__ long long g_bwd_trace[2048];
#define NRB_TRACE()                                                                                      \
  do {                                                                                                   \
    if ((dbg & 4) && t == 0 && blockIdx.x == 0 && trace_n < 1024) g_bwd_trace[trace_n++] = clock64();    \
  } while (0)

template <int S>
__global__ void __launch_bounds__(kRows * S + 32, 1) field_mlp_bwd_split_kernel(const __grid_constant__ FieldParams prm,
                                                                                const __grid_constant__ FieldBwdIn in,
                                                                                const __grid_constant__ FieldBwdOut out,
                                                                                int samples_per_ray, int64_t M, int dbg) {
  constexpr int W = 32 / S, T = kRows * S, NPF = 1024 / T, NSH = (16 * 32 + T - 1) / T;
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int h = warp >> 2, quarter = warp & 3, r = t & (kRows - 1), col0 = h * W;
  const bool issuer = warp == 4 * S;  // the extra warp
  auto dt_hi = [&](int b) { return smem + FieldBwd2Smem::dt + (2 * b) * kDt16Bytes; };
  auto dt_mid = [&](int b) { return smem + FieldBwd2Smem::dt + (2 * b + 1) * kDt16Bytes; };
  auto at_hi = [&](int b) { return smem + FieldBwd2Smem::at + (2 * b) * kAt16Bytes; };
  auto at_mid = [&](int b) { return smem + FieldBwd2Smem::at + (2 * b + 1) * kAt16Bytes; };
  float* w1row0 = reinterpret_cast<float*>(smem + FieldBwd2Smem::w1row0);
  uint64_t* mbar_data = reinterpret_cast<uint64_t*>(smem + FieldBwd2Smem::mbar);
  uint64_t* mbar_dw = mbar_data + 1;  // [2], one per operand buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FieldBwd2Smem::tmem);
  char* bounce = smem + FieldBwd2Smem::bounce + (warp & (4 * S - 1)) * (32 * W * 4);

  // W_l^T as the B operand of the data-gradient GEMM; layer 1 without its sdf row (rank-1 term below)
  stage_weight_transposed_split(prm.w[0], 32, 32, 32, smem + field_w_hi(0), smem + field_w_lo(0));
  stage_weight_transposed_split(prm.w[1] + 32, 32, 32, 32, smem + field_w_hi(1), smem + field_w_lo(1));
  stage_weight_transposed_split(prm.w[2], 32, 32, 48, smem + field_w_hi(2), smem + field_w_lo(2));
  stage_weight_transposed_split(prm.w[3], 32, 32, 32, smem + field_w_hi(3), smem + field_w_lo(3));
  stage_weight_transposed_split(prm.w[4], 32, 32, 32, smem + field_w_hi(4), smem + field_w_lo(4));
  if (t < 32) w1row0[t] = __ldg(prm.w[1] + t);
  // operand buffers start as zeros (delta^T rows 32..63 and the padding rows of input^T stay zero / finite)
  for (int e = t; e < (4 * kDt16Bytes + 4 * kAt16Bytes) / 16; e += T + 32)
    reinterpret_cast<uint4*>(smem + FieldBwd2Smem::dt)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // [ones, 0 x 7] rows of the input^T operand: rows 32..39 (layers with 32 inputs) and 48..55 (layer 2)
  auto write_ones_rows = [&](int b, int row_base, int stride) {
    for (int e = t; e < 8 * 16; e += stride) {
      const uint32_t v = (e & 7) == 0 ? 0x3F803F80u : 0u;  // bf16 1.0 pairs
      const uint32_t off = tile16_offset(row_base + (e & 7), e >> 3);
      *reinterpret_cast<uint4*>(at_hi(b) + off) = make_uint4(v, v, v, v);
      *reinterpret_cast<uint4*>(at_mid(b) + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  for (int b = 0; b < 2; ++b) {
    write_ones_rows(b, 32, T + 32);
    write_ones_rows(b, 48, T + 32);
  }
  if (warp == 0) tmem_alloc<kBwd2TmemCols>(tmem_slot);
  if (t == 0) {
    mbar_init(mbar_data, 1);
    mbar_init(mbar_dw, 1);
    mbar_init(mbar_dw + 1, 1);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t tiles = (M + kRows - 1) / kRows;
  const bool any = static_cast<int64_t>(blockIdx.x) < tiles;

  if (issuer) {
    // ---- the issuing warp: per layer wait for the workers' operands, then dIn chain (committed first: the workers
    // wait for it) and the weight-gradient chain on this layer's operand buffer
    bool first = true;
    int it = 0;
    int itrace = 1024;
#define NRB_ITRACE()                                                                                   \
  do {                                                                                                 \
    if ((dbg & 4) && blockIdx.x == 0 && itrace < 2048) g_bwd_trace[itrace++] = clock64();              \
  } while (0)
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int l = 4; l >= 0; --l, ++it) {
        const int b = it & 1;
        asm volatile("bar.sync 1, %0;" ::"n"(T + 32) : "memory");
        if (elect_one()) {
          NRB_ITRACE();  // 0: all workers arrived
          fence_after_sync();
          if (!(dbg & 2))
            issue_gemm_ts(tmem_base, kK[l], tmem_base + kBwd2AHi, tmem_base + kBwd2ALo, smem_u32(smem + field_w_hi(l)),
                          smem_u32(smem + field_w_lo(l)), 32, 32);
          mma_commit(mbar_data);
          NRB_ITRACE();  // 1: dIn chain issued
          if (!(dbg & 1))
            issue_gemm_bf16_stacked(tmem_base + kDw2Col[l], kDw2N[l], smem_u32(dt_hi(b)), smem_u32(at_hi(b)),
                                    smem_u32(at_mid(b)), !first);
          mma_commit(mbar_dw + b);
          NRB_ITRACE();  // 2: dW chain issued
        }
        __syncwarp();
      }
      first = false;
    }
  } else {
    // ---- workers
    const float beta = fabsf(__ldg(prm.beta)) + prm.beta_min;
    uint32_t phase_data = 0, phase_dw = 0, dw_pending = 0;  // bit b: operand buffer b
    int it = 0;
    float dbeta_acc = 0.0f;
    int trace_n = 0;
    // operands staged + this thread's TMEM reads complete -> arrive on the issuer's barrier
    auto signal_issuer = [&](int b) {
      fence_async_smem();
      fence_before_sync();
      asm volatile("bar.arrive 1, %0;" ::"n"(T + 32) : "memory");
      dw_pending |= 1u << b;
    };
    auto wait_data = [&]() {
      mbar_wait(mbar_data, phase_data);
      phase_data ^= 1;
      fence_after_sync();
    };
    auto wait_dw = [&](int b) {  // the chain that last read operand buffer b
      if ((dw_pending >> b) & 1u) {
        mbar_wait(mbar_dw + b, (phase_dw >> b) & 1u);
        phase_dw ^= 1u << b;
        dw_pending &= ~(1u << b);
      }
    };
    auto prefetch_act = [&](const float* __restrict__ fm, int64_t row0, float4 (&pf)[NPF]) {
#pragma unroll
      for (int q = 0; q < NPF; ++q) {
        const int e = t + q * T;
        const int j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
        pf[q] = __ldg(reinterpret_cast<const float4*>(fm + j * in.ld + row0 + 4 * c));
      }
    };
    auto commit_act = [&](int b, const float4 (&pf)[NPF]) {  // rows 0..31 of input^T
#pragma unroll
      for (int q = 0; q < NPF; ++q) {
        const int e = t + q * T;
        const int j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
        store_bf16x4(at_hi(b), at_mid(b), j, c, pf[q].x, pf[q].y, pf[q].z, pf[q].w);
      }
    };
    auto stage_delta = [&](int b, const float (&v)[W]) {
      tmem_store_row_split<W>(tmem_base, warp, kBwd2AHi, kBwd2ALo, col0, v);
      store_part_transposed_bf16<W>(dt_hi(b), dt_mid(b), quarter, lane, v, col0);
    };
    auto load_din = [&](float (&v)[W]) { tmem_load_cols<W>(tmem_base, warp, col0, v); };
    auto apply_mask = [&](uint32_t m, float (&v)[W]) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = ((m >> (col0 + j)) & 1u) ? v[j] : 0.0f;
    };

    const uint32_t last_ray = static_cast<uint32_t>((M - 1) / samples_per_ray);
    float4 pfa[NPF], pfb[NPF];  // two prefetch sets: every activation tile is requested two layers before its use
    float4 df[W / 4];  // this thread's share of the tile's d feature rows, loaded one tile ahead
    if (any) {
      prefetch_act(in.g2, static_cast<int64_t>(blockIdx.x) * kRows, pfa);
      warp_load_part_coalesced<W>(in.dfeature, static_cast<int64_t>(blockIdx.x) * kRows + quarter * 32, M, col0, lane, df);
    }
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int64_t row0 = tile * kRows;
      const int64_t row = row0 + r;
      const bool ok = row < M;
      const uint32_t m_h1 = __ldg(in.masks + row), m_g1 = __ldg(in.masks + in.ld + row), m_g2 = __ldg(in.masks + 2 * in.ld + row);
      float delta[W], demb[W];
      NRB_TRACE();  // 0: tile start
      prefetch_act(in.g1, row0, pfb);
      warp_bounce_part_to_rows<W>(bounce, lane, df, delta);
      // loads consumed layers later are issued now: x rows (layer 0), SH basis (layer 2), sdf / alpha terms (layer 1)
      float4 xpf[W / 4];
      warp_load_part_coalesced<W>(in.x, row0 + quarter * 32, M, col0, lane, xpf);
      float shv[NSH][4];
#pragma unroll
      for (int q = 0; q < NSH; ++q) {
        const int e = t + q * T;
        const int k = (e & 7) + 8 * ((e >> 8) & 1), c = (e >> 3) & 31;
        const uint32_t s0 = static_cast<uint32_t>(row0) + 4u * c;  // first sample of the chunk (M < 2^31)
        const uint32_t ray0 = s0 / static_cast<uint32_t>(samples_per_ray);
        const uint32_t rem = s0 - ray0 * static_cast<uint32_t>(samples_per_ray);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t ray = ray0;
          if (samples_per_ray >= 4) {
            ray += (rem + i >= static_cast<uint32_t>(samples_per_ray)) ? 1u : 0u;
          } else {
            ray += (rem + i) / static_cast<uint32_t>(samples_per_ray);
          }
          shv[q][i] = __ldg(in.sh + static_cast<int64_t>(min(ray, last_ray)) * 16 + k);
        }
      }
      const int64_t rc = ok ? row : (M - 1);
      const float a_v = __ldg(in.alpha + rc), sd_v = __ldg(in.sdf + rc);
      const float da_v = in.dalpha != nullptr ? __ldg(in.dalpha + rc) : 0.0f;
      const float ds_v = in.dsdf != nullptr ? __ldg(in.dsdf + rc) : 0.0f;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        delta[j] = ok ? delta[j] : 0.0f;
        demb[j] = delta[j];  // residual branch
      }
      NRB_TRACE();  // 1
      // ---- layer 4 (mlp_feature.layers.2): delta = d feature, input g2
      int b = it & 1;
      wait_dw(b);
      NRB_TRACE();  // 2
      stage_delta(b, delta);
      commit_act(b, pfa);
      write_ones_rows(b, 32, T);
      NRB_TRACE();  // 3
      signal_issuer(b);
      ++it;
      prefetch_act(in.emb, row0, pfa);
      wait_data();
      NRB_TRACE();  // 4
      load_din(delta);
      apply_mask(m_g2, delta);
      // ---- layer 3 (mlp_feature.layers.1): input g1
      b = it & 1;
      wait_dw(b);
      NRB_TRACE();  // 5
      stage_delta(b, delta);
      commit_act(b, pfb);
      write_ones_rows(b, 32, T);
      NRB_TRACE();  // 6
      signal_issuer(b);
      ++it;
      prefetch_act(in.h1, row0, pfb);
      wait_data();
      NRB_TRACE();  // 7
      load_din(delta);
      apply_mask(m_g1, delta);
      // ---- layer 2 (mlp_feature.layers.0): input [emb | sh | 1]
      b = it & 1;
      wait_dw(b);
      NRB_TRACE();  // 8
      stage_delta(b, delta);
      commit_act(b, pfa);
#pragma unroll
      for (int q = 0; q < NSH; ++q) {  // rows 32..47 of the input^T tile: the ray's SH basis
        const int e = t + q * T;
        if (e < 16 * 32) {
          const int k = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
          store_bf16x4(at_hi(b), at_mid(b), 32 + k, c, shv[q][0], shv[q][1], shv[q][2], shv[q][3]);
        }
      }
      NRB_TRACE();  // 9
      signal_issuer(b);
      ++it;
      {
        const int64_t ntile = tile + gridDim.x;
        if (ntile < tiles) {
          prefetch_act(in.g2, ntile * kRows, pfa);
          warp_load_part_coalesced<W>(in.dfeature, ntile * kRows + quarter * 32, M, col0, lane, df);
        }
      }
      wait_data();
      NRB_TRACE();  // 10
      load_din(delta);  // the SH part of the input carries no gradient
#pragma unroll
      for (int j = 0; j < W; ++j) demb[j] += delta[j];
      // ---- layer 1 (mlp_geo.layers.1): delta = [d sdf | d emb]; input h1
      float dsdf_v = 0.0f;
      if (ok) {
        const float sg = a_v * (1.0f - a_v);
        dsdf_v = ds_v - da_v * beta * sg;
        if (h == 0) dbeta_acc -= da_v * sd_v * sg;
      }
      b = it & 1;
      wait_dw(b);
      NRB_TRACE();  // 11
      stage_delta(b, demb);  // accumulator rows 0..31 <-> W1 rows 1..32
      if (h == 0) {          // accumulator row 32 <-> W1 row 0 (sdf)
        const uint32_t hb = __float_as_uint(dsdf_v) & 0xFFFF0000u;
        const __nv_bfloat16 mid = __float2bfloat16_rn(dsdf_v - __uint_as_float(hb));
        const uint32_t off = tile16_offset(32, r >> 3) + (r & 7) * 2;
        *reinterpret_cast<uint16_t*>(dt_hi(b) + off) = static_cast<uint16_t>(hb >> 16);
        *reinterpret_cast<__nv_bfloat16*>(dt_mid(b) + off) = mid;
      }
      commit_act(b, pfb);
      write_ones_rows(b, 32, T);
      NRB_TRACE();  // 12
      signal_issuer(b);
      ++it;
      float xrow[W];
      warp_bounce_part_to_rows<W>(bounce, lane, xpf, xrow);
      wait_data();
      NRB_TRACE();  // 13
      load_din(delta);
#pragma unroll
      for (int j = 0; j < W; ++j) delta[j] = fmaf(dsdf_v, w1row0[col0 + j], delta[j]);
      apply_mask(m_h1, delta);
      // ---- layer 0 (mlp_geo.layers.0): input x (row-major in memory: transposed through registers)
      b = it & 1;
      wait_dw(b);
      NRB_TRACE();  // 14
      stage_delta(b, delta);
      store_part_transposed_bf16<W>(at_hi(b), at_mid(b), quarter, lane, xrow, col0);
      write_ones_rows(b, 32, T);
      NRB_TRACE();  // 15
      signal_issuer(b);
      ++it;
      wait_data();
      NRB_TRACE();  // 16
      if (out.dx != nullptr) {
        load_din(delta);
        warp_store_part_coalesced<W>(out.dx, row0 + quarter * 32, M, col0, bounce, lane, delta);
      }
      NRB_TRACE();  // 17
    }
    wait_dw(0);
    wait_dw(1);
    const float sb = warp_sum(dbeta_acc);
    if (any && h == 0 && lane == 0 && out.dbeta != nullptr) atomicAdd(out.dbeta, sb);
  }
  // ---- flush: weight and bias gradients from TMEM
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (any && h == 0) {
    const int ar = r & 63;  // rows 0..63: hi part, 64..127: mid part of the same weight-gradient row
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      float acc[64];
      tmem_load_row<64>(tmem_base, warp, kDw2Col[l], acc);
      const int wrow = (l == 1) ? (ar == 32 ? 0 : ar + 1) : ar;  // row of the weight matrix
      if (wrow < kOut[l] && ar <= 32) {
        if (out.dw[l] != nullptr) {
#pragma unroll
          for (int k = 0; k < 48; ++k)
            if (k < kK[l]) atomicAdd(out.dw[l] + wrow * kK[l] + k, acc[k]);
        }
        if (out.db[l] != nullptr) atomicAdd(out.db[l] + wrow, l == 2 ? acc[48] : acc[32]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<kBwd2TmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Two tiles in flight per CTA.  The per-layer chain stage -> signal -> issue -> commit -> wake -> tcgen05.ld is serial
// for one tile (clock64 trace: ~4 K cycles per layer with the SM mostly waiting), so this variant runs TWO independent
// 128-sample tiles per CTA, each owned by a group of 8 warps (thread = 16 columns of a row), phase-shifted by
// construction because one warp issues the tensor-core chains of both groups alternately.  Resources per group: one
// delta^T / input^T operand buffer (the two buffers of the split kernel), 48 + 64 TMEM columns (data-gradient
// accumulator, A operand); the five weight-gradient accumulators (256 columns) are shared: both groups' chains add into
// them, issued by the same thread and therefore ordered.  While one group waits for its chain, the other stages.
constexpr int kPairThreads = 2 * 256 + 96;  // two groups of 8 worker warps + three issuing warps
constexpr int kPairDwCol0 = 96;  // dIn accumulators at 0 and 48, dW at 96 .. 351, A operands at 352 .. 479
__device__ constexpr int kPairDwOff[5] = {0, 48, 96, 160, 208};
constexpr int kPairACol0 = 352;

struct FieldBwdPairSmem {
  static constexpr int w1row0 = field_w_hi(5);
  static constexpr int dt = w1row0 + 128;                // [2 groups][hi, mid] delta^T
  static constexpr int at = dt + 4 * kDt16Bytes;         // [2 groups][hi, mid] [input | ones]^T
  static constexpr int bounce = at + 4 * kAt16Bytes;     // 16 warps x 2 KB
  static constexpr int mbar = bounce + 32768;            // data[2], dw[2], ready[2]
  static constexpr int tmem = mbar + 48;
  static constexpr int total = tmem + 16;
};

__global__ void __launch_bounds__(kPairThreads, 1) field_mlp_bwd_pair_kernel(const __grid_constant__ FieldParams prm,
                                                                             const __grid_constant__ FieldBwdIn in,
                                                                             const __grid_constant__ FieldBwdOut out,
                                                                             int samples_per_ray, int64_t M, int dbg) {
  constexpr int W = 16, T = 256, NPF = 4, NSH = 2;
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool issuer = warp >= 16;            // warps 16 / 17: data-gradient chains of group 0 / 1, warp 18: weight-gradient chains
  const int g = (warp >> 3) & 1;            // tile group
  const int tg = t & 255;                    // thread index within the group
  const int h = (warp >> 2) & 1, quarter = warp & 3, r = t & (kRows - 1), col0 = h * W;
  char* dt_hi = smem + FieldBwdPairSmem::dt + (2 * g) * kDt16Bytes;
  char* dt_mid = dt_hi + kDt16Bytes;
  char* at_hi = smem + FieldBwdPairSmem::at + (2 * g) * kAt16Bytes;
  char* at_mid = at_hi + kAt16Bytes;
  float* w1row0 = reinterpret_cast<float*>(smem + FieldBwdPairSmem::w1row0);
  uint64_t* mbar_data = reinterpret_cast<uint64_t*>(smem + FieldBwdPairSmem::mbar);  // [2]
  uint64_t* mbar_dw = mbar_data + 2;                                                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FieldBwdPairSmem::tmem);
  char* bounce = smem + FieldBwdPairSmem::bounce + (warp & 15) * (32 * W * 4);

  stage_weight_transposed_split(prm.w[0], 32, 32, 32, smem + field_w_hi(0), smem + field_w_lo(0));
  stage_weight_transposed_split(prm.w[1] + 32, 32, 32, 32, smem + field_w_hi(1), smem + field_w_lo(1));
  stage_weight_transposed_split(prm.w[2], 32, 32, 48, smem + field_w_hi(2), smem + field_w_lo(2));
  stage_weight_transposed_split(prm.w[3], 32, 32, 32, smem + field_w_hi(3), smem + field_w_lo(3));
  stage_weight_transposed_split(prm.w[4], 32, 32, 32, smem + field_w_hi(4), smem + field_w_lo(4));
  if (t < 32) w1row0[t] = __ldg(prm.w[1] + t);
  for (int e = t; e < (4 * kDt16Bytes + 4 * kAt16Bytes) / 16; e += kPairThreads)
    reinterpret_cast<uint4*>(smem + FieldBwdPairSmem::dt)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // [ones, 0 x 7] rows of a group's input^T operand: rows 32..39 (32-input layers) and 48..55 (layer 2)
  auto write_ones_rows = [&](char* hi_tile, char* mid_tile, int row_base, int first, int stride) {
    for (int e = first; e < 8 * 16; e += stride) {
      const uint32_t v = (e & 7) == 0 ? 0x3F803F80u : 0u;  // bf16 1.0 pairs
      const uint32_t off = tile16_offset(row_base + (e & 7), e >> 3);
      *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(v, v, v, v);
      *reinterpret_cast<uint4*>(mid_tile + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  for (int b = 0; b < 2; ++b) {
    char* hi_tile = smem + FieldBwdPairSmem::at + (2 * b) * kAt16Bytes;
    write_ones_rows(hi_tile, hi_tile + kAt16Bytes, 32, t, kPairThreads);
    write_ones_rows(hi_tile, hi_tile + kAt16Bytes, 48, t, kPairThreads);
  }
  if (warp == 0) tmem_alloc<kBwd2TmemCols>(tmem_slot);
  if (t == 0)
    for (int i = 0; i < 4; ++i) mbar_init(mbar_data + i, 1);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t tiles = (M + kRows - 1) / kRows;
  // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0, 1, ...; group g takes k = g, g + 2, ...
  const int64_t mine = static_cast<int64_t>(blockIdx.x) < tiles ? (tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const bool any = mine > 0;

  if (issuer) {
    // Three issuing warps.  Warps 16 / 17 issue the data-gradient chain of group 0 / 1 as soon as THAT group has staged
    // its operands (the group waits for it: nothing else sits on its critical path).  Warp 18 issues the weight-gradient
    // chains of both groups in strict alternation: ONE thread issues everything that accumulates into the shared dW
    // accumulators, so those accumulations are ordered.  (History: one issuer for everything 1.64 ms - it was busy
    // ~60 % of the time; serving whichever group is ready first through polled mbarriers 1.81 ms - a spinning issuer
    // steals issue slots; data / weight issuers 1.50 ms.)
    if (warp < 18) {
      const int gg = warp - 16;
      int itrace = 1024;
      for (int64_t k = gg; k < mine; k += 2) {
#pragma unroll 1
        for (int l = 4; l >= 0; --l) {
          if (gg == 0)
            asm volatile("bar.sync 1, %0;" ::"n"(T + 64) : "memory");
          else
            asm volatile("bar.sync 2, %0;" ::"n"(T + 64) : "memory");
          if (elect_one()) {
            if (gg == 0) NRB_ITRACE();
            fence_after_sync();
            const uint32_t a_col = tmem_base + kPairACol0 + 64 * gg;
            issue_gemm_ts(tmem_base + 48 * gg, kK[l], a_col, a_col + 32, smem_u32(smem + field_w_hi(l)),
                          smem_u32(smem + field_w_lo(l)), 32, 32);
            mma_commit(mbar_data + gg);
            if (gg == 0) NRB_ITRACE();
          }
          __syncwarp();
        }
      }
    } else {
      bool first = true;
      for (int64_t k0 = 0; k0 < mine; k0 += 2) {
#pragma unroll 1
        for (int l = 4; l >= 0; --l) {
#pragma unroll 1
          for (int gg = 0; gg < 2; ++gg) {
            if (k0 + gg >= mine) continue;
            if (gg == 0)
              asm volatile("bar.sync 1, %0;" ::"n"(T + 64) : "memory");
            else
              asm volatile("bar.sync 2, %0;" ::"n"(T + 64) : "memory");
            if (elect_one()) {
              fence_after_sync();
              char* dth = smem + FieldBwdPairSmem::dt + (2 * gg) * kDt16Bytes;
              char* ath = smem + FieldBwdPairSmem::at + (2 * gg) * kAt16Bytes;
              issue_gemm_bf16_stacked(tmem_base + kPairDwCol0 + kPairDwOff[l], kDw2N[l], smem_u32(dth), smem_u32(ath),
                                      smem_u32(ath + kAt16Bytes), !(first && gg == 0));
              mma_commit(mbar_dw + gg);
            }
            __syncwarp();
          }
        }
        first = false;
      }
    }
  } else {
    const float beta = fabsf(__ldg(prm.beta)) + prm.beta_min;
    uint32_t phase_data = 0, phase_dw = 0;
    bool dw_pending = false;
    float dbeta_acc = 0.0f;
    int trace_n = 0;
    const uint32_t a_hi_col = kPairACol0 + 64 * g, a_lo_col = a_hi_col + 32, din_col = 48 * g;
    const uint32_t last_ray = static_cast<uint32_t>((M - 1) / samples_per_ray);
    auto signal_issuer = [&]() {
      NRB_TRACE();  // staged
      fence_async_smem();
      fence_before_sync();
      if (g == 0)
        asm volatile("bar.arrive 1, %0;" ::"n"(T + 64) : "memory");
      else
        asm volatile("bar.arrive 2, %0;" ::"n"(T + 64) : "memory");
      dw_pending = true;
    };
    auto wait_data = [&]() {
      mbar_wait(mbar_data + g, phase_data);
      phase_data ^= 1;
      fence_after_sync();
      NRB_TRACE();  // dIn done
    };
    auto wait_dw = [&]() {  // the chain that last read this group's operand buffer
      NRB_TRACE();  // before wait_dw
      if (dw_pending) {
        mbar_wait(mbar_dw + g, phase_dw);
        phase_dw ^= 1;
        dw_pending = false;
      }
      NRB_TRACE();  // after wait_dw
    };
    auto prefetch_act = [&](const float* __restrict__ fm, int64_t row0, float4 (&pf)[NPF]) {
#pragma unroll
      for (int q = 0; q < NPF; ++q) {
        const int e = tg + q * T;
        const int j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
        pf[q] = __ldg(reinterpret_cast<const float4*>(fm + j * in.ld + row0 + 4 * c));
      }
    };
    auto commit_act = [&](const float4 (&pf)[NPF]) {  // rows 0..31 of input^T
#pragma unroll
      for (int q = 0; q < NPF; ++q) {
        const int e = tg + q * T;
        const int j = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
        store_bf16x4(at_hi, at_mid, j, c, pf[q].x, pf[q].y, pf[q].z, pf[q].w);
      }
    };
    auto stage_delta = [&](const float (&v)[W]) {
      tmem_store_row_split<W>(tmem_base, warp, a_hi_col, a_lo_col, col0, v);
      store_part_transposed_bf16<W>(dt_hi, dt_mid, quarter, lane, v, col0);
    };
    auto load_din = [&](float (&v)[W]) { tmem_load_cols<W>(tmem_base, warp, din_col + col0, v); };
    auto apply_mask = [&](uint32_t m, float (&v)[W]) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = ((m >> (col0 + j)) & 1u) ? v[j] : 0.0f;
    };

    float4 pf[NPF];
    for (int64_t k = g; k < mine; k += 2) {
      const int64_t tile = blockIdx.x + k * gridDim.x;
      const int64_t row0 = tile * kRows;
      const int64_t row = row0 + r;
      const bool ok = row < M;
      const uint32_t m_h1 = __ldg(in.masks + row), m_g1 = __ldg(in.masks + in.ld + row), m_g2 = __ldg(in.masks + 2 * in.ld + row);
      float delta[W], demb[W];
      prefetch_act(in.g2, row0, pf);
      {
        float4 df[W / 4];
        warp_load_part_coalesced<W>(in.dfeature, row0 + quarter * 32, M, col0, lane, df);
        warp_bounce_part_to_rows<W>(bounce, lane, df, delta);
      }
#pragma unroll
      for (int j = 0; j < W; ++j) {
        delta[j] = ok ? delta[j] : 0.0f;
        demb[j] = delta[j];  // residual branch
      }
      // ---- layer 4 (mlp_feature.layers.2): delta = d feature, input g2
      wait_dw();
      stage_delta(delta);
      commit_act(pf);
      write_ones_rows(at_hi, at_mid, 32, tg, T);
      signal_issuer();
      prefetch_act(in.g1, row0, pf);
      wait_data();
      load_din(delta);
      apply_mask(m_g2, delta);
      // ---- layer 3 (mlp_feature.layers.1): input g1
      wait_dw();
      stage_delta(delta);
      commit_act(pf);
      signal_issuer();
      prefetch_act(in.emb, row0, pf);
      float shv[NSH][4];  // the ray's SH basis for rows 32..47 of layer 2's input^T
#pragma unroll
      for (int q = 0; q < NSH; ++q) {
        const int e = tg + q * T;
        const int kk = (e & 7) + 8 * ((e >> 8) & 1), c = (e >> 3) & 31;
        const uint32_t s0 = static_cast<uint32_t>(row0) + 4u * c;
        const uint32_t ray0 = s0 / static_cast<uint32_t>(samples_per_ray);
        const uint32_t rem = s0 - ray0 * static_cast<uint32_t>(samples_per_ray);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t ray = ray0;
          if (samples_per_ray >= 4) {
            ray += (rem + i >= static_cast<uint32_t>(samples_per_ray)) ? 1u : 0u;
          } else {
            ray += (rem + i) / static_cast<uint32_t>(samples_per_ray);
          }
          shv[q][i] = __ldg(in.sh + static_cast<int64_t>(min(ray, last_ray)) * 16 + kk);
        }
      }
      wait_data();
      load_din(delta);
      apply_mask(m_g1, delta);
      // ---- layer 2 (mlp_feature.layers.0): input [emb | sh | 1]
      wait_dw();
      stage_delta(delta);
      commit_act(pf);
#pragma unroll
      for (int q = 0; q < NSH; ++q) {
        const int e = tg + q * T;
        const int kk = (e & 7) + 8 * (e >> 8), c = (e >> 3) & 31;
        store_bf16x4(at_hi, at_mid, 32 + kk, c, shv[q][0], shv[q][1], shv[q][2], shv[q][3]);
      }
      signal_issuer();
      prefetch_act(in.h1, row0, pf);
      const int64_t rc = ok ? row : (M - 1);
      const float a_v = __ldg(in.alpha + rc), sd_v = __ldg(in.sdf + rc);
      const float da_v = in.dalpha != nullptr ? __ldg(in.dalpha + rc) : 0.0f;
      const float ds_v = in.dsdf != nullptr ? __ldg(in.dsdf + rc) : 0.0f;
      wait_data();
      load_din(delta);  // the SH part of the input carries no gradient
#pragma unroll
      for (int j = 0; j < W; ++j) demb[j] += delta[j];
      // ---- layer 1 (mlp_geo.layers.1): delta = [d sdf | d emb]; input h1
      float dsdf_v = 0.0f;
      if (ok) {
        const float sg = a_v * (1.0f - a_v);
        dsdf_v = ds_v - da_v * beta * sg;
        if (h == 0) dbeta_acc -= da_v * sd_v * sg;
      }
      wait_dw();
      stage_delta(demb);  // accumulator rows 0..31 <-> W1 rows 1..32
      if (h == 0) {       // accumulator row 32 <-> W1 row 0 (sdf)
        const uint32_t hb = __float_as_uint(dsdf_v) & 0xFFFF0000u;
        const __nv_bfloat16 mid = __float2bfloat16_rn(dsdf_v - __uint_as_float(hb));
        const uint32_t off = tile16_offset(32, r >> 3) + (r & 7) * 2;
        *reinterpret_cast<uint16_t*>(dt_hi + off) = static_cast<uint16_t>(hb >> 16);
        *reinterpret_cast<__nv_bfloat16*>(dt_mid + off) = mid;
      }
      commit_act(pf);
      write_ones_rows(at_hi, at_mid, 32, tg, T);  // rows 32..39 held SH values during layer 2
      signal_issuer();
      warp_load_part_coalesced<W>(in.x, row0 + quarter * 32, M, col0, lane, pf);  // x rows (layer 0's input)
      wait_data();
      load_din(delta);
#pragma unroll
      for (int j = 0; j < W; ++j) delta[j] = fmaf(dsdf_v, w1row0[col0 + j], delta[j]);
      apply_mask(m_h1, delta);
      // ---- layer 0 (mlp_geo.layers.0): input x (row-major in memory: transposed through registers)
      {
        float xrow[W];
        warp_bounce_part_to_rows<W>(bounce, lane, pf, xrow);
        wait_dw();
        stage_delta(delta);
        store_part_transposed_bf16<W>(at_hi, at_mid, quarter, lane, xrow, col0);
      }
      signal_issuer();
      wait_data();
      if (out.dx != nullptr) {
        load_din(delta);
        warp_store_part_coalesced<W>(out.dx, row0 + quarter * 32, M, col0, bounce, lane, delta);
      }
    }
    wait_dw();
    const float sb = warp_sum(dbeta_acc);
    if (h == 0 && lane == 0 && out.dbeta != nullptr && sb != 0.0f) atomicAdd(out.dbeta, sb);
  }
  // ---- flush: weight and bias gradients from TMEM (group 0's first column group reads them)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (any && !issuer && g == 0 && h == 0) {
    const int ar = r & 63;  // rows 0..63: hi part, 64..127: mid part of the same weight-gradient row
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      float acc[64];
      tmem_load_row<64>(tmem_base, warp, kPairDwCol0 + kPairDwOff[l], acc);
      const int wrow = (l == 1) ? (ar == 32 ? 0 : ar + 1) : ar;
      if (wrow < kOut[l] && ar <= 32) {
        if (out.dw[l] != nullptr) {
#pragma unroll
          for (int kk = 0; kk < 48; ++kk)
            if (kk < kK[l]) atomicAdd(out.dw[l] + wrow * kK[l] + kk, acc[kk]);
        }
        if (out.db[l] != nullptr) atomicAdd(out.db[l] + wrow, l == 2 ? acc[48] : acc[32]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<kBwd2TmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Generic probe of descriptor conventions: P and Q are [128,32] row-major host-provided matrices staged as canonical
// tiles; one chain of `ksteps` tf32 MMAs is issued with caller-chosen majors, M, N, LBO/SBO and per-step address
// advances; the [128 lanes][32 columns] accumulator block is dumped.
struct ProbeCfg {
  int a_major, b_major, m, n;
  int a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, ksteps;
};

__global__ void __launch_bounds__(kRows) tc_probe_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                         ProbeCfg cfg, float* __restrict__ dump) {
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5;
  char* p_hi = smem;
  char* p_lo = p_hi + kRows * 32 * 4;
  char* q_hi = p_lo + kRows * 32 * 4;
  char* q_lo = q_hi + kRows * 32 * 4;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(q_lo + kRows * 32 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  float p[32], q[32];
  load_row32(P + t * 32, p);
  load_row32(Q + t * 32, q);
  store_row_split<32>(p_hi, p_lo, t, p);
  store_row_split<32>(q_hi, q_lo, t, q);
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  if (t == 0) mbar_init(mbar, 1);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (t == 0) {
    const uint32_t idesc = make_idesc_m(cfg.m, cfg.n, cfg.a_major, cfg.b_major);
    for (int k = 0; k < cfg.ksteps; ++k) {
      const uint64_t ad = make_desc(smem_u32(p_hi) + k * cfg.a_step, cfg.a_lbo, cfg.a_sbo);
      const uint64_t bd = make_desc(smem_u32(q_hi) + k * cfg.b_step, cfg.b_lbo, cfg.b_sbo);
      mma_tf32(tmem_base, ad, bd, idesc, k > 0);
    }
    mma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  fence_after_sync();
  float acc[32];
  tmem_load_row<32>(tmem_base, warp, 0, acc);
  for (int j = 0; j < 32; ++j) dump[t * 32 + j] = acc[j];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<64>(tmem_base);
}

// Single linear layer y = x W^T + b (optionally ReLU) through the same building blocks: unit test of the descriptor
// and layout conventions.  K in {32, 48}, N (padded) in {32, 48}.
__global__ void __launch_bounds__(kRows) tc_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, int K, int n_out, int relu,
                                                          int64_t M, float* __restrict__ y, int dbg) {
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5;
  const int N = (n_out + 15) & ~15;
  char* w_hi = smem;
  char* w_lo = w_hi + 48 * 48 * 4;
  char* a_hi = w_lo + 48 * 48 * 4;
  char* a_lo = a_hi + kRows * 48 * 4;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(a_lo + kRows * 48 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  stage_weight_split(w, n_out, N, K, w_hi, w_lo);
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
  if (t == 0) mbar_init(mbar, 1);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t phase = 0;
  const int64_t tiles = (M + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row = tile * kRows + t;
    const bool ok = row < M;
    const int64_t rr = ok ? row : (M - 1);
    float v[48];
#pragma unroll
    for (int j = 0; j < 48; ++j) v[j] = (j < K) ? __ldg(x + rr * K + j) : 0.0f;
    if (!(dbg & 2)) {
      if (K == 32) {
        float v32[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v32[j] = v[j];
        store_row_split<32>(a_hi, a_lo, t, v32);
      } else {
        store_row_split<48>(a_hi, a_lo, t, v);
      }
    }
    if (!(dbg & 8)) fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (!(dbg & 1)) {
      if (t == 0) {
        fence_after_sync();
        issue_gemm(tmem_base, 128, N, smem_u32(a_hi), smem_u32(a_lo), K, smem_u32(w_hi), smem_u32(w_lo), K, K, false);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
    }
    fence_after_sync();
    float acc[48];
    if (!(dbg & 4)) {
      tmem_load_row<48>(tmem_base, warp, 0, acc);
    } else {
#pragma unroll
      for (int j = 0; j < 48; ++j) acc[j] = v[j];
    }
    if (ok) {
      for (int j = 0; j < n_out; ++j) {
        float r = acc[j] + (b != nullptr ? __ldg(b + j) : 0.0f);
        y[row * n_out + j] = relu ? fmaxf(r, 0.0f) : r;
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  if (warp == 0) tmem_free<kTmemCols>(tmem_base);
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_tc_linear(const float* x, const float* w, const float* b, int32_t K, int32_t n_out, int32_t relu,
                             int64_t M, float* y, nrb_stream_t stream) {
  NRB_REQUIRE(x && w && y && M >= 0, NRB_ERR_BAD_ARG, "nrb_tc_linear: null pointer or negative M");
  NRB_REQUIRE((K == 32 || K == 48) && n_out >= 1 && n_out <= 48, NRB_ERR_UNSUPPORTED,
              "nrb_tc_linear: K must be 32 or 48 and n_out <= 48");
  if (M == 0) return NRB_OK;
  const size_t smem = 2 * 48 * 48 * 4 + 2 * tc::kRows * 48 * 4 + 16;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(tc_linear_kernel), static_cast<int>(smem));
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_tc_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + tc::kRows - 1) / tc::kRows;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, 2 * sm_count()));
  static const int dbg = std::getenv("NRB_TC_LINEAR_DEBUG") ? std::atoi(std::getenv("NRB_TC_LINEAR_DEBUG")) : 0;
  tc_linear_kernel<<<grid, tc::kRows, smem, static_cast<cudaStream_t>(stream)>>>(x, w, b, K, n_out, relu, M, y, dbg);
  return finish_launch("nrb_tc_linear");
}

extern "C" int nrb_tc_probe(const float* P, const float* Q, const int32_t* cfg11, float* dump, nrb_stream_t stream) {
  NRB_REQUIRE(P && Q && cfg11 && dump, NRB_ERR_BAD_ARG, "nrb_tc_probe: null pointer");
  ProbeCfg c{cfg11[0], cfg11[1], cfg11[2], cfg11[3], cfg11[4], cfg11[5], cfg11[6], cfg11[7], cfg11[8], cfg11[9], cfg11[10]};
  const size_t smem = 4 * tc::kRows * 32 * 4 + 16;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(tc_probe_kernel), static_cast<int>(smem));
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_tc_probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  tc_probe_kernel<<<1, tc::kRows, smem, static_cast<cudaStream_t>(stream)>>>(P, Q, c, dump);
  return finish_launch("nrb_tc_probe");
}

static FieldParams to_params(const nrb_field_mlp_t* p) {
  FieldParams prm;
  for (int l = 0; l < 5; ++l) {
    prm.w[l] = p->weights[l];
    prm.b[l] = p->biases[l];
  }
  prm.beta = p->beta;
  prm.beta_min = p->beta_min;
  return prm;
}

extern "C" int64_t nrb_field_saved_ld(int64_t M) { return (M + tc::kRows - 1) / tc::kRows * tc::kRows; }

extern "C" int nrb_field_mlp_fwd(const nrb_field_mlp_t* p, const float* x, const float* sh, int32_t samples_per_ray,
                                 int64_t M, float* feature, float* sdf, float* alpha, const nrb_field_saved_t* saved,
                                 nrb_stream_t stream) {
  NRB_REQUIRE(p && x && sh && feature && sdf && alpha && M >= 0 && samples_per_ray > 0, NRB_ERR_BAD_ARG,
              "nrb_field_mlp_fwd: null pointer or bad size");
  for (int l = 0; l < 5; ++l) NRB_REQUIRE(p->weights[l] != nullptr, NRB_ERR_BAD_ARG, "nrb_field_mlp_fwd: weights[%d] is null", l);
  NRB_REQUIRE(p->beta != nullptr, NRB_ERR_BAD_ARG, "nrb_field_mlp_fwd: beta is null");
  NRB_REQUIRE(aligned16(x) && aligned16(sh) && aligned16(feature), NRB_ERR_ALIGNMENT,
              "nrb_field_mlp_fwd: x, sh and feature must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  FieldSaved sv{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  if (saved != nullptr && saved->h1 != nullptr) {
    NRB_REQUIRE(saved->emb && saved->g1 && saved->g2 && saved->masks, NRB_ERR_BAD_ARG,
                "nrb_field_mlp_fwd: saved activations must be given together");
    NRB_REQUIRE(saved->ld >= M && saved->ld % tc::kRows == 0, NRB_ERR_BAD_ARG,
                "nrb_field_mlp_fwd: saved->ld must be nrb_field_saved_ld(M)");
    sv = FieldSaved{saved->h1, saved->emb, saved->g1, saved->g2, saved->masks, saved->ld};
  }
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(field_mlp_fwd_kernel), FieldSmem::total);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_field_mlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + tc::kRows - 1) / tc::kRows;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, 2 * sm_count()));
  static const int dbg = std::getenv("NRB_FIELD_FWD_DEBUG") ? std::atoi(std::getenv("NRB_FIELD_FWD_DEBUG")) : 0;
  field_mlp_fwd_kernel<<<grid, tc::kRows, FieldSmem::total, static_cast<cudaStream_t>(stream)>>>(
      to_params(p), sv, x, sh, samples_per_ray, M, feature, sdf, alpha, dbg);
  return finish_launch("nrb_field_mlp_fwd");
}

extern "C" int nrb_field_mlp_bwd(const nrb_field_mlp_t* p, const nrb_field_bwd_in_t* in, const nrb_field_bwd_out_t* out,
                                 int32_t samples_per_ray, int64_t M, nrb_stream_t stream) {
  NRB_REQUIRE(p && in && out && M >= 0 && samples_per_ray > 0, NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: null pointer or bad size");
  NRB_REQUIRE(in->x && in->saved.h1 && in->saved.emb && in->saved.g1 && in->saved.g2 && in->saved.masks && in->sh &&
                  in->sdf && in->alpha && in->dfeature,
              NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: a required input is null");
  NRB_REQUIRE(in->saved.ld >= M && in->saved.ld % tc::kRows == 0, NRB_ERR_BAD_ARG,
              "nrb_field_mlp_bwd: saved.ld must be nrb_field_saved_ld(M)");
  for (int l = 0; l < 5; ++l) NRB_REQUIRE(p->weights[l] != nullptr, NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: weights[%d] is null", l);
  NRB_REQUIRE(p->beta != nullptr, NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: beta is null");
  NRB_REQUIRE(aligned16(in->x) && aligned16(in->saved.h1) && aligned16(in->saved.emb) && aligned16(in->saved.g1) &&
                  aligned16(in->saved.g2) && aligned16(in->sh) && aligned16(in->dfeature) &&
                  (out->dx == nullptr || aligned16(out->dx)),
              NRB_ERR_ALIGNMENT, "nrb_field_mlp_bwd: arrays must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  FieldBwdIn bi{in->x,  in->saved.h1, in->saved.emb, in->saved.g1,  in->saved.g2, in->saved.masks, in->saved.ld,
                in->sh, in->sdf,      in->alpha,     in->dfeature, in->dsdf,     in->dalpha};
  FieldBwdOut bo;
  bo.dx = out->dx;
  for (int l = 0; l < 5; ++l) {
    bo.dw[l] = out->dweights[l];
    bo.db[l] = out->dbiases[l];
  }
  bo.dbeta = out->dbeta;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(field_mlp_bwd_kernel), FieldBwdSmem::total);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_field_mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + tc::kRows - 1) / tc::kRows;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, sm_count()));
  static const int dbg = std::getenv("NRB_FIELD_BWD_DEBUG") ? std::atoi(std::getenv("NRB_FIELD_BWD_DEBUG")) : 0;
  // 22 (default) = two tiles in flight per CTA; 2 / 4 = the column-split kernel with that many threads per row;
  // 1 = the one-thread-per-row kernel
  static const int split = std::getenv("NRB_FIELD_BWD_SPLIT") ? std::atoi(std::getenv("NRB_FIELD_BWD_SPLIT")) : 22;
  if (split == 22 && M < (int64_t{1} << 31)) {  // two tiles in flight per CTA
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(field_mlp_bwd_pair_kernel), FieldBwdPairSmem::total);
    NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_field_mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const unsigned pgrid = static_cast<unsigned>(std::min<int64_t>((tiles + 1) / 2, sm_count()));
    field_mlp_bwd_pair_kernel<<<pgrid, kPairThreads, FieldBwdPairSmem::total, static_cast<cudaStream_t>(stream)>>>(
        to_params(p), bi, bo, samples_per_ray, M, dbg);
    return finish_launch("nrb_field_mlp_bwd");
  }
  if ((split == 2 || split == 4) && M < (int64_t{1} << 31)) {
    auto kern = split == 2 ? field_mlp_bwd_split_kernel<2> : field_mlp_bwd_split_kernel<4>;
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), FieldBwd2Smem::total);
    NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_field_mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, tc::kRows * split + 32, FieldBwd2Smem::total, static_cast<cudaStream_t>(stream)>>>(to_params(p), bi, bo,
                                                                                                   samples_per_ray, M, dbg);
    return finish_launch("nrb_field_mlp_bwd");
  }
  field_mlp_bwd_kernel<<<grid, tc::kRows, FieldBwdSmem::total, static_cast<cudaStream_t>(stream)>>>(
      to_params(p), bi, bo, samples_per_ray, M, dbg);
  return finish_launch("nrb_field_mlp_bwd");
}

// Debug: copy the phase trace of the last column-split backward launch (see NRB_TRACE) to the host.
extern "C" int nrb_debug_bwd_trace(long long* host, int32_t n) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(host, nrb::g_bwd_trace, sizeof(long long) * std::min(n, 2048));
  return static_cast<int>(e);
}
