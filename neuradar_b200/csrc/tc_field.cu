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

constexpr int kTmemCols = 64;
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
  auto run_layer = [&](int l) {
    if (!(dbg & 8)) fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (dbg & 1) return;
    if (uwarp == 0) {
      if (elect_one()) {
        fence_after_sync();
        if (!(dbg & 2))
          issue_gemm(tmem_base, 128, kN[l], a_hi_u, a_lo_u, kK[l], smem_u32(smem + field_w_hi(l)),
                     smem_u32(smem + field_w_lo(l)), kK[l], kK[l], false);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
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
    store_row_split<32>(a_hi, a_lo, t, v);
    const int64_t ntile = tile + gridDim.x;
    if (ntile < tiles) warp_load_rows_coalesced(x, ntile * kRows + warp * 32, M, lane, xpf);
    run_layer(0);
    tmem_load_row<32>(tmem_base, warp, 0, v);
    const uint32_t m_h1 = relu_bias_mask(v, s_bias);
    if (train) store_col32(sv.h1, sv.ld, row, v);
    store_row_split<32>(a_hi, a_lo, t, v);
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
      store_row_split<48>(a_hi, a_lo, t, in48);
    }
    run_layer(2);
    tmem_load_row<32>(tmem_base, warp, 0, v);
    const uint32_t m_g1 = relu_bias_mask(v, s_bias + 2 * 48);
    if (train) store_col32(sv.g1, sv.ld, row, v);
    store_row_split<32>(a_hi, a_lo, t, v);
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
    store_row_split<32>(a_hi, a_lo, t, v);
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
  cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
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
  cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
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
  cudaError_t e = cudaFuncSetAttribute(field_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FieldSmem::total);
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
  cudaError_t e = cudaFuncSetAttribute(field_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FieldBwdSmem::total);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_field_mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + tc::kRows - 1) / tc::kRows;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, sm_count()));
  static const int dbg = std::getenv("NRB_FIELD_BWD_DEBUG") ? std::atoi(std::getenv("NRB_FIELD_BWD_DEBUG")) : 0;
  field_mlp_bwd_kernel<<<grid, tc::kRows, FieldBwdSmem::total, static_cast<cudaStream_t>(stream)>>>(
      to_params(p), bi, bo, samples_per_ray, M, dbg);
  return finish_launch("nrb_field_mlp_bwd");
}
