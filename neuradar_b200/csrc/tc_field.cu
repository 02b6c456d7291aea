// Fused NeuRAD field MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Semantics: NeuRADField.forward after the hash grid (nerfstudio/fields/neurad_field.py:132-152):
//   geo   = mlp_geo(x)                      32 -> 32 (ReLU) -> 33         (field_components/mlp.py:159-178)
//   sdf, emb = split(geo, [1, 32])
//   feat  = emb + mlp_feature([emb, sh])    48 -> 32 (ReLU) -> 32 (ReLU) -> 32
//   alpha = sigmoid(-sdf * (|beta| + 1e-4))                               (model_components/utils.py:30-41)
// One CTA = 128 threads = one 128-sample tile at a time (persistent over tiles); thread t owns row t: it stages its
// row of the next layer's A operand (hi/lo tf32 halves) in shared memory, one elected thread issues the layer's
// tcgen05.mma chain, and every thread reads its accumulator row back with tcgen05.ld for bias + ReLU.  Weights are
// staged once per CTA.  Two CTAs per SM overlap one tile's MMA/TMEM latency with the other's epilogue.
#include "tc_common.cuh"

namespace nrb {

using namespace tc;

struct FieldParams {
  const float* w[5];  // mlp_geo.layers.{0,1}.weight, mlp_feature.layers.{0,1,2}.weight   ([out, in] row-major)
  const float* b[5];
  const float* beta;  // sdf_to_density.beta [1]
  float beta_min;
};

struct FieldSaved {  // activations kept for the backward pass (all optional), row-major
  float* h1;   // [M,32] post-ReLU hidden of mlp_geo
  float* emb;  // [M,32] geometry embedding (pre-activation output columns 1..32 of mlp_geo)
  float* g1;   // [M,32] post-ReLU hidden 1 of mlp_feature
  float* g2;   // [M,32] post-ReLU hidden 2 of mlp_feature
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
  __host__ __device__ static constexpr int w_hi(int l) { return field_w_hi(l); }
  __host__ __device__ static constexpr int w_lo(int l) { return field_w_lo(l); }
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

__global__ void __launch_bounds__(kRows, 2) field_mlp_fwd_kernel(const __grid_constant__ FieldParams prm,
                                                                 const __grid_constant__ FieldSaved sv,
                                                                 const float* __restrict__ x,   // [M,32]
                                                                 const float* __restrict__ sh,  // [N_rays,16]
                                                                 int samples_per_ray, int64_t M,
                                                                 float* __restrict__ feature,  // [M,32]
                                                                 float* __restrict__ sdf,      // [M]
                                                                 float* __restrict__ alpha) {  // [M]
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5;
  float* s_bias = reinterpret_cast<float*>(smem + FieldSmem::bias);
  char* a_hi = smem + FieldSmem::a_hi;
  char* a_lo = smem + FieldSmem::a_lo;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + FieldSmem::mbar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FieldSmem::tmem);

  for (int l = 0; l < 5; ++l) {
    stage_weight_split(prm.w[l], kOut[l], kN[l], kK[l], smem + FieldSmem::w_hi(l), smem + FieldSmem::w_lo(l));
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
  uint32_t phase = 0;

  // one layer: operands are staged, every thread has fenced; issue, wait, read the accumulator row
  auto run_layer = [&](int l) {
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (t == 0) {
      fence_after_sync();
      issue_gemm_kmajor(tmem_base, a_hi_u, a_lo_u, smem_u32(smem + FieldSmem::w_hi(l)), smem_u32(smem + FieldSmem::w_lo(l)),
                        kK[l], kN[l], false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
  };

  const int64_t tiles = (M + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row = tile * kRows + t;
    const bool ok = row < M;
    const int64_t rr = ok ? row : (M - 1);  // clamp: out-of-range rows compute on a valid row and are not stored
    float v[32];
    // ---- mlp_geo layer 0: 32 -> 32, ReLU
    load_row32(x + rr * 32, v);
    store_row_split<32>(a_hi, a_lo, t, v);
    run_layer(0);
    tmem_load_row<32>(tmem_base, warp, 0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + s_bias[j], 0.0f);
    if (ok && sv.h1 != nullptr) store_row32(sv.h1 + row * 32, v);
    store_row_split<32>(a_hi, a_lo, t, v);
    // ---- mlp_geo layer 1: 32 -> 33 (sdf | embedding), no activation
    run_layer(1);
    float geo[48];
    tmem_load_row<48>(tmem_base, warp, 0, geo);
    const float sdf_v = geo[0] + s_bias[48 + 0];
    float emb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) emb[j] = geo[1 + j] + s_bias[48 + 1 + j];
    if (ok && sv.emb != nullptr) store_row32(sv.emb + row * 32, emb);
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
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + s_bias[2 * 48 + j], 0.0f);
    if (ok && sv.g1 != nullptr) store_row32(sv.g1 + row * 32, v);
    store_row_split<32>(a_hi, a_lo, t, v);
    // ---- mlp_feature layer 1: 32 -> 32, ReLU
    run_layer(3);
    tmem_load_row<32>(tmem_base, warp, 0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + s_bias[3 * 48 + j], 0.0f);
    if (ok && sv.g2 != nullptr) store_row32(sv.g2 + row * 32, v);
    store_row_split<32>(a_hi, a_lo, t, v);
    // ---- mlp_feature layer 2: 32 -> 32, no activation; residual with the embedding
    run_layer(4);
    tmem_load_row<32>(tmem_base, warp, 0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = emb[j] + (v[j] + s_bias[4 * 48 + j]);
    if (ok) {
      store_row32(feature + row * 32, v);
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
// the tensor cores from the SAME shared-memory tiles:
//     dW_l (+)= delta_l^T * in_l        M = 64 (rows >= out_l ignored), N = in_l, reduction over the 128 samples,
//                                        accumulated in TMEM across ALL tiles of the CTA and flushed once at the end;
//     dIn_l  = delta_l * W_l            M = 128, N = in_l, reduction over out_l (W_l read MN-major as staged).
// Bias gradients are column sums of delta_l (31-shuffle transpose-reduce per warp) kept in shared memory.
struct FieldBwdIn {
  const float* x;     // [M,32] hash features (input of the forward pass)
  const float* h1;    // saved activations, [M,32] each
  const float* emb;
  const float* g1;
  const float* g2;
  const float* sh;    // [N_rays,16]
  const float* sdf;   // [M]
  const float* alpha; // [M]
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
  static constexpr int s_raw = d_lo + kRows * 48 * 4;   // layer input rows, [128 x 48] raw fp32 (transpose source)
  static constexpr int dt_hi = s_raw + kRows * 48 * 4;  // delta^T, [48 x 128] (A of the weight-gradient GEMM)
  static constexpr int dt_lo = dt_hi + 48 * kRows * 4;
  static constexpr int at_hi = dt_lo + 48 * kRows * 4;  // input^T, [48 x 128] (B of the weight-gradient GEMM)
  static constexpr int at_lo = at_hi + 48 * kRows * 4;
  static constexpr int dbacc = at_lo + 48 * kRows * 4;  // [4 warps][5 layers][48]
  static constexpr int mbar = dbacc + 4 * 5 * 48 * 4;
  static constexpr int tmem = mbar + 8;
  static constexpr int total = tmem + 8;
};

// TMEM lane that holds row r of an M = 64 accumulator (cta_group::1): 16 rows per 32-lane quarter (measured with
// tools_tc_probe.py: lane (r / 16) * 32 + r % 16).
__device__ __forceinline__ int m64_row_of_lane(int lane128) {
  const int q = lane128 >> 5, i = lane128 & 31;
  return i < 16 ? q * 16 + i : -1;
}

__global__ void __launch_bounds__(kRows, 1) field_mlp_bwd_kernel(const __grid_constant__ FieldParams prm,
                                                                 const __grid_constant__ FieldBwdIn in,
                                                                 const __grid_constant__ FieldBwdOut out,
                                                                 int samples_per_ray, int64_t M) {
  extern __shared__ __align__(128) char smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  char* d_hi = smem + FieldBwdSmem::d_hi;
  char* d_lo = smem + FieldBwdSmem::d_lo;
  char* s_raw = smem + FieldBwdSmem::s_raw;
  char* dt_hi = smem + FieldBwdSmem::dt_hi;
  char* dt_lo = smem + FieldBwdSmem::dt_lo;
  char* at_hi = smem + FieldBwdSmem::at_hi;
  char* at_lo = smem + FieldBwdSmem::at_lo;
  float* dbacc = reinterpret_cast<float*>(smem + FieldBwdSmem::dbacc);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + FieldBwdSmem::mbar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FieldBwdSmem::tmem);

  for (int l = 0; l < 5; ++l)
    stage_weight_transposed_split(prm.w[l], kOut[l], kN[l], kK[l], smem + field_w_hi(l), smem + field_w_lo(l));
  for (int i = t; i < 4 * 5 * 48; i += kRows) dbacc[i] = 0.0f;
  if (warp == 0) tmem_alloc<kBwdTmemCols>(tmem_slot);
  if (t == 0) mbar_init(mbar, 1);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const float beta = fabsf(__ldg(prm.beta)) + prm.beta_min;
  uint32_t phase = 0;
  bool first = true;
  float dbeta_acc = 0.0f;
  float* my_db = dbacc + warp * 5 * 48;

  // delta rows are in d_hi/d_lo (dcols columns), input rows in s_raw (acols columns): transpose both, then issue
  //   dW_l (+)= delta^T in   (M = 64, N = acols, reduction over the 128 samples)   -> TMEM column kDwCol[l]
  //   dIn   = delta W_l      (M = 128, N = acols, reduction over kred delta columns) -> TMEM column 0
  auto run_layer = [&](int l, int dcols, int acols, int kred) {
    __syncthreads();
    transpose_tile<false>(d_hi, d_lo, dcols, dt_hi, dt_lo);
    transpose_tile<true>(s_raw, nullptr, acols, at_hi, at_lo);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (t == 0) {
      fence_after_sync();
      issue_gemm(tmem_base + kDwCol[l], 64, acols, smem_u32(dt_hi), smem_u32(dt_lo), kRows, smem_u32(at_hi),
                 smem_u32(at_lo), kRows, kRows, !first);
      issue_gemm(tmem_base, 128, acols, smem_u32(d_hi), smem_u32(d_lo), dcols, smem_u32(smem + field_w_hi(l)),
                 smem_u32(smem + field_w_lo(l)), kN[l], kred, false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
  };

  const int64_t tiles = (M + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row = tile * kRows + t;
    const bool ok = row < M;
    const int64_t rr = ok ? row : (M - 1);
    float delta[32], act[32], tmp[32];
    // ---- layer 4 (mlp_feature.layers.2): delta = d feature
    load_row32(in.dfeature + rr * 32, delta);
    if (!ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) delta[j] = 0.0f;
    }
    float demb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) demb[j] = delta[j];  // residual branch
    load_row32(in.g2 + rr * 32, act);
    store_row_split<32>(d_hi, d_lo, t, delta);
    store_row_raw<32>(s_raw, t, act);
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    my_db[4 * 48 + lane] += warp_column_sums(tmp, lane);
    run_layer(4, 32, 32, 32);
    tmem_load_row<32>(tmem_base, warp, 0, delta);
#pragma unroll
    for (int j = 0; j < 32; ++j) delta[j] = act[j] > 0.0f ? delta[j] : 0.0f;
    // ---- layer 3 (mlp_feature.layers.1)
    load_row32(in.g1 + rr * 32, act);
    store_row_split<32>(d_hi, d_lo, t, delta);
    store_row_raw<32>(s_raw, t, act);
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    my_db[3 * 48 + lane] += warp_column_sums(tmp, lane);
    run_layer(3, 32, 32, 32);
    tmem_load_row<32>(tmem_base, warp, 0, delta);
#pragma unroll
    for (int j = 0; j < 32; ++j) delta[j] = act[j] > 0.0f ? delta[j] : 0.0f;
    // ---- layer 2 (mlp_feature.layers.0): input [emb | sh]
    {
      float in48[48];
      load_row32(in.emb + rr * 32, act);
#pragma unroll
      for (int j = 0; j < 32; ++j) in48[j] = act[j];
      const float4* shp = reinterpret_cast<const float4*>(in.sh + (rr / samples_per_ray) * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 q = __ldg(shp + c);
        in48[32 + 4 * c] = q.x;
        in48[32 + 4 * c + 1] = q.y;
        in48[32 + 4 * c + 2] = q.z;
        in48[32 + 4 * c + 3] = q.w;
      }
      store_row_split<32>(d_hi, d_lo, t, delta);
      store_row_raw<48>(s_raw, t, in48);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    my_db[2 * 48 + lane] += warp_column_sums(tmp, lane);
    run_layer(2, 32, 48, 32);
    {
      float din[48];
      tmem_load_row<48>(tmem_base, warp, 0, din);
#pragma unroll
      for (int j = 0; j < 32; ++j) demb[j] += din[j];  // the SH part of the input carries no gradient
    }
    // ---- layer 1 (mlp_geo.layers.1): delta = [d sdf | d emb], padded to 48 columns
    float dsdf_v = 0.0f;
    if (ok) {
      const float a = __ldg(in.alpha + row), sd = __ldg(in.sdf + row);
      const float da = in.dalpha != nullptr ? __ldg(in.dalpha + row) : 0.0f;
      const float s = a * (1.0f - a);
      dsdf_v = (in.dsdf != nullptr ? __ldg(in.dsdf + row) : 0.0f) - da * beta * s;
      dbeta_acc -= da * sd * s;
    }
    {
      float d48[48];
      d48[0] = dsdf_v;
#pragma unroll
      for (int j = 0; j < 32; ++j) d48[1 + j] = demb[j];
#pragma unroll
      for (int j = 33; j < 48; ++j) d48[j] = 0.0f;
      load_row32(in.h1 + rr * 32, act);
      store_row_split<48>(d_hi, d_lo, t, d48);
      store_row_raw<32>(s_raw, t, act);
    }
    {
      const float s0 = warp_sum(dsdf_v);
      if (lane == 0) my_db[1 * 48 + 0] += s0;
#pragma unroll
      for (int j = 0; j < 32; ++j) tmp[j] = demb[j];
      my_db[1 * 48 + 1 + lane] += warp_column_sums(tmp, lane);
    }
    run_layer(1, 48, 32, 40);
    tmem_load_row<32>(tmem_base, warp, 0, delta);
#pragma unroll
    for (int j = 0; j < 32; ++j) delta[j] = act[j] > 0.0f ? delta[j] : 0.0f;
    // ---- layer 0 (mlp_geo.layers.0)
    load_row32(in.x + rr * 32, act);
    store_row_split<32>(d_hi, d_lo, t, delta);
    store_row_raw<32>(s_raw, t, act);
#pragma unroll
    for (int j = 0; j < 32; ++j) tmp[j] = delta[j];
    my_db[0 * 48 + lane] += warp_column_sums(tmp, lane);
    run_layer(0, 32, 32, 32);
    if (out.dx != nullptr) {
      tmem_load_row<32>(tmem_base, warp, 0, delta);
      if (ok) store_row32(out.dx + row * 32, delta);
    }
    first = false;
  }
  // ---- flush: weight gradients from TMEM, bias gradients from shared memory, beta
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
                                                          int64_t M, float* __restrict__ y) {
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
    if (K == 32) {
      float v32[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v32[j] = v[j];
      store_row_split<32>(a_hi, a_lo, t, v32);
    } else {
      store_row_split<48>(a_hi, a_lo, t, v);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (t == 0) {
      fence_after_sync();
      issue_gemm_kmajor(tmem_base, smem_u32(a_hi), smem_u32(a_lo), smem_u32(w_hi), smem_u32(w_lo), K, N, false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    fence_after_sync();
    float acc[48];
    tmem_load_row<48>(tmem_base, warp, 0, acc);
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
  tc_linear_kernel<<<grid, tc::kRows, smem, static_cast<cudaStream_t>(stream)>>>(x, w, b, K, n_out, relu, M, y);
  return finish_launch("nrb_tc_linear");
}

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
  FieldParams prm;
  for (int l = 0; l < 5; ++l) {
    prm.w[l] = p->weights[l];
    prm.b[l] = p->biases[l];
  }
  prm.beta = p->beta;
  prm.beta_min = p->beta_min;
  FieldSaved sv{nullptr, nullptr, nullptr, nullptr};
  if (saved != nullptr) {
    sv.h1 = saved->h1;
    sv.emb = saved->emb;
    sv.g1 = saved->g1;
    sv.g2 = saved->g2;
  }
  cudaError_t e = cudaFuncSetAttribute(field_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FieldSmem::total);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_field_mlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + tc::kRows - 1) / tc::kRows;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, 2 * sm_count()));
  field_mlp_fwd_kernel<<<grid, tc::kRows, FieldSmem::total, static_cast<cudaStream_t>(stream)>>>(
      prm, sv, x, sh, samples_per_ray, M, feature, sdf, alpha);
  return finish_launch("nrb_field_mlp_fwd");
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

extern "C" int nrb_field_mlp_bwd(const nrb_field_mlp_t* p, const nrb_field_bwd_in_t* in, const nrb_field_bwd_out_t* out,
                                 int32_t samples_per_ray, int64_t M, nrb_stream_t stream) {
  NRB_REQUIRE(p && in && out && M >= 0 && samples_per_ray > 0, NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: null pointer or bad size");
  NRB_REQUIRE(in->x && in->h1 && in->emb && in->g1 && in->g2 && in->sh && in->sdf && in->alpha && in->dfeature,
              NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: a required input is null");
  for (int l = 0; l < 5; ++l) NRB_REQUIRE(p->weights[l] != nullptr, NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: weights[%d] is null", l);
  NRB_REQUIRE(p->beta != nullptr, NRB_ERR_BAD_ARG, "nrb_field_mlp_bwd: beta is null");
  NRB_REQUIRE(aligned16(in->x) && aligned16(in->h1) && aligned16(in->emb) && aligned16(in->g1) && aligned16(in->g2) &&
                  aligned16(in->sh) && aligned16(in->dfeature) && (out->dx == nullptr || aligned16(out->dx)),
              NRB_ERR_ALIGNMENT, "nrb_field_mlp_bwd: row-major [M,32] arrays must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  FieldParams prm;
  for (int l = 0; l < 5; ++l) {
    prm.w[l] = p->weights[l];
    prm.b[l] = p->biases[l];
  }
  prm.beta = p->beta;
  prm.beta_min = p->beta_min;
  FieldBwdIn bi{in->x, in->h1, in->emb, in->g1, in->g2, in->sh, in->sdf, in->alpha, in->dfeature, in->dsdf, in->dalpha};
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
  field_mlp_bwd_kernel<<<grid, tc::kRows, FieldBwdSmem::total, static_cast<cudaStream_t>(stream)>>>(prm, bi, bo, samples_per_ray, M);
  return finish_launch("nrb_field_mlp_bwd");
}
