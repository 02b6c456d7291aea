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
