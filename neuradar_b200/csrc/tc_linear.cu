// One linear layer y = x W^T + b (optionally ReLU) on the 5th-generation tensor cores (tcgen05.mma as 3xTF32, the
// accumulator in tensor memory): the unit test of the descriptor, layout and split conventions of tc_common.cuh that the
// fused field kernels (field_fused.cu) are built from, and the tensor-core route of `functional.tc_linear`.
// One CTA = 128 threads = one 128-row tile at a time (persistent over tiles); thread t stages row t of the A operand as
// tf32 hi / lo halves in shared memory, one thread issues the MMA chain, every thread reads its accumulator row back
// with tcgen05.ld.  K in {32, 48}, N (padded to a multiple of 16) <= 48.
#include <algorithm>

#include "tc_common.cuh"

namespace nrb {

using namespace tc;

constexpr int kTmemCols = 256;

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
    if (warp == 0) {  // one elected lane of a converged warp issues (straight-line UTCHMMA sequence, see elect_one)
      if (elect_one()) {
        fence_after_sync();
        issue_gemm(tmem_base, 128, N, smem_u32(a_hi), smem_u32(a_lo), K, smem_u32(w_hi), smem_u32(w_lo), K, K, false);
        mma_commit(mbar);
      }
      __syncwarp();
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
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(tc_linear_kernel), static_cast<int>(smem));
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_tc_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + tc::kRows - 1) / tc::kRows;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, 2 * sm_count()));
  tc_linear_kernel<<<grid, tc::kRows, smem, static_cast<cudaStream_t>(stream)>>>(x, w, b, K, n_out, relu, M, y);
  return finish_launch("nrb_tc_linear");
}

