// Ray samplers: PowerSampler initial bins and warp-per-ray inverse-CDF importance sampling.
// Semantics: SpacedSampler / PDFSampler.generate_ray_samples (nerfstudio/model_components/ray_samplers.py:80-132,
// 280-376).  All randomness is drawn by the host with torch.rand and passed in; the linspace tables are passed in
// as well, so every value that feeds torch.searchsorted is formed with the reference's own fp32 roundings.
#include "common.cuh"

namespace nrb {

__global__ void __launch_bounds__(256) spaced_bins_kernel(const float* __restrict__ nears,
                                                          const float* __restrict__ fars, nrb_spacing_t sp,
                                                          const float* __restrict__ base_bins,
                                                          const float* __restrict__ jitter, int jitter_per_bin, int S,
                                                          float* __restrict__ sbins, float* __restrict__ ebins,
                                                          int64_t total) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int nb = S + 1;
  const int64_t n = gid / nb;
  const int j = static_cast<int>(gid - n * nb);
  float b = base_bins[j];
  if (jitter != nullptr) {
    // bin_centers / bin_upper / bin_lower of ray_samplers.py:112-115
    const float upper = (j < S) ? mul(add(base_bins[j + 1], base_bins[j]), 0.5f) : base_bins[S];
    const float lower = (j > 0) ? mul(add(base_bins[j], base_bins[j - 1]), 0.5f) : base_bins[0];
    const float t = jitter_per_bin ? jitter[gid] : jitter[n];
    b = add(lower, mul(sub(upper, lower), t));
  }
  const float s_near = power_fn(mul(nears[n], sp.scaling), sp.lambda);
  const float s_far = power_fn(mul(fars[n], sp.scaling), sp.lambda);
  sbins[gid] = b;
  ebins[gid] = spacing_to_euclidean(b, s_near, s_far, sp);
}

// One warp per ray.  The CDF (S_in+1 values) and the existing bin edges live in shared memory; each lane then
// resolves its own query points with a binary search, which is exactly torch.searchsorted(side="right").
constexpr int kPdfWarps = 4;

__global__ void __launch_bounds__(kPdfWarps * 32) pdf_sample_kernel(
    const float* __restrict__ nears, const float* __restrict__ fars, nrb_spacing_t sp,
    const float* __restrict__ weights, const float* __restrict__ sbins_in, int S_in,
    const float* __restrict__ u_base, const float* __restrict__ jitter, int S_out, float hist_pad, float eps,
    float* __restrict__ sbins_out, float* __restrict__ ebins_out, int64_t* __restrict__ inds_out,
    float* __restrict__ cdf_out, int64_t N) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kPdfWarps + warp;
  if (n >= N) return;
  const int nin = S_in + 1;
  float* cdf = smem + warp * 2 * nin;
  float* bins = cdf + nin;

  // weights + histogram padding, total, zero-weight guard (ray_samplers.py:308-316)
  const float* wrow = weights + n * S_in;
  double part = 0.0;
  for (int i = lane; i < S_in; i += 32) part += static_cast<double>(add(wrow[i], hist_pad));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
  float total = static_cast<float>(part);
  const float pad = fmaxf(sub(eps, total), 0.0f);
  const float pad_each = div(pad, static_cast<float>(S_in));
  total = add(total, pad);

  // cdf = [0, min(1, cumsum(pdf))]; torch's CPU cumsum accumulates in double and rounds each prefix to fp32
  double carry = 0.0;
  for (int base = 0; base < S_in; base += 32) {
    const int i = base + lane;
    const float pdf = (i < S_in) ? div(add(add(wrow[i], hist_pad), pad_each), total) : 0.0f;
    const double incl = carry + warp_inclusive_sum(static_cast<double>(pdf), lane);
    if (i < S_in) cdf[i + 1] = fminf(1.0f, static_cast<float>(incl));
    carry = __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) cdf[0] = 0.0f;
  const float* brow = sbins_in + n * nin;
  for (int i = lane; i < nin; i += 32) bins[i] = brow[i];
  __syncwarp();
  if (cdf_out != nullptr)
    for (int i = lane; i < nin; i += 32) cdf_out[n * nin + i] = cdf[i];

  const int nb = S_out + 1;
  const float s_near = power_fn(mul(nears[n], sp.scaling), sp.lambda);
  const float s_far = power_fn(mul(fars[n], sp.scaling), sp.lambda);
  // u = linspace + rand/nb (training) or + 1/(2 nb) (eval), ray_samplers.py:321-335
  const float shift = (jitter != nullptr) ? div(jitter[n], static_cast<float>(nb))
                                          : static_cast<float>(1.0 / static_cast<double>(2 * nb));
  for (int j = lane; j < nb; j += 32) {
    const float u = add(u_base[j], shift);
    int lo = 0, hi = nin;  // first index with cdf[idx] > u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    const int below = min(max(lo - 1, 0), S_in);
    const int above = min(max(lo, 0), S_in);
    const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float t = div(sub(u, c0), sub(c1, c0));
    t = isnan(t) ? 0.0f : t;  // nan_to_num(nan=0); +-inf are clipped below
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float b = add(b0, mul(t, sub(b1, b0)));
    const int64_t o = n * nb + j;
    sbins_out[o] = b;
    ebins_out[o] = spacing_to_euclidean(b, s_near, s_far, sp);
    if (inds_out != nullptr) inds_out[o] = lo;
  }
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_spaced_bins(const nrb_rays_t* rays, nrb_spacing_t spacing, const float* base_bins,
                               const float* jitter, int32_t jitter_per_bin, int32_t S, float* sbins, float* ebins,
                               nrb_stream_t stream) {
  NRB_REQUIRE(rays && rays->nears && rays->fars && rays->num_rays >= 0, NRB_ERR_BAD_ARG,
              "nrb_spaced_bins: rays.nears/fars must be set");
  NRB_REQUIRE(base_bins && sbins && ebins && S > 0, NRB_ERR_BAD_ARG, "nrb_spaced_bins: null pointer or S <= 0");
  NRB_REQUIRE(spacing.lambda != 0.f && spacing.lambda != 1.f && spacing.scaling > 0.f, NRB_ERR_UNSUPPORTED,
              "nrb_spaced_bins: lambda must not be 0 or 1 and scaling must be positive");
  if (rays->num_rays == 0) return NRB_OK;
  const int64_t total = rays->num_rays * (S + 1);
  spaced_bins_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rays->nears, rays->fars, spacing, base_bins, jitter, jitter_per_bin, S, sbins, ebins, total);
  return finish_launch("nrb_spaced_bins");
}

extern "C" int nrb_pdf_sample(const nrb_rays_t* rays, nrb_spacing_t spacing, const float* weights,
                              const float* sbins_in, int32_t S_in, const float* u_base, const float* jitter,
                              int32_t S_out, float histogram_padding, float eps, float* sbins_out, float* ebins_out,
                              int64_t* inds, float* cdf, nrb_stream_t stream) {
  NRB_REQUIRE(rays && rays->nears && rays->fars && rays->num_rays >= 0, NRB_ERR_BAD_ARG,
              "nrb_pdf_sample: rays.nears/fars must be set");
  NRB_REQUIRE(weights && sbins_in && u_base && sbins_out && ebins_out, NRB_ERR_BAD_ARG, "nrb_pdf_sample: null pointer");
  NRB_REQUIRE(S_in > 0 && S_in <= NRB_MAX_SAMPLES && S_out > 0 && S_out <= NRB_MAX_SAMPLES, NRB_ERR_BAD_ARG,
              "nrb_pdf_sample: sample counts must be in [1,%d]", NRB_MAX_SAMPLES);
  NRB_REQUIRE(spacing.lambda != 0.f && spacing.lambda != 1.f && spacing.scaling > 0.f, NRB_ERR_UNSUPPORTED,
              "nrb_pdf_sample: lambda must not be 0 or 1 and scaling must be positive");
  if (rays->num_rays == 0) return NRB_OK;
  const size_t smem = sizeof(float) * kPdfWarps * 2 * (S_in + 1);
  pdf_sample_kernel<<<blocks_for(rays->num_rays, kPdfWarps), kPdfWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      rays->nears, rays->fars, spacing, weights, sbins_in, S_in, u_base, jitter, S_out, histogram_padding, eps,
      sbins_out, ebins_out, inds, cdf, rays->num_rays);
  return finish_launch("nrb_pdf_sample");
}
