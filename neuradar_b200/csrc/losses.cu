// Per-ray losses that run every training step on the path's outputs: the MipNeRF-360
// distortion loss and the ZipNeRF anti-aliased interlevel loss.  One warp per ray; all per-ray state lives in shared
// memory; each kernel returns the per-ray loss and the factor d loss_ray / d w, so the backward pass is one scale.
// Semantics: nerfstudio/model_components/losses.py:137-156 (lossfun_distortion, distortion_loss) and :620-705
// (_blur_stepfun, _sorted_interp_quad, zipnerf_interlevel_loss).
#include "common.cuh"

namespace nrb {

constexpr int kLossWarps = 4;

// loss_ray = sum_ij w_i w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1} - t_i) / 3, u = bin midpoints
// grad_i   = 2 sum_j w_j |u_i - u_j| + 2 w_i (t_{i+1} - t_i) / 3
__global__ void __launch_bounds__(kLossWarps * 32) distortion_kernel(const float* __restrict__ sbins, int64_t bin_stride,
                                                                     const float* __restrict__ w, int64_t N, int S,
                                                                     float* __restrict__ loss, float* __restrict__ grad) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kLossWarps + warp;
  if (n >= N) return;
  float* u = smem + warp * 2 * S;
  float* ww = u + S;
  const float* t = sbins + n * bin_stride;
  for (int i = lane; i < S; i += 32) {
    u[i] = (t[i + 1] + t[i]) * 0.5f;
    ww[i] = w[n * S + i];
  }
  __syncwarp();
  float acc = 0.0f;
  for (int i = lane; i < S; i += 32) {
    const float ui = u[i], wi = ww[i];
    float inner = 0.0f;
    for (int j = 0; j < S; ++j) inner = fmaf(ww[j], fabsf(ui - u[j]), inner);
    const float width = t[i + 1] - t[i];
    acc += wi * inner + wi * wi * width * (1.0f / 3.0f);
    if (grad != nullptr) grad[n * S + i] = 2.0f * inner + 2.0f * wi * width * (1.0f / 3.0f);
  }
  acc = warp_sum(acc);
  if (lane == 0) loss[n] = acc;
}

// Inclusive scan over a shared array of `len` floats by one warp (in place).  The running sum is carried in fp64 and
// rounded to fp32 per element, which is what torch.cumsum does on the host (at::acc_type<float, false> is double):
// the blurred histogram is a difference of large cancelling prefix sums, and an fp32 carry moves the loss by percents.
__device__ __forceinline__ void warp_scan_inplace(float* a, int len, int lane) {
  double carry = 0.0;
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    const double v = (i < len) ? static_cast<double>(a[i]) : 0.0;
    const double incl = carry + warp_inclusive_sum(v, lane);
    if (i < len) a[i] = static_cast<float>(incl);
    carry = __shfl_sync(kFull, incl, 31);
  }
  __syncwarp();
}

// ZipNeRF interlevel loss of ONE proposal round against the final level (c, w constants).
//   c [N, Sc+1], w [N, Sc]: final-level spacing bins / weights; cp [N, Sp+1], wp [N, Sp]: the proposal's.
//   loss_ray = sum_q relu(ws_q - wp_q)^2 / (wp_q + 1e-5), ws = diff of the blurred final-level CDF at cp.
__global__ void __launch_bounds__(kLossWarps * 32) interlevel_kernel(
    const float* __restrict__ c_bins, int64_t c_stride, const float* __restrict__ w, int Sc,
    const float* __restrict__ cp_bins, int64_t cp_stride, const float* __restrict__ wp, int Sp, float r, int64_t N,
    float* __restrict__ loss, float* __restrict__ grad) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kLossWarps + warp;
  if (n >= N) return;
  const int n1 = Sc + 1;      // step-function knots
  const int m = 2 * n1;       // knots of the blurred (piecewise linear) function
  const int len = m + 2;      // after padding with (0, 0) and (1, 0)
  // per-warp arrays: cs[n1] y1[n1] | xp[len] pdf[len] cdf[len] | s2[m] | q[Sp+1]
  float* cs = smem + warp * (2 * n1 + 3 * len + m + (Sp + 1));
  float* y1 = cs + n1;
  float* xp = y1 + n1;
  float* pdf = xp + len;
  float* cdf = pdf + len;
  float* s2 = cdf + len;
  float* qv = s2 + m;
  const float* c = c_bins + n * c_stride;
  const float* wr = w + n * Sc;
  // final-level weights with the remaining accumulation put on the last sample, normalised by the bin widths
  float part = 0.0f;
  for (int i = lane; i < Sc; i += 32) part += wr[i];
  const float accum = warp_sum(part);
  for (int i = lane; i < n1; i += 32) cs[i] = c[i];
  __syncwarp();
  // y1_k = (y_k - y_{k-1}) / (2r) with y_{-1} = y_{Sc} = 0, y = w_norm
  for (int k = lane; k < n1; k += 32) {
    float yk = 0.0f, ykm = 0.0f;
    if (k < Sc) {
      const float wk = wr[k] + ((k == Sc - 1) ? (1.0f - accum) : 0.0f);
      yk = wk / (cs[k + 1] - cs[k]);
    }
    if (k > 0) {
      const float wk = wr[k - 1] + ((k - 1 == Sc - 1) ? (1.0f - accum) : 0.0f);
      ykm = wk / (cs[k] - cs[k - 1]);
    }
    y1[k] = (yk - ykm) / (2.0f * r);
  }
  __syncwarp();
  // merge the two sorted knot sequences a_k = c_k - r (slope change +y1_k) and b_k = c_k + r (slope change -y1_k);
  // equal keys keep the order of the concatenation [a | b]
  for (int k = lane; k < n1; k += 32) {
    const float ak = cs[k] - r, bk = cs[k] + r;
    int lo = 0, hi = n1;  // number of b_j < a_k
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cs[mid] + r < ak) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    const int pa = k + lo;
    lo = 0, hi = n1;  // number of a_j <= b_k
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cs[mid] - r <= bk) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    const int pb = k + lo;
    xp[1 + pa] = ak;
    s2[pa] = y1[k];
    xp[1 + pb] = bk;
    s2[pb] = -y1[k];
  }
  __syncwarp();
  // slope = cumsum(y2) (the last knot's jump is not needed), pdf = [0, clamp(cumsum(width * slope), 0)]
  warp_scan_inplace(s2, m - 1, lane);
  for (int k = lane; k < m - 1; k += 32) pdf[2 + k] = (xp[2 + k] - xp[1 + k]) * s2[k];
  __syncwarp();
  warp_scan_inplace(pdf + 2, m - 1, lane);
  for (int k = lane; k < m - 1; k += 32) pdf[2 + k] = fmaxf(pdf[2 + k], 0.0f);
  if (lane == 0) {
    pdf[1] = 0.0f;        // first knot of the blurred function
    pdf[0] = 0.0f;        // padding (0, 0)
    pdf[len - 1] = 0.0f;  // padding (1, 0)
    xp[0] = 0.0f;
    xp[len - 1] = 1.0f;
  }
  __syncwarp();
  // cdf over the m blurred knots: [0, cumsum(trapezoids)], then padded with 0 in front and 1 behind
  for (int k = lane; k < m - 1; k += 32) cdf[2 + k] = 0.5f * (pdf[2 + k] + pdf[1 + k]) * (xp[2 + k] - xp[1 + k]);
  __syncwarp();
  warp_scan_inplace(cdf + 2, m - 1, lane);
  if (lane == 0) {
    cdf[0] = 0.0f;
    cdf[1] = 0.0f;
    cdf[len - 1] = 1.0f;
  }
  __syncwarp();
  // piecewise-quadratic interpolation of the cdf at the proposal's bin edges
  const float* cp = cp_bins + n * cp_stride;
  for (int q = lane; q <= Sp; q += 32) {
    const float x = cp[q];
    int lo = 0, hi = len;  // searchsorted(side="left"): first index with xp[idx] >= x
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (xp[mid] < x) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    const int left = max(lo - 1, 0), right = min(lo, len - 1);
    const float x0 = xp[left], x1 = xp[right], p0 = pdf[left], p1 = pdf[right];
    float off = (x - x0) / (x1 - x0);
    off = isnan(off) ? 0.0f : off;
    off = fminf(fmaxf(off, 0.0f), 1.0f);
    qv[q] = cdf[left] + (x - x0) * (p0 + p1 * off + p0 * (1.0f - off)) * 0.5f;
  }
  __syncwarp();
  float acc = 0.0f;
  for (int q = lane; q < Sp; q += 32) {
    const float ws = qv[q + 1] - qv[q];
    const float wq = wp[n * Sp + q];
    const float d = fmaxf(ws - wq, 0.0f), den = wq + 1e-5f;
    acc += d * d / den;
    if (grad != nullptr) grad[n * Sp + q] = -2.0f * d / den - d * d / (den * den);
  }
  acc = warp_sum(acc);
  if (lane == 0) loss[n] = acc;
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_distortion_loss(const float* sbins, int64_t bin_stride, const float* weights, int64_t N, int32_t S,
                                   float* loss_per_ray, float* grad_factor, nrb_stream_t stream) {
  NRB_REQUIRE(sbins && weights && loss_per_ray && N >= 0, NRB_ERR_BAD_ARG, "nrb_distortion_loss: null pointer");
  NRB_REQUIRE(S > 0 && S <= NRB_MAX_SAMPLES && bin_stride >= S + 1, NRB_ERR_BAD_ARG,
              "nrb_distortion_loss: S must be in [1,%d] and bin_stride >= S+1", NRB_MAX_SAMPLES);
  if (N == 0) return NRB_OK;
  const size_t smem = sizeof(float) * kLossWarps * 2 * S;
  distortion_kernel<<<blocks_for(N, kLossWarps), kLossWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      sbins, bin_stride, weights, N, S, loss_per_ray, grad_factor);
  return finish_launch("nrb_distortion_loss");
}

extern "C" int nrb_interlevel_loss(const float* c_bins, int64_t c_stride, const float* w, int32_t Sc, const float* cp_bins,
                                   int64_t cp_stride, const float* wp, int32_t Sp, float pulse_width, int64_t N,
                                   float* loss_per_ray, float* grad_factor, nrb_stream_t stream) {
  NRB_REQUIRE(c_bins && w && cp_bins && wp && loss_per_ray && N >= 0, NRB_ERR_BAD_ARG, "nrb_interlevel_loss: null pointer");
  NRB_REQUIRE(Sc > 0 && Sc <= NRB_MAX_SAMPLES && Sp > 0 && Sp <= NRB_MAX_SAMPLES && c_stride >= Sc + 1 && cp_stride >= Sp + 1,
              NRB_ERR_BAD_ARG, "nrb_interlevel_loss: sample counts must be in [1,%d] and strides cover the bins", NRB_MAX_SAMPLES);
  NRB_REQUIRE(pulse_width > 0.f, NRB_ERR_BAD_ARG, "nrb_interlevel_loss: pulse_width must be positive");
  if (N == 0) return NRB_OK;
  const int n1 = Sc + 1, m = 2 * n1, len = m + 2;
  const size_t smem = sizeof(float) * kLossWarps * (2 * n1 + 3 * len + m + (Sp + 1));
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(interlevel_kernel), 96 * 1024);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_interlevel_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  interlevel_kernel<<<blocks_for(N, kLossWarps), kLossWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      c_bins, c_stride, w, Sc, cp_bins, cp_stride, wp, Sp, pulse_width, N, loss_per_ray, grad_factor);
  return finish_launch("nrb_interlevel_loss");
}
