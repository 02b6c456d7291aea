// Compositing along rays: one warp per ray, samples striped over the lanes in chunks of 32, exclusive scans done
// with shuffles and a running carry, per-ray segments read with coalesced row accesses.
// Semantics: RaySamples.get_weights (nerfstudio/cameras/rays.py:188-210), the alpha-compositing tail of
// NeuRadarModel.get_nff_outputs (nerfstudio/models/neuradar.py:504-517) incl. nerfacc.render_weight_from_alpha's
// dense contract, FeatureRenderer / AccumulationRenderer (model_components/renderers.py:59-90,322-350) and
// render_depth_simple (models/neurad.py:721-728).
#include "common.cuh"

namespace nrb {

constexpr int kRayWarps = 4;
constexpr int kMaxChunks = NRB_MAX_SAMPLES / 32;

__global__ void __launch_bounds__(kRayWarps * 32) density_weights_fwd_kernel(const float* __restrict__ dens,
                                                                             nrb_intervals_t iv, int64_t N,
                                                                             float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kRayWarps + (threadIdx.x >> 5);
  if (n >= N) return;
  const int S = iv.num_samples;
  const float* st = iv.starts + n * iv.row_stride;
  const float* en = iv.ends + n * iv.row_stride;
  float carry = 0.0f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const bool ok = i < S;
    const float dd = ok ? mul(sub(en[i], st[i]), dens[n * S + i]) : 0.0f;
    const float incl = warp_inclusive_sum(dd, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.0f;
    const float alpha = 1.0f - expf(-dd);
    const float trans = expf(-(carry + excl));
    if (ok) weights[n * S + i] = nan_to_num(alpha * trans);
    carry += __shfl_sync(kFull, incl, 31);
  }
}

__global__ void __launch_bounds__(kRayWarps * 32) density_weights_bwd_kernel(const float* __restrict__ dens,
                                                                             nrb_intervals_t iv,
                                                                             const float* __restrict__ dweights,
                                                                             int64_t N, float* __restrict__ ddens) {
  const int lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kRayWarps + (threadIdx.x >> 5);
  if (n >= N) return;
  const int S = iv.num_samples;
  const float* st = iv.starts + n * iv.row_stride;
  const float* en = iv.ends + n * iv.row_stride;
  // forward recompute, keeping per-chunk values in registers
  float delta[kMaxChunks], dd[kMaxChunks], trans[kMaxChunks];
  float carry = 0.0f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int i = c * 32 + lane;
    if (c * 32 < S) {
      const bool ok = i < S;
      delta[c] = ok ? sub(en[i], st[i]) : 0.0f;
      dd[c] = ok ? mul(delta[c], dens[n * S + i]) : 0.0f;
      const float incl = warp_inclusive_sum(dd[c], lane);
      float excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = 0.0f;
      trans[c] = expf(-(carry + excl));
      carry += __shfl_sync(kFull, incl, 31);
    }
  }
  // w = (1 - e^-dd) T:   d/ddd_i = g_i T_i e^-dd_i - sum_{k>i} g_k w_k
  float suffix = 0.0f;  // sum over later chunks
#pragma unroll
  for (int c = kMaxChunks - 1; c >= 0; --c) {
    const int i = c * 32 + lane;
    if (c * 32 < S) {
      const bool ok = i < S;
      const float e = expf(-dd[c]);
      const float w = (1.0f - e) * trans[c];
      float g = ok ? dweights[n * S + i] : 0.0f;
      if (isnan(w) || isinf(w)) g = 0.0f;  // nan_to_num passes no gradient at non-finite values
      float gw = ok ? g * w : 0.0f;
      // inclusive suffix scan over lanes
      float incl = gw;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_down_sync(kFull, incl, o);
        if (lane + o < 32) incl += v;
      }
      float excl = __shfl_down_sync(kFull, incl, 1);
      if (lane == 31) excl = 0.0f;
      const float later = suffix + excl;
      if (ok) ddens[n * S + i] = (g * trans[c] * e - later) * delta[c];
      suffix += __shfl_sync(kFull, incl, 0);
    }
  }
}

// Shared forward of the alpha compositor: fills per-warp shared arrays with alpha, T and the (sky-corrected) weights
// and returns the accumulation.
__device__ __forceinline__ float alpha_scan(const float* __restrict__ arow, int S, float eps, int sky, int lane,
                                            float* s_alpha, float* s_trans, float* s_w) {
  float carry = 1.0f, acc = 0.0f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const bool ok = i < S;
    const float a = ok ? arow[i] : 0.0f;
    const float om = ok ? add(sub(1.0f, a), eps) : 1.0f;
    const float incl = warp_inclusive_prod(om, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = carry * excl;
    const float w = a * T;
    if (ok) {
      s_alpha[i] = a;
      s_trans[i] = T;
      s_w[i] = w;
      acc += w;
    }
    carry *= __shfl_sync(kFull, incl, 31);
  }
  acc = warp_sum(acc);
  __syncwarp();
  if (sky && lane == 0) s_w[S - 1] = sub(add(s_w[S - 1], 1.0f), acc);
  __syncwarp();
  return acc;
}

__global__ void __launch_bounds__(kRayWarps * 32) alpha_composite_fwd_kernel(
    const float* __restrict__ alphas, const float* __restrict__ feats, nrb_intervals_t iv, int64_t N, int C, float eps,
    int sky, float* __restrict__ weights, float* __restrict__ features, float* __restrict__ depth,
    float* __restrict__ accumulation, float* __restrict__ transmittance) {
  extern __shared__ float smem[];
  const int S = iv.num_samples;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (n >= N) return;
  float* s_alpha = smem + warp * 3 * S;
  float* s_trans = s_alpha + S;
  float* s_w = s_trans + S;
  const float acc = alpha_scan(alphas + n * S, S, eps, sky, lane, s_alpha, s_trans, s_w);
  if (lane == 0 && accumulation != nullptr) accumulation[n] = acc;
  if (weights != nullptr)
    for (int i = lane; i < S; i += 32) weights[n * S + i] = s_w[i];
  if (transmittance != nullptr)
    for (int i = lane; i < S; i += 32) transmittance[n * S + i] = s_trans[i];
  if (depth != nullptr) {
    const float* st = iv.starts + n * iv.row_stride;
    const float* en = iv.ends + n * iv.row_stride;
    const int last = sky ? S - 1 : S;
    float d = 0.0f;
    for (int i = lane; i < last; i += 32) d += s_w[i] * mul(add(st[i], en[i]), 0.5f);
    d = warp_sum(d);
    if (lane == 0) depth[n] = d;
  }
  if (features != nullptr && feats != nullptr) {
    // lanes own channels: each sample row of C floats is one coalesced read
    for (int c = lane; c < C; c += 32) {
      const float* f = feats + n * S * C + c;
      float s = 0.0f;
#pragma unroll 16
      for (int i = 0; i < S; ++i) s += s_w[i] * __ldg(f + static_cast<size_t>(i) * C);
      features[n * C + c] = s;
    }
  }
}

__global__ void __launch_bounds__(kRayWarps * 32) alpha_composite_bwd_kernel(
    const float* __restrict__ alphas, const float* __restrict__ feats, nrb_intervals_t iv, int64_t N, int C, float eps,
    int sky, const float* __restrict__ dweights, const float* __restrict__ dfeatures,
    const float* __restrict__ ddepth, const float* __restrict__ dacc, float* __restrict__ dalphas,
    float* __restrict__ dfeats) {
  extern __shared__ float smem[];
  const int S = iv.num_samples;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (n >= N) return;
  float* s_alpha = smem + warp * (4 * S + NRB_MAX_MLP_WIDTH);
  float* s_trans = s_alpha + S;
  float* s_w = s_trans + S;
  float* s_g = s_w + S;      // dL/dw' then dL/dw (raw)
  float* s_df = s_g + S;     // upstream feature gradient [C]
  alpha_scan(alphas + n * S, S, eps, sky, lane, s_alpha, s_trans, s_w);
  const bool has_f = dfeatures != nullptr && feats != nullptr;
  if (has_f)
    for (int c = lane; c < C; c += 32) s_df[c] = dfeatures[n * C + c];
  __syncwarp();
  const float* st = iv.starts + n * iv.row_stride;
  const float* en = iv.ends + n * iv.row_stride;
  const float gd = ddepth != nullptr ? ddepth[n] : 0.0f;
  const int last = sky ? S - 1 : S;
  // G_s = dL/dw'_s ; lanes own samples here: every lane walks its own contiguous feature row (a channel-major variant
  // with coalesced rows and a transpose-reduce for the dot products was measured 35 % slower: 4x the instructions)
  for (int i = lane; i < S; i += 32) {
    float g = dweights != nullptr ? dweights[n * S + i] : 0.0f;
    if (i < last) g += gd * mul(add(st[i], en[i]), 0.5f);
    if (has_f) {
      const float4* f = reinterpret_cast<const float4*>(feats + (n * S + i) * C);
      float4* df = dfeats != nullptr ? reinterpret_cast<float4*>(dfeats + (n * S + i) * C) : nullptr;
      const float w = s_w[i];
      float dot = 0.0f;
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 v = __ldg(f + c4);
        const float4 u = *reinterpret_cast<const float4*>(s_df + 4 * c4);
        dot += v.x * u.x + v.y * u.y + v.z * u.z + v.w * u.w;
        if (df != nullptr) df[c4] = make_float4(w * u.x, w * u.y, w * u.z, w * u.w);
      }
      g += dot;
    }
    s_g[i] = g;
  }
  __syncwarp();
  // back through the sky fix-up and the accumulation: dL/dw_s = G_s - [sky] G_{S-1} + dacc
  const float g_last = sky ? s_g[S - 1] : 0.0f;
  const float ga = dacc != nullptr ? dacc[n] : 0.0f;
  __syncwarp();
  for (int i = lane; i < S; i += 32) s_g[i] = s_g[i] - g_last + ga;
  __syncwarp();
  // w_s = a_s T_s, T_s = prod_{j<s} om_j:  dL/da_s = g_s T_s - (1/om_s) sum_{k>s} g_k w_k(raw)
  float suffix = 0.0f;
  const int chunks = (S + 31) / 32;
  for (int c = chunks - 1; c >= 0; --c) {
    const int i = c * 32 + lane;
    const bool ok = i < S;
    const float a = ok ? s_alpha[i] : 0.0f;
    const float T = ok ? s_trans[i] : 0.0f;
    const float gw = ok ? s_g[i] * a * T : 0.0f;
    float incl = gw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float v = __shfl_down_sync(kFull, incl, o);
      if (lane + o < 32) incl += v;
    }
    float excl = __shfl_down_sync(kFull, incl, 1);
      if (lane == 31) excl = 0.0f;
      const float later = suffix + excl;
    if (ok) {
      const float om = add(sub(1.0f, a), eps);
      float tail;
      if (om != 0.0f) {
        tail = later / om;
      } else {
        // an opaque sample zeroes every later T; differentiate the product with this factor left out
        tail = 0.0f;
        float p = T;
        for (int k = i + 1; k < S; ++k) {
          tail += s_g[k] * s_alpha[k] * p;
          p *= add(sub(1.0f, s_alpha[k]), eps);
        }
      }
      dalphas[n * S + i] = s_g[i] * T - tail;
    }
    suffix += __shfl_sync(kFull, incl, 0);
  }
}

// accumulate_along_rays on dense samples: out[n,c] = sum_s w[n,s] v[n,s,c]  (v == nullptr: out[n] = sum_s w[n,s]).
__global__ void __launch_bounds__(kRayWarps * 32) accumulate_fwd_kernel(const float* __restrict__ w,
                                                                        const float* __restrict__ v, int64_t N, int S,
                                                                        int C, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kRayWarps + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* wr = w + n * S;
  if (v == nullptr) {
    float s = 0.0f;
    for (int i = lane; i < S; i += 32) s += wr[i];
    s = warp_sum(s);
    if (lane == 0) out[n] = s;
    return;
  }
  if (C >= 16) {  // lanes own channels, rows are coalesced
    for (int c = lane; c < C; c += 32) {
      const float* f = v + n * S * C + c;
      float s = 0.0f;
      for (int i = 0; i < S; ++i) s += __ldg(wr + i) * __ldg(f + static_cast<size_t>(i) * C);
      out[n * C + c] = s;
    }
  } else {  // few channels: lanes own samples
    for (int c = 0; c < C; ++c) {
      float s = 0.0f;
      for (int i = lane; i < S; i += 32) s += wr[i] * v[(n * S + i) * C + c];
      s = warp_sum(s);
      if (lane == 0) out[n * C + c] = s;
    }
  }
}

__global__ void __launch_bounds__(256) accumulate_bwd_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                             const float* __restrict__ dout, int64_t total, int S,
                                                             int C, float* __restrict__ dw, float* __restrict__ dv) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread per sample
  if (gid >= total) return;
  const int64_t n = gid / S;
  if (v == nullptr) {
    if (dw != nullptr) dw[gid] = dout[n];
    return;
  }
  const float wi = w[gid];
  float dot = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float g = __ldg(dout + n * C + c);
    dot = fmaf(g, v[gid * C + c], dot);
    if (dv != nullptr) dv[gid * C + c] = wi * g;
  }
  if (dw != nullptr) dw[gid] = dot;
}

// render_depth_simple (nerfstudio/models/neurad.py:721-728) for given weights: depth[n] = sum_s w[n,s] (start + end) / 2,
// and its derivative with respect to the weights.  The midpoints are formed in registers from the bin edges.
__global__ void __launch_bounds__(kRayWarps * 32) weighted_depth_fwd_kernel(const float* __restrict__ w, nrb_intervals_t iv,
                                                                            int64_t N, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * kRayWarps + (threadIdx.x >> 5);
  if (n >= N) return;
  const int S = iv.num_samples;
  const float* st = iv.starts + n * iv.row_stride;
  const float* en = iv.ends + n * iv.row_stride;
  float s = 0.0f;
  for (int i = lane; i < S; i += 32) s += w[n * S + i] * ((st[i] + en[i]) / 2.0f);
  s = warp_sum(s);
  if (lane == 0) out[n] = s;
}

__global__ void __launch_bounds__(256) weighted_depth_bwd_kernel(nrb_intervals_t iv, const float* __restrict__ dout,
                                                                 int64_t total, float* __restrict__ dw) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int S = iv.num_samples;
  const int64_t n = gid / S;
  const int i = static_cast<int>(gid - n * S);
  dw[gid] = dout[n] * ((iv.starts[n * iv.row_stride + i] + iv.ends[n * iv.row_stride + i]) / 2.0f);
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_weighted_depth_fwd(const float* weights, const nrb_intervals_t* iv, int64_t N, float* depth,
                                      nrb_stream_t stream) {
  if (int rc = check_intervals("nrb_weighted_depth_fwd", iv)) return rc;
  NRB_REQUIRE(weights && depth && N >= 0, NRB_ERR_BAD_ARG, "nrb_weighted_depth_fwd: null pointer");
  if (N == 0) return NRB_OK;
  weighted_depth_fwd_kernel<<<blocks_for(N, kRayWarps), kRayWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(weights, *iv, N,
                                                                                                             depth);
  return finish_launch("nrb_weighted_depth_fwd");
}

extern "C" int nrb_weighted_depth_bwd(const nrb_intervals_t* iv, const float* ddepth, int64_t N, float* dweights,
                                      nrb_stream_t stream) {
  if (int rc = check_intervals("nrb_weighted_depth_bwd", iv)) return rc;
  NRB_REQUIRE(ddepth && dweights && N >= 0, NRB_ERR_BAD_ARG, "nrb_weighted_depth_bwd: null pointer");
  if (N == 0) return NRB_OK;
  const int64_t total = N * iv->num_samples;
  weighted_depth_bwd_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(*iv, ddepth, total, dweights);
  return finish_launch("nrb_weighted_depth_bwd");
}

extern "C" int nrb_accumulate_fwd(const float* weights, const float* values, int64_t N, int32_t S, int32_t C,
                                  float* out, nrb_stream_t stream) {
  NRB_REQUIRE(weights && out && N >= 0 && S > 0, NRB_ERR_BAD_ARG, "nrb_accumulate_fwd: null pointer or bad size");
  NRB_REQUIRE(values == nullptr || C > 0, NRB_ERR_BAD_ARG, "nrb_accumulate_fwd: C must be positive");
  if (N == 0) return NRB_OK;
  accumulate_fwd_kernel<<<blocks_for(N, kRayWarps), kRayWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      weights, values, N, S, C, out);
  return finish_launch("nrb_accumulate_fwd");
}

extern "C" int nrb_accumulate_bwd(const float* weights, const float* values, const float* dout, int64_t N, int32_t S,
                                  int32_t C, float* dweights, float* dvalues, nrb_stream_t stream) {
  NRB_REQUIRE(weights && dout && N >= 0 && S > 0, NRB_ERR_BAD_ARG, "nrb_accumulate_bwd: null pointer or bad size");
  if (N == 0) return NRB_OK;
  const int64_t total = N * S;
  accumulate_bwd_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weights, values, dout, total, S, C, dweights, dvalues);
  return finish_launch("nrb_accumulate_bwd");
}

extern "C" int nrb_density_weights_fwd(const float* densities, const nrb_intervals_t* iv, int64_t N, float* weights,
                                       nrb_stream_t stream) {
  if (int rc = check_intervals("nrb_density_weights_fwd", iv)) return rc;
  NRB_REQUIRE(densities && weights && N >= 0, NRB_ERR_BAD_ARG, "nrb_density_weights_fwd: null pointer");
  if (N == 0) return NRB_OK;
  density_weights_fwd_kernel<<<blocks_for(N, kRayWarps), kRayWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      densities, *iv, N, weights);
  return finish_launch("nrb_density_weights_fwd");
}

extern "C" int nrb_density_weights_bwd(const float* densities, const nrb_intervals_t* iv, const float* dweights,
                                       int64_t N, float* ddensities, nrb_stream_t stream) {
  if (int rc = check_intervals("nrb_density_weights_bwd", iv)) return rc;
  NRB_REQUIRE(densities && dweights && ddensities && N >= 0, NRB_ERR_BAD_ARG,
              "nrb_density_weights_bwd: null pointer");
  if (N == 0) return NRB_OK;
  density_weights_bwd_kernel<<<blocks_for(N, kRayWarps), kRayWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      densities, *iv, dweights, N, ddensities);
  return finish_launch("nrb_density_weights_bwd");
}

static int check_composite(const char* who, const float* alphas, const nrb_intervals_t* iv, int64_t N, int32_t C,
                           const float* feats) {
  if (int rc = check_intervals(who, iv)) return rc;
  NRB_REQUIRE(alphas && N >= 0, NRB_ERR_BAD_ARG, "%s: null pointer", who);
  NRB_REQUIRE(C >= 0 && C <= NRB_MAX_MLP_WIDTH && C % 4 == 0, NRB_ERR_BAD_ARG,
              "%s: C must be a multiple of 4 in [0,%d]", who, NRB_MAX_MLP_WIDTH);
  NRB_REQUIRE(feats == nullptr || aligned16(feats), NRB_ERR_ALIGNMENT, "%s: feats must be 16-byte aligned", who);
  return NRB_OK;
}

extern "C" int nrb_alpha_composite_fwd(const float* alphas, const float* feats, const nrb_intervals_t* iv, int64_t N,
                                       int32_t C, float trans_eps, int32_t sky_sample, float* weights,
                                       float* features, float* depth, float* accumulation, float* transmittance,
                                       nrb_stream_t stream) {
  if (int rc = check_composite("nrb_alpha_composite_fwd", alphas, iv, N, C, feats)) return rc;
  if (N == 0) return NRB_OK;
  const size_t smem = sizeof(float) * kRayWarps * 3 * iv->num_samples;
  alpha_composite_fwd_kernel<<<blocks_for(N, kRayWarps), kRayWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      alphas, feats, *iv, N, C, trans_eps, sky_sample, weights, features, depth, accumulation, transmittance);
  return finish_launch("nrb_alpha_composite_fwd");
}

extern "C" int nrb_alpha_composite_bwd(const float* alphas, const float* feats, const nrb_intervals_t* iv, int64_t N,
                                       int32_t C, float trans_eps, int32_t sky_sample, const float* dweights,
                                       const float* dfeatures, const float* ddepth, const float* daccumulation,
                                       float* dalphas, float* dfeats, nrb_stream_t stream) {
  if (int rc = check_composite("nrb_alpha_composite_bwd", alphas, iv, N, C, feats)) return rc;
  NRB_REQUIRE(dalphas != nullptr, NRB_ERR_BAD_ARG, "nrb_alpha_composite_bwd: dalphas is null");
  NRB_REQUIRE(dfeats == nullptr || aligned16(dfeats), NRB_ERR_ALIGNMENT,
              "nrb_alpha_composite_bwd: dfeats must be 16-byte aligned");
  if (N == 0) return NRB_OK;
  const size_t smem = sizeof(float) * kRayWarps * (4 * iv->num_samples + NRB_MAX_MLP_WIDTH);
  alpha_composite_bwd_kernel<<<blocks_for(N, kRayWarps), kRayWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      alphas, feats, *iv, N, C, trans_eps, sky_sample, dweights, dfeatures, ddepth, daccumulation, dalphas, dfeats);
  return finish_launch("nrb_alpha_composite_bwd");
}
