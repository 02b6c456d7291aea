// Fused Adam / AdamW step over a flat fp32 segment: the optimiser the reference configures for
// its hash grids (Adam, lr 1e-2, eps 1e-15) and field MLPs (AdamW, weight decay 1e-7), configs/method_configs.py:393-400,
// stepped through GradScaler at engine/optimizers.py:159-181.  torch runs it as an unscale pass + several foreach
// kernels; here ONE pass reads p, g, m, v and writes p, m, v (28 B / parameter; 32 with the fused zero_grad), with the
// data-parallel average and the loss-scale division folded into the gradient read.  Pure HBM streaming: float4 accesses,
// grid-stride over a grid sized in multiples of the SM count.
// Arithmetic follows torch/optim/adam.py (_single_tensor_adam): lerp for the first moment, mul + addcmul for the second,
// denom = sqrt(v) / sqrt(1 - beta2^t) + eps, p -= lr / (1 - beta1^t) * m / denom.
#include <cmath>

#include "common.cuh"

namespace nrb {

struct AdamDev {  // every constant is formed in double on the host (as torch's Python floats are) and rounded once
  float one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay;
  float decay;           // AdamW: 1 - lr * weight_decay
  float step_size;       // lr / (1 - beta1^t)
  float bc2_sqrt;        // sqrt(1 - beta2^t)
  float grad_mult;
  int decoupled, zero_grad;
  double lr_d, beta1_d, beta2_d;  // for the on-device bias corrections when steps can be skipped
  int step;
};

__device__ __forceinline__ void adam_update(float& p, float& g, float& m, float& v, const AdamDev& c, float gmul,
                                            float step_size, float bc2_sqrt) {
  float grad = g * gmul;
  if (c.decoupled) {
    p *= c.decay;
  } else if (c.weight_decay != 0.0f) {
    grad = fmaf(c.weight_decay, p, grad);
  }
  m = fmaf(c.one_minus_beta1, grad - m, m);
  v = fmaf(c.one_minus_beta2, grad * grad, v * c.beta2);
  const float denom = sqrtf(v) / bc2_sqrt + c.eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const AdamDev c,
                                                        const float* __restrict__ loss_scale,
                                                        const float* __restrict__ found_inf, float* skipped_steps) {
  const bool skip = found_inf != nullptr && __ldg(found_inf) != 0.0f;  // GradScaler: skip the step, keep the state
  float step_size = c.step_size, bc2_sqrt = c.bc2_sqrt;
  if (skipped_steps != nullptr) {
    // torch does not count a skipped step (optimizer.step() is never called), so the bias corrections use
    // step - (number of skipped steps so far), kept on the device to stay free of host synchronisation.  Within one
    // launch the counter is either only read (no skip) or only written (skip, nobody needs it).
    if (skip) {
      if (blockIdx.x == 0 && threadIdx.x == 0) *skipped_steps += 1.0f;
    } else {
      const int s = c.step - static_cast<int>(*skipped_steps);
      step_size = static_cast<float>(c.lr_d / (1.0 - pow(c.beta1_d, static_cast<double>(s))));
      bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(c.beta2_d, static_cast<double>(s))));
    }
  }
  const float gmul = c.grad_mult / (loss_scale != nullptr ? __ldg(loss_scale) : 1.0f);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    if (!skip) {
      // moments and gradients stream through (evict-first); the parameters are written with the default policy so that
      // the tables are still L2 resident when the next forward gathers from them
      float4 pp = reinterpret_cast<float4*>(p)[i], gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
      float4 mm = __ldcs(reinterpret_cast<const float4*>(m) + i), vv = __ldcs(reinterpret_cast<const float4*>(v) + i);
      adam_update(pp.x, gg.x, mm.x, vv.x, c, gmul, step_size, bc2_sqrt);
      adam_update(pp.y, gg.y, mm.y, vv.y, c, gmul, step_size, bc2_sqrt);
      adam_update(pp.z, gg.z, mm.z, vv.z, c, gmul, step_size, bc2_sqrt);
      adam_update(pp.w, gg.w, mm.w, vv.w, c, gmul, step_size, bc2_sqrt);
      reinterpret_cast<float4*>(p)[i] = pp;
      __stcs(reinterpret_cast<float4*>(m) + i, mm);
      __stcs(reinterpret_cast<float4*>(v) + i, vv);
    }
    if (c.zero_grad) __stcs(reinterpret_cast<float4*>(g) + i, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  // tail (n % 4 elements)
  const int64_t tail0 = n4 << 2;
  const int64_t j = tail0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j < n) {
    if (!skip) {
      float pp = p[j], gg = g[j], mm = m[j], vv = v[j];
      adam_update(pp, gg, mm, vv, c, gmul, step_size, bc2_sqrt);
      p[j] = pp;
      m[j] = mm;
      v[j] = vv;
    }
    if (c.zero_grad) g[j] = 0.0f;
  }
}

// found_inf[0] = 1 if any gradient is inf / nan (never cleared here: one flag can cover several segments)
__global__ void __launch_bounds__(256) grad_check_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ found_inf) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  bool bad = false;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(g) + i);
    // x - x is 0 for finite x and nan for inf / nan
    bad |= ((t.x - t.x) + (t.y - t.y) + (t.z - t.z) + (t.w - t.w)) != 0.0f;
  }
  const int64_t j = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j < n) bad |= (g[j] - g[j]) != 0.0f;
  if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;
}

}  // namespace nrb

using namespace nrb;

static unsigned stream_grid(int64_t n4) {
  const int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;  // 8 resident 256-thread CTAs per SM
  return static_cast<unsigned>(std::max<int64_t>(1, std::min(blocks, cap)));
}

extern "C" int nrb_adam_step(float* p, float* g, float* m, float* v, int64_t n, const nrb_adam_t* cfg,
                             const float* loss_scale, const float* found_inf, float* skipped_steps, nrb_stream_t stream) {
  NRB_REQUIRE(p && g && m && v && cfg && n >= 0, NRB_ERR_BAD_ARG, "nrb_adam_step: null pointer or negative size");
  NRB_REQUIRE(cfg->step >= 1, NRB_ERR_BAD_ARG, "nrb_adam_step: step counts from 1");
  NRB_REQUIRE(cfg->beta1 >= 0. && cfg->beta1 < 1. && cfg->beta2 >= 0. && cfg->beta2 < 1. && cfg->eps >= 0.,
              NRB_ERR_BAD_ARG, "nrb_adam_step: betas must be in [0,1) and eps >= 0");
  NRB_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), NRB_ERR_ALIGNMENT,
              "nrb_adam_step: arrays must be 16-byte aligned");
  if (n == 0) return NRB_OK;
  AdamDev c;
  c.one_minus_beta1 = static_cast<float>(1.0 - cfg->beta1);
  c.beta2 = static_cast<float>(cfg->beta2);
  c.one_minus_beta2 = static_cast<float>(1.0 - cfg->beta2);
  c.eps = static_cast<float>(cfg->eps);
  c.weight_decay = static_cast<float>(cfg->weight_decay);
  c.decay = static_cast<float>(1.0 - cfg->lr * cfg->weight_decay);
  const double bc1 = 1.0 - std::pow(cfg->beta1, cfg->step), bc2 = 1.0 - std::pow(cfg->beta2, cfg->step);
  c.step_size = static_cast<float>(cfg->lr / bc1);
  c.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  c.lr_d = cfg->lr;
  c.beta1_d = cfg->beta1;
  c.beta2_d = cfg->beta2;
  c.step = cfg->step;
  c.grad_mult = cfg->grad_mult;
  c.decoupled = cfg->decoupled_weight_decay;
  c.zero_grad = cfg->zero_grad;
  adam_step_kernel<<<stream_grid(std::max<int64_t>(n >> 2, 1)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, c, loss_scale, found_inf, skipped_steps);
  return finish_launch("nrb_adam_step");
}

extern "C" int nrb_grad_check(const float* g, int64_t n, float* found_inf, nrb_stream_t stream) {
  NRB_REQUIRE(g && found_inf && n >= 0, NRB_ERR_BAD_ARG, "nrb_grad_check: null pointer or negative size");
  NRB_REQUIRE(aligned16(g), NRB_ERR_ALIGNMENT, "nrb_grad_check: g must be 16-byte aligned");
  if (n == 0) return NRB_OK;
  grad_check_kernel<<<stream_grid(std::max<int64_t>(n >> 2, 1)), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, found_inf);
  return finish_launch("nrb_grad_check");
}
