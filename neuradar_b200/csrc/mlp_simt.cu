// Generic tiny-MLP forward/backward on the CUDA cores (fp32 FMA): any Linear+ReLU chain with <= 4 layers and
// widths <= 64.  This is the shape-agnostic path behind the `MLP` module (lidar decoder, radar heads, ...) and the
// numerical yardstick for the tensor-core field kernel.  Semantics: MLP.pytorch_fwd
// (nerfstudio/field_components/mlp.py:142-178): x -> relu(W0 x + b0) -> ... -> W_last h + b_last.
//
// One thread owns one sample.  Activations of a 128-sample tile sit feature-major in shared memory with a row
// stride of 129 floats, so the owner's accesses are conflict-free and the tile can still be staged to and from the
// row-major global layout with coalesced accesses.  Weights are read as broadcast float4s, 8 outputs per pass.
// The backward pass keeps per-CTA weight-gradient accumulators in shared memory across the tiles of a persistent
// grid and flushes them with one atomicAdd per entry at the end.
#include "common.cuh"

namespace nrb {

constexpr int kTile = 128;
constexpr int kStride = kTile + 1;

struct MlpDev {
  const float* w[NRB_MAX_MLP_LAYERS];
  const float* b[NRB_MAX_MLP_LAYERS];
  int dims[NRB_MAX_MLP_LAYERS + 1];
  int n;
};

struct MlpGradDev {
  float* w[NRB_MAX_MLP_LAYERS];
  float* b[NRB_MAX_MLP_LAYERS];
};

__host__ __device__ inline int pad8(int v) { return (v + 7) & ~7; }

// shared-memory carve-up helpers (identical on host and device)
__host__ __device__ inline int wt_floats(const int* dims, int n) {
  int t = 0;
  for (int i = 0; i < n; ++i) t += dims[i] * pad8(dims[i + 1]) + pad8(dims[i + 1]);
  return t;
}
__host__ __device__ inline int max_dim(const int* dims, int n) {
  int m = 0;
  for (int i = 0; i <= n; ++i) m = dims[i] > m ? dims[i] : m;
  return m;
}

// stage a row-major [rows, width] global tile into feature-major shared memory (zero-filled past `rows`)
__device__ __forceinline__ void load_tile_rowmajor(const float* __restrict__ g, int64_t row0, int rows, int width,
                                                   float* s) {
  const int total = kTile * width;
  for (int e = threadIdx.x; e < total; e += kTile) {
    const int r = e / width, c = e - r * width;
    s[c * kStride + r] = (r < rows) ? __ldg(g + (row0 + r) * width + c) : 0.0f;
  }
}
__device__ __forceinline__ void store_tile_rowmajor(float* __restrict__ g, int64_t row0, int rows, int width,
                                                    const float* s) {
  const int total = kTile * width;
  for (int e = threadIdx.x; e < total; e += kTile) {
    const int r = e / width, c = e - r * width;
    if (r < rows) g[(row0 + r) * width + c] = s[c * kStride + r];
  }
}

__global__ void __launch_bounds__(kTile) mlp_fwd_kernel(const __grid_constant__ MlpDev p, const float* __restrict__ x, float* __restrict__ y,
                                                        float* __restrict__ hidden, int64_t M) {
  extern __shared__ __align__(16) float smem[];
  const int t = threadIdx.x;
  const int md = max_dim(p.dims, p.n);
  float* wt = smem;  // per layer: Wt[in][pad8(out)] then bias[pad8(out)]
  float* act0 = smem + ((wt_floats(p.dims, p.n) + 3) & ~3);
  float* act1 = act0 + md * kStride;
  {
    float* dst = wt;
    for (int i = 0; i < p.n; ++i) {
      const int in = p.dims[i], out = p.dims[i + 1], op = pad8(out);
      for (int e = t; e < in * op; e += kTile) {
        const int k = e / op, j = e - k * op;
        dst[e] = (j < out) ? __ldg(p.w[i] + j * in + k) : 0.0f;
      }
      dst += in * op;
      for (int j = t; j < op; j += kTile) dst[j] = (j < out && p.b[i] != nullptr) ? __ldg(p.b[i] + j) : 0.0f;
      dst += op;
    }
  }
  const int64_t tiles = (M + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int rows = static_cast<int>(min(static_cast<int64_t>(kTile), M - row0));
    __syncthreads();
    load_tile_rowmajor(x, row0, rows, p.dims[0], act0);
    __syncthreads();
    float* a_in = act0;
    float* a_out = act1;
    const float* wl = wt;
    int64_t hid_off = 0;
    for (int i = 0; i < p.n; ++i) {
      const int in = p.dims[i], out = p.dims[i + 1], op = pad8(out);
      const float* bias = wl + in * op;
      const bool last = (i == p.n - 1);
      for (int jb = 0; jb < op; jb += 8) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = bias[jb + q];
        for (int k = 0; k < in; ++k) {
          const float a = a_in[k * kStride + t];
          const float4 w0 = *reinterpret_cast<const float4*>(wl + k * op + jb);
          const float4 w1 = *reinterpret_cast<const float4*>(wl + k * op + jb + 4);
          acc[0] = fmaf(a, w0.x, acc[0]);
          acc[1] = fmaf(a, w0.y, acc[1]);
          acc[2] = fmaf(a, w0.z, acc[2]);
          acc[3] = fmaf(a, w0.w, acc[3]);
          acc[4] = fmaf(a, w1.x, acc[4]);
          acc[5] = fmaf(a, w1.y, acc[5]);
          acc[6] = fmaf(a, w1.z, acc[6]);
          acc[7] = fmaf(a, w1.w, acc[7]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = jb + q;
          if (j < out) {
            const float v = last ? acc[q] : fmaxf(acc[q], 0.0f);
            a_out[j * kStride + t] = v;
            if (!last && hidden != nullptr && t < rows) hidden[(hid_off + j) * M + row0 + t] = v;
          }
        }
      }
      if (!last) hid_off += out;
      wl += in * op + op;
      float* tmp = a_in;
      a_in = a_out;
      a_out = tmp;
      // the next layer reads only this thread's own column: no barrier needed between layers
    }
    __syncthreads();
    store_tile_rowmajor(y, row0, rows, p.dims[p.n], a_in);
  }
}

__global__ void __launch_bounds__(kTile) mlp_bwd_kernel(const __grid_constant__ MlpDev p, const __grid_constant__ MlpGradDev gr, const float* __restrict__ x,
                                                        const float* __restrict__ hidden,
                                                        const float* __restrict__ dy, float* __restrict__ dx,
                                                        int64_t M) {
  extern __shared__ __align__(16) float smem[];
  const int t = threadIdx.x;
  const int md = max_dim(p.dims, p.n);
  // carve: W row-major per layer [out][pad8(in)], dW accumulators [out][in] + db[out], two tile buffers
  int w_floats = 0, g_floats = 0;
  for (int i = 0; i < p.n; ++i) {
    w_floats += p.dims[i + 1] * pad8(p.dims[i]);
    g_floats += p.dims[i + 1] * p.dims[i] + p.dims[i + 1];
  }
  float* ws = smem;
  float* gs = ws + ((w_floats + 3) & ~3);
  float* bufA = gs + ((g_floats + 3) & ~3);
  float* bufD = bufA + md * kStride;
  {
    float* dst = ws;
    for (int i = 0; i < p.n; ++i) {
      const int in = p.dims[i], out = p.dims[i + 1], ip = pad8(in);
      for (int e = t; e < out * ip; e += kTile) {
        const int j = e / ip, k = e - j * ip;
        dst[e] = (k < in) ? __ldg(p.w[i] + j * in + k) : 0.0f;
      }
      dst += out * ip;
    }
    for (int e = t; e < g_floats; e += kTile) gs[e] = 0.0f;
  }
  const int64_t tiles = (M + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int rows = static_cast<int>(min(static_cast<int64_t>(kTile), M - row0));
    __syncthreads();
    load_tile_rowmajor(dy, row0, rows, p.dims[p.n], bufD);
    float* D = bufD;
    float* A = bufA;
    for (int i = p.n - 1; i >= 0; --i) {
      const int in = p.dims[i], out = p.dims[i + 1], ip = pad8(in);
      // offsets of this layer inside the packed arrays
      int w_off = 0, g_off = 0;
      int64_t hid_off = 0;
      for (int q = 0; q < i; ++q) {
        w_off += p.dims[q + 1] * pad8(p.dims[q]);
        g_off += p.dims[q + 1] * p.dims[q] + p.dims[q + 1];
        if (q < i - 1) hid_off += p.dims[q + 1];
      }
      // stage the layer's input activations
      if (i == 0) {
        load_tile_rowmajor(x, row0, rows, in, A);
      } else {
        for (int e = t; e < in * kTile; e += kTile) {
          const int k = e / kTile;  // e % kTile == t
          A[k * kStride + t] = (t < rows) ? __ldg(hidden + (hid_off + k) * M + row0 + t) : 0.0f;
        }
      }
      __syncthreads();
      // weight / bias gradients of this tile: each thread owns (j,k) pairs p = t + 128 r
      {
        float* gw = gs + g_off;
        float* gb = gw + out * in;
        for (int pbase = 0; pbase < out * in; pbase += kTile * 4) {
          int jj[4], kk[4];
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int pi = pbase + r * kTile + t;
            const int pc = pi < out * in ? pi : 0;
            jj[r] = pc / in;
            kk[r] = pc - jj[r] * in;
          }
          for (int s = 0; s < kTile; ++s) {
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(D[jj[r] * kStride + s], A[kk[r] * kStride + s], acc[r]);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int pi = pbase + r * kTile + t;
            if (pi < out * in) gw[pi] += acc[r];
          }
        }
        if (t < out) {
          float s = 0.0f;
          for (int q = 0; q < kTile; ++q) s += D[t * kStride + q];
          gb[t] += s;
        }
      }
      __syncthreads();
      // data gradient: d_in[k] = sum_j W[j][k] D[j]; masked by ReLU of the producing layer; written over A
      if (i > 0 || dx != nullptr) {
        const float* wl = ws + w_off;
        for (int kb = 0; kb < ip; kb += 8) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < out; ++j) {
            const float d = D[j * kStride + t];
            const float4 w0 = *reinterpret_cast<const float4*>(wl + j * ip + kb);
            const float4 w1 = *reinterpret_cast<const float4*>(wl + j * ip + kb + 4);
            acc[0] = fmaf(d, w0.x, acc[0]);
            acc[1] = fmaf(d, w0.y, acc[1]);
            acc[2] = fmaf(d, w0.z, acc[2]);
            acc[3] = fmaf(d, w0.w, acc[3]);
            acc[4] = fmaf(d, w1.x, acc[4]);
            acc[5] = fmaf(d, w1.y, acc[5]);
            acc[6] = fmaf(d, w1.z, acc[6]);
            acc[7] = fmaf(d, w1.w, acc[7]);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = kb + q;
            if (k < in) {
              const float a = A[k * kStride + t];
              A[k * kStride + t] = (i > 0 && !(a > 0.0f)) ? 0.0f : acc[q];
            }
          }
        }
      }
      __syncthreads();
      float* tmp = D;
      D = A;
      A = tmp;
    }
    if (dx != nullptr) store_tile_rowmajor(dx, row0, rows, p.dims[0], D);
  }
  __syncthreads();
  // flush the per-CTA accumulators
  {
    int g_off = 0;
    for (int i = 0; i < p.n; ++i) {
      const int in = p.dims[i], out = p.dims[i + 1];
      if (gr.w[i] != nullptr)
        for (int e = t; e < out * in; e += kTile) atomicAdd(gr.w[i] + e, gs[g_off + e]);
      if (gr.b[i] != nullptr)
        for (int e = t; e < out; e += kTile) atomicAdd(gr.b[i] + e, gs[g_off + out * in + e]);
      g_off += out * in + out;
    }
  }
}

static int check_mlp(const char* who, const nrb_mlp_t* m) {
  NRB_REQUIRE(m != nullptr, NRB_ERR_BAD_ARG, "%s: null mlp", who);
  NRB_REQUIRE(m->num_layers >= 1 && m->num_layers <= NRB_MAX_MLP_LAYERS, NRB_ERR_BAD_ARG,
              "%s: num_layers %d not in [1,%d]", who, m->num_layers, NRB_MAX_MLP_LAYERS);
  for (int i = 0; i <= m->num_layers; ++i)
    NRB_REQUIRE(m->dims[i] >= 1 && m->dims[i] <= NRB_MAX_MLP_WIDTH, NRB_ERR_UNSUPPORTED,
                "%s: layer width %d not in [1,%d]", who, m->dims[i], NRB_MAX_MLP_WIDTH);
  for (int i = 0; i < m->num_layers; ++i)
    NRB_REQUIRE(m->weights[i] != nullptr, NRB_ERR_BAD_ARG, "%s: weights[%d] is null", who, i);
  return NRB_OK;
}

static MlpDev to_dev(const nrb_mlp_t* m) {
  MlpDev d{};
  d.n = m->num_layers;
  for (int i = 0; i < m->num_layers; ++i) {
    d.w[i] = m->weights[i];
    d.b[i] = m->biases[i];
  }
  for (int i = 0; i <= m->num_layers; ++i) d.dims[i] = m->dims[i];
  return d;
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_mlp_fwd(const nrb_mlp_t* mlp, const float* x, float* y, float* hidden, int64_t M,
                           nrb_stream_t stream) {
  if (int rc = check_mlp("nrb_mlp_fwd", mlp)) return rc;
  NRB_REQUIRE(x && y && M >= 0, NRB_ERR_BAD_ARG, "nrb_mlp_fwd: null pointer or negative M");
  if (M == 0) return NRB_OK;
  const MlpDev d = to_dev(mlp);
  const size_t smem =
      sizeof(float) * (((wt_floats(d.dims, d.n) + 3) & ~3) + 2 * static_cast<size_t>(max_dim(d.dims, d.n)) * kStride);
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(mlp_fwd_kernel), 200 * 1024);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_mlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + kTile - 1) / kTile;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, static_cast<int64_t>(sm_count()) * 3));
  mlp_fwd_kernel<<<grid, kTile, smem, static_cast<cudaStream_t>(stream)>>>(d, x, y, hidden, M);
  return finish_launch("nrb_mlp_fwd");
}

extern "C" int nrb_mlp_bwd(const nrb_mlp_t* mlp, const float* x, const float* hidden, const float* dy, float* dx,
                           const nrb_mlp_grad_t* grads, int64_t M, nrb_stream_t stream) {
  if (int rc = check_mlp("nrb_mlp_bwd", mlp)) return rc;
  NRB_REQUIRE(x && dy && grads && M >= 0, NRB_ERR_BAD_ARG, "nrb_mlp_bwd: null pointer or negative M");
  NRB_REQUIRE(mlp->num_layers == 1 || hidden != nullptr, NRB_ERR_BAD_ARG, "nrb_mlp_bwd: hidden activations required");
  if (M == 0) return NRB_OK;
  const MlpDev d = to_dev(mlp);
  MlpGradDev g{};
  for (int i = 0; i < d.n; ++i) {
    g.w[i] = grads->weights[i];
    g.b[i] = grads->biases[i];
  }
  int w_floats = 0, g_floats = 0;
  for (int i = 0; i < d.n; ++i) {
    w_floats += d.dims[i + 1] * pad8(d.dims[i]);
    g_floats += d.dims[i + 1] * d.dims[i] + d.dims[i + 1];
  }
  const size_t smem = sizeof(float) * (((w_floats + 3) & ~3) + ((g_floats + 3) & ~3) +
                                       2 * static_cast<size_t>(max_dim(d.dims, d.n)) * kStride);
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(mlp_bwd_kernel), 200 * 1024);
  NRB_REQUIRE(e == cudaSuccess, static_cast<int>(e), "nrb_mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int64_t tiles = (M + kTile - 1) / kTile;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, static_cast<int64_t>(sm_count()) * 2));
  mlp_bwd_kernel<<<grid, kTile, smem, static_cast<cudaStream_t>(stream)>>>(d, g, x, hidden, dy, dx, M);
  return finish_launch("nrb_mlp_bwd");
}
