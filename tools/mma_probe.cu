// Scratch probe (not part of the product library): tcgen05.mma operand layouts and instruction throughput on sm_100a.
// The HOST builds byte images of the operand tiles for a layout hypothesis; the kernel copies them to shared memory
// (or to tensor memory for the ".ts" form), issues a chain of MMAs and dumps the accumulator.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_probe tools/mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

struct Cfg {
  int kind;        // 0 tf32, 1 bf16
  int a_tmem;      // 1: A from tensor memory
  int a_major, b_major;
  int m, n, ksteps;
  uint32_t a_lbo, a_sbo, a_step;  // bytes; a_step = descriptor advance per k-step (bytes) or TMEM columns if a_tmem
  uint32_t b_lbo, b_sbo, b_step;
  int a_bytes, b_bytes;           // image sizes
  int a_cols;                     // TMEM columns of the A image (a_tmem)
  int d_lane;                     // lane offset of the accumulator address
  int repeat;                     // timing: issue the chain this many times
  int issuers;                    // timing: number of issuing warps (each issues the chain `repeat` times)
  int swizzle;                    // layout_type field of both descriptors
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7) << 61;
  return d;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

template <int KIND, int TS>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a_desc, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool acc) {
  const uint32_t accu = acc;
  if constexpr (KIND == 0 && TS == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accu) : "memory");
  if constexpr (KIND == 1 && TS == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accu) : "memory");
  if constexpr (KIND == 0 && TS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accu) : "memory");
  if constexpr (KIND == 1 && TS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accu) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = smem_u32(mbar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}

template <int KIND, int TS>
__global__ void __launch_bounds__(256) probe_kernel(Cfg cfg, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
                                                    float* __restrict__ dump, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar[8];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((cfg.a_bytes + 1023) / 1024) * 1024;
  if (!TS)
    for (int i = t; i < cfg.a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (int i = t; i < cfg.b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0)
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t a_col0 = 256;  // TMEM columns of the A image
  if (TS && warp < 4) {         // image: [128 lanes][a_cols] uint32
    for (int c = 0; c < cfg.a_cols; ++c) {
      const uint32_t v = reinterpret_cast<const uint32_t*>(a_img)[(warp * 32 + lane) * cfg.a_cols + c];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + a_col0 + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // zero the accumulator region so that untouched lanes read as zero
  if (warp < 4) {
    for (int c = 0; c < 256; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t fmt = KIND == 0 ? 2u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)cfg.a_major << 15) | ((uint32_t)cfg.b_major << 16) |
                         ((uint32_t)(cfg.n >> 3) << 17) | ((uint32_t)(cfg.m >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp < cfg.issuers) {
    if (elect_one()) {
      const uint32_t d = tmem + ((uint32_t)cfg.d_lane << 16) + (uint32_t)(warp * (cfg.issuers > 1 ? 64 : 0));
      t0 = clock64();
      for (int r = 0; r < cfg.repeat; ++r) {
        uint64_t ad = make_desc(smem_u32(a_s), cfg.a_lbo, cfg.a_sbo, cfg.swizzle);
        uint64_t bd = make_desc(smem_u32(b_s), cfg.b_lbo, cfg.b_sbo, cfg.swizzle);
        uint32_t at = tmem + a_col0;
        for (int k = 0; k < cfg.ksteps; ++k) {
          mma<KIND, TS>(d, ad, at, bd, idesc, (k > 0) || (r > 0));
          ad += cfg.a_step >> 4;
          at += cfg.a_step;
          bd += cfg.b_step >> 4;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[warp])) : "memory");
      const long long ti = clock64();
      mbar_wait(&mbar[warp], 0);
      t1 = clock64();
      cycles[2 * warp] = t1 - t0;
      cycles[2 * warp + 1] = ti - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4 && dump != nullptr) {
    for (int c = 0; c < cfg.n; ++c) {
      uint32_t v;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      dump[(warp * 32 + lane) * 256 + c] = __uint_as_float(v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
static uint16_t bf16_of(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  return static_cast<uint16_t>(u >> 16);
}

struct Mat {
  int rows, cols;
  std::vector<float> v;
  Mat(int r, int c) : rows(r), cols(c), v(static_cast<size_t>(r) * c) {}
  float& at(int r, int c) { return v[static_cast<size_t>(r) * cols + c]; }
  float at(int r, int c) const { return v[static_cast<size_t>(r) * cols + c]; }
};

static Mat rand_mat(int r, int c, unsigned seed) {
  Mat m(r, c);
  srand(seed);
  for (auto& x : m.v) x = static_cast<float>(rand() % 7 - 3);
  return m;
}

// image of a [mn x k] operand.  K-major: byte(mn, k) = (mn/8)*sbo + (k/T)*lbo + (mn%8)*16 + (k%T)*e
//                               MN-major: byte(mn, k) = (mn/T)*sbo + (k/8)*lbo + (k%8)*16 + (mn%T)*e     (T = 16/e)
static std::vector<uint8_t> image(const Mat& m, int e, int major, uint32_t lbo, uint32_t sbo, size_t bytes) {
  std::vector<uint8_t> img(bytes, 0);
  const int T = 16 / e;
  for (int r = 0; r < m.rows; ++r)
    for (int k = 0; k < m.cols; ++k) {
      size_t off = major == 0 ? static_cast<size_t>(r / 8) * sbo + static_cast<size_t>(k / T) * lbo + (r % 8) * 16 + (k % T) * e
                              : static_cast<size_t>(r / T) * sbo + static_cast<size_t>(k / 8) * lbo + (k % 8) * 16 + (r % T) * e;
      if (off + e > bytes) {
        printf("image overflow\n");
        exit(1);
      }
      const float v = m.at(r, k);
      if (e == 4)
        memcpy(&img[off], &v, 4);
      else {
        const uint16_t h = bf16_of(v);
        memcpy(&img[off], &h, 2);
      }
    }
  return img;
}

template <int KIND, int TS>
static void launch(const Cfg& cfg, const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, std::vector<float>* dump, long long* cyc) {
  uint8_t *da, *db;
  float* dd;
  long long* dc;
  CK(cudaMalloc(&da, a.size() + 16));
  CK(cudaMalloc(&db, b.size() + 16));
  CK(cudaMalloc(&dd, 128 * 256 * 4));
  CK(cudaMalloc(&dc, 64 * 8));
  CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0, 128 * 256 * 4));
  CK(cudaMemset(dc, 0, 64 * 8));
  const int smem = ((cfg.a_bytes + 1023) / 1024) * 1024 + cfg.b_bytes + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<KIND, TS><<<1, 256, smem>>>(cfg, da, db, dd, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  kernel failed: %s\n", cudaGetErrorString(e));
    exit(2);
  }
  if (dump) {
    dump->resize(128 * 256);
    CK(cudaMemcpy(dump->data(), dd, 128 * 256 * 4, cudaMemcpyDeviceToHost));
  }
  if (cyc) CK(cudaMemcpy(cyc, dc, 64 * 8, cudaMemcpyDeviceToHost));
  cudaFree(da);
  cudaFree(db);
  cudaFree(dd);
  cudaFree(dc);
}

static void run(const Cfg& cfg, const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, std::vector<float>* dump, long long* cyc) {
  if (cfg.kind == 0 && !cfg.a_tmem) launch<0, 0>(cfg, a, b, dump, cyc);
  if (cfg.kind == 1 && !cfg.a_tmem) launch<1, 0>(cfg, a, b, dump, cyc);
  if (cfg.kind == 0 && cfg.a_tmem) launch<0, 1>(cfg, a, b, dump, cyc);
  if (cfg.kind == 1 && cfg.a_tmem) launch<1, 1>(cfg, a, b, dump, cyc);
}

// compare dump lanes against expected rows; reports the lane -> row mapping found
static void check(const char* tag, const std::vector<float>& dump, const Mat& expect, int n) {
  int matched = 0, identity = 0;
  int first_lane[4] = {-1, -1, -1, -1};
  int nonzero = 0;
  for (int lane = 0; lane < 128; ++lane) {
    bool nz = false;
    for (int c = 0; c < n; ++c) nz |= dump[lane * 256 + c] != 0.0f;
    nonzero += nz;
    for (int r = 0; r < expect.rows; ++r) {
      bool eq = true;
      for (int c = 0; c < n && eq; ++c) eq = dump[lane * 256 + c] == expect.at(r, c);
      if (eq) {
        ++matched;
        if (r == lane) ++identity;
        if (r < 4 && first_lane[r] < 0) first_lane[r] = lane;
        if (r == 16 || r == 32 || r == 63) printf("    row %d found in lane %d\n", r, lane);
        break;
      }
    }
  }
  printf("%-58s rows=%d matched=%d identity=%d nonzero_lanes=%d (row0..3 at lanes %d %d %d %d)\n", tag, expect.rows, matched, identity,
         nonzero, first_lane[0], first_lane[1], first_lane[2], first_lane[3]);
}

static Mat matmul_abt(const Mat& a, const Mat& b) {  // a [m x k], b [n x k] -> [m x n]
  Mat d(a.rows, b.rows);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.rows; ++j) {
      float s = 0;
      for (int k = 0; k < a.cols; ++k) s += a.at(i, k) * b.at(j, k);
      d.at(i, j) = s;
    }
  return d;
}

int main(int argc, char** argv) {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d, %d SMs, L2 %d MB\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.l2CacheSize >> 20);

  // ---------------- correctness of layouts ----------------
  for (int kind = 0; kind < 2; ++kind) {
    const int e = kind == 0 ? 4 : 2, T = 16 / e, kk = kind == 0 ? 8 : 16;  // kk = K per instruction
    const char* kn = kind == 0 ? "tf32" : "bf16";
    char tag[200];
    // (1) K-major A [128 x 32], K-major B [32 x 32]
    {
      Mat A = rand_mat(128, 32, 1), B = rand_mat(32, 32, 2);
      Cfg c{};
      c.kind = kind; c.m = 128; c.n = 32; c.ksteps = 32 / kk;
      c.a_lbo = 128; c.a_sbo = (32 / T) * 128; c.a_step = 2 * 128;
      c.b_lbo = 128; c.b_sbo = (32 / T) * 128; c.b_step = 2 * 128;
      c.a_bytes = 128 * 32 * e; c.b_bytes = 32 * 32 * e; c.repeat = 1; c.issuers = 1;
      auto ai = image(A, e, 0, c.a_lbo, c.a_sbo, c.a_bytes), bi = image(B, e, 0, c.b_lbo, c.b_sbo, c.b_bytes);
      std::vector<float> d;
      run(c, ai, bi, &d, nullptr);
      snprintf(tag, sizeof tag, "%s K/K M128 N32 K32", kn);
      check(tag, d, matmul_abt(A, B), 32);
      // same with M = 64 (first 64 rows), accumulator at lane 0 and lane 16
      for (int dl = 0; dl <= 16; dl += 16) {
        Mat A64(64, 32);
        for (int r = 0; r < 64; ++r) for (int k = 0; k < 32; ++k) A64.at(r, k) = A.at(r, k);
        c.m = 64; c.d_lane = dl;
        run(c, ai, bi, &d, nullptr);
        snprintf(tag, sizeof tag, "%s K/K M64 N32 K32 d_lane=%d", kn, dl);
        check(tag, d, matmul_abt(A64, B), 32);
      }
    }
    // (2) A from tensor memory: tf32 one element per column; bf16 two per column (even element in the low half)
    {
      Mat A = rand_mat(128, 32, 3), B = rand_mat(32, 32, 4);
      Cfg c{};
      c.kind = kind; c.a_tmem = 1; c.m = 128; c.n = 32; c.ksteps = 32 / kk;
      c.a_cols = kind == 0 ? 32 : 16; c.a_step = 8;  // columns per k-step: 8 tf32 / 16 bf16 = 8 columns
      c.b_lbo = 128; c.b_sbo = (32 / T) * 128; c.b_step = 2 * 128;
      c.a_bytes = 128 * c.a_cols * 4; c.b_bytes = 32 * 32 * e; c.repeat = 1; c.issuers = 1;
      std::vector<uint8_t> ai(c.a_bytes);
      for (int r = 0; r < 128; ++r)
        for (int col = 0; col < c.a_cols; ++col) {
          uint32_t w;
          if (kind == 0) { float v = A.at(r, col); memcpy(&w, &v, 4); }
          else w = static_cast<uint32_t>(bf16_of(A.at(r, 2 * col))) | (static_cast<uint32_t>(bf16_of(A.at(r, 2 * col + 1))) << 16);
          memcpy(&ai[(static_cast<size_t>(r) * c.a_cols + col) * 4], &w, 4);
        }
      auto bi = image(B, e, 0, c.b_lbo, c.b_sbo, c.b_bytes);
      std::vector<float> d;
      run(c, ai, bi, &d, nullptr);
      snprintf(tag, sizeof tag, "%s A in TMEM (.ts) M128 N32 K32", kn);
      check(tag, d, matmul_abt(A, B), 32);
    }
    // (3) MN-major A: reduction over 128 samples.  A^T stored as [mn = 64 features][k = 128 samples]
    for (int variant = 0; variant < 2; ++variant) {
      Mat A = rand_mat(64, 128, 5), B = rand_mat(48, 128, 6);  // A [m x k], B [n x k]
      Cfg c{};
      c.kind = kind; c.m = 64; c.n = 48; c.ksteps = 128 / kk; c.a_major = 1; c.b_major = variant;
      // MN-major image: byte(mn, k) = (mn/T)*sbo + (k/8)*lbo + (k%8)*16 + (mn%T)*e  with lbo = 128, sbo = 16 k-groups * 128
      c.a_lbo = 128; c.a_sbo = 16 * 128; c.a_step = (kk / 8) * 128;
      c.a_bytes = (64 / T) * 16 * 128;
      if (variant == 0) {  // B K-major [48 x 128]
        c.b_lbo = 128; c.b_sbo = (128 / T) * 128; c.b_step = 2 * 128; c.b_bytes = 48 * 128 * e;
      } else {  // B MN-major
        c.b_lbo = 128; c.b_sbo = 16 * 128; c.b_step = (kk / 8) * 128; c.b_bytes = (48 / T) * 16 * 128;
      }
      c.repeat = 1; c.issuers = 1;
      auto ai = image(A, e, 1, c.a_lbo, c.a_sbo, c.a_bytes);
      auto bi = image(B, e, variant, c.b_lbo, c.b_sbo, c.b_bytes);
      std::vector<float> d;
      run(c, ai, bi, &d, nullptr);
      snprintf(tag, sizeof tag, "%s A MN-major, B %s-major M64 N48 K128", kn, variant ? "MN" : "K");
      check(tag, d, matmul_abt(A, B), 48);
      // the same with lbo / sbo exchanged in the descriptor only (image unchanged), to learn the field roles
      std::swap(c.a_lbo, c.a_sbo);
      if (variant) std::swap(c.b_lbo, c.b_sbo);
      run(c, ai, bi, &d, nullptr);
      snprintf(tag, sizeof tag, "%s   .. descriptor lbo<->sbo swapped", kn);
      check(tag, d, matmul_abt(A, B), 48);
    }
    // (4) MN-major with M = 128 (stacked [hi; mid] style): A [128 x 128]
    {
      Mat A = rand_mat(128, 128, 7), B = rand_mat(48, 128, 8);
      Cfg c{};
      c.kind = kind; c.m = 128; c.n = 48; c.ksteps = 128 / kk; c.a_major = 1; c.b_major = 1;
      c.a_lbo = 128; c.a_sbo = 16 * 128; c.a_step = (kk / 8) * 128; c.a_bytes = (128 / T) * 16 * 128;
      c.b_lbo = 128; c.b_sbo = 16 * 128; c.b_step = (kk / 8) * 128; c.b_bytes = (48 / T) * 16 * 128;
      c.repeat = 1; c.issuers = 1;
      auto ai = image(A, e, 1, c.a_lbo, c.a_sbo, c.a_bytes), bi = image(B, e, 1, c.b_lbo, c.b_sbo, c.b_bytes);
      std::vector<float> d;
      run(c, ai, bi, &d, nullptr);
      snprintf(tag, sizeof tag, "%s MN/MN M128 N48 K128", kn);
      check(tag, d, matmul_abt(A, B), 48);
    }
  }

  // ---------------- instruction throughput ----------------
  printf("\nthroughput: cycles per MMA (total incl. completion / issue only), one CTA, chain repeated\n");
  for (int kind = 0; kind < 2; ++kind) {
    const int e = kind == 0 ? 4 : 2, T = 16 / e;
    for (int ts = 0; ts < 2; ++ts)
      for (int n : {32, 48, 64, 128, 256})
        for (int issuers : {1, 2, 4}) {
          if (issuers > 1 && n > 64) continue;
          Cfg c{};
          c.kind = kind; c.a_tmem = ts; c.m = 128; c.n = n; c.ksteps = 4;
          c.a_lbo = 128; c.a_sbo = (32 / T) * 128; c.a_step = ts ? 8 : 2 * 128;
          if (kind == 1) { c.ksteps = 2; }
          c.b_lbo = 128; c.b_sbo = (32 / T) * 128; c.b_step = 2 * 128;
          c.a_cols = kind == 0 ? 32 : 16;
          c.a_bytes = 128 * 32 * 4; c.b_bytes = 256 * 32 * e; c.repeat = 256; c.issuers = issuers;
          std::vector<uint8_t> ai(c.a_bytes, 0), bi(c.b_bytes, 0);
          long long cyc[64] = {0};
          run(c, ai, bi, nullptr, cyc);
          const double per = static_cast<double>(cyc[0]) / (c.repeat * c.ksteps);
          const double iss = static_cast<double>(cyc[1]) / (c.repeat * c.ksteps);
          printf("  %s %s N=%3d issuers=%d: %.1f cycles/MMA per issuer (issue %.1f) -> %.1f cycles per MMA overall\n", kind ? "bf16" : "tf32",
                 ts ? "A=tmem" : "A=smem", n, issuers, per, iss, per / issuers);
        }
  }
  return 0;
}
