// tcgen05 / TMEM building blocks for the fused NeuRAD field MLP (sm_100a).
//
// Operand layout.  Every operand tile lives in shared memory in the canonical no-swizzle ("interleave") UMMA
// layout: 8-row x 16-byte core matrices (8 rows x 4 tf32 values, 128 contiguous bytes).  For a tile of R rows and
// K columns:   byte_offset(r, k) = (r / 8) * (K / 4) * 128 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4.
//   * As a K-major operand (reduction over the columns): LBO = 128 (next 4 columns), SBO = (K/4)*128 (next 8 rows).
//   * The same bytes are a valid MN-major operand with the roles of rows and columns swapped (reduction over the
//     ROWS): LBO = (K/4)*128, SBO = 128.  The backward pass uses this to form X^T dY and dY W without re-staging.
// One thread owns one row, so a row is written as K/4 16-byte stores; the 8 lanes of a quarter warp cover 128
// contiguous bytes, i.e. the stores are bank-conflict free.
//
// Precision.  The reference MLPs run in fp32 and parity is 1e-3 relative on outputs AND gradients, through five
// chained layers and a sigmoid with slope 20.  kind::tf32 keeps 10 mantissa bits, so every operand is split into
// hi = top 11 significant bits and lo = x - hi (exact), and each product is issued as hi*hi + hi*lo + lo*hi with
// fp32 accumulation in TMEM ("3xTF32"), which is accurate to ~2^-20.  The tensor pipe has two orders of magnitude of
// headroom for this workload, so the 3x MMA count is free.
#pragma once

#include "common.cuh"

namespace nrb {
namespace tc {

constexpr int kRows = 128;  // rows (samples) per tile = UMMA M = TMEM lanes

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 (SWIZZLE_NONE) [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, M = 128.
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}

// One lane of a converged warp (CUTLASS's elect_one_sync): ptxas recognises the elected region as single-threaded and
// emits straight-line UTCHMMA sequences; under a plain `threadIdx.x == 0` test it wraps every tcgen05.mma in an
// ELECT / BRA.U.ANY loop (seen in SASS), which costs ~50 cycles per instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// warp index that the compiler can prove warp-uniform
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = smem_u32(mbar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// non-blocking test of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* mbar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(mbar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int kCols>
__device__ __forceinline__ void tmem_free(uint32_t taddr) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// TMEM -> registers: lane i of warp w receives row 32*(w%4)+i, 16 consecutive fp32 columns starting at `taddr`.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Read kN (multiple of 16) accumulator columns of this thread's row.
template <int kN>
__device__ __forceinline__ void tmem_load_row(uint32_t tmem_base, int warp, int col0, float (&v)[kN]) {
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(col0);
#pragma unroll
  for (int c = 0; c < kN / 16; ++c) {
    float t[16];
    tmem_ld16(taddr + c * 16, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[c * 16 + i] = t[i];
  }
  tmem_ld_wait();
}

// byte offset of (row, 4-column chunk) in a canonical tile with K columns
__device__ __forceinline__ uint32_t tile_offset(int row, int chunk, int K) {
  return static_cast<uint32_t>((row >> 3) * (K / 4) * 128 + chunk * 128 + (row & 7) * 16);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Write one row (K values, K % 4 == 0) of an operand tile as its hi / lo tf32 halves.
template <int K>
__device__ __forceinline__ void store_row_split(char* hi_tile, char* lo_tile, int row, const float (&v)[K]) {
#pragma unroll
  for (int c = 0; c < K / 4; ++c) {
    float4 h, l;
    h.x = tf32_hi(v[4 * c + 0]);
    h.y = tf32_hi(v[4 * c + 1]);
    h.z = tf32_hi(v[4 * c + 2]);
    h.w = tf32_hi(v[4 * c + 3]);
    l.x = v[4 * c + 0] - h.x;
    l.y = v[4 * c + 1] - h.y;
    l.z = v[4 * c + 2] - h.z;
    l.w = v[4 * c + 3] - h.w;
    const uint32_t off = tile_offset(row, c, K);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

// Stage a row-major weight matrix w[n_rows][K] (global) as a canonical tile padded to n_pad rows, hi and lo halves.
// Called by the whole CTA.
__device__ __forceinline__ void stage_weight_split(const float* __restrict__ w, int n_rows, int n_pad, int K, char* hi_tile,
                                                   char* lo_tile) {
  for (int e = threadIdx.x; e < n_pad * K; e += blockDim.x) {
    const int r = e / K, k = e - r * K;
    const float v = (r < n_rows) ? __ldg(w + r * K + k) : 0.0f;
    const float h = tf32_hi(v);
    const uint32_t off = tile_offset(r, k >> 2, K) + (k & 3) * 4;
    *reinterpret_cast<float*>(hi_tile + off) = h;
    *reinterpret_cast<float*>(lo_tile + off) = v - h;
  }
}

// D[128, N] (+)= A[128, K] * B[N, K]^T with A and B K-major canonical tiles (hi/lo pairs), 3xTF32.  One thread.
__device__ __forceinline__ void issue_gemm_kmajor(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                  uint32_t b_lo, int K, int N, bool accumulate_first) {
  const uint32_t idesc = make_idesc(N, 0, 0);
  const uint32_t sbo = static_cast<uint32_t>(K / 4) * 128u;
  bool acc = accumulate_first;
  for (int k = 0; k < K / 8; ++k) {
    const uint32_t koff = static_cast<uint32_t>(k) * 256u;  // 8 tf32 = two 16-byte chunks = two core matrices
    const uint64_t ah = make_desc(a_hi + koff, 128, sbo), al = make_desc(a_lo + koff, 128, sbo);
    const uint64_t bh = make_desc(b_hi + koff, 128, sbo), bl = make_desc(b_lo + koff, 128, sbo);
    mma_tf32(d_tmem, al, bh, idesc, acc);  // small terms first
    mma_tf32(d_tmem, ah, bl, idesc, true);
    mma_tf32(d_tmem, ah, bh, idesc, true);
    acc = true;
  }
}

}  // namespace tc
}  // namespace nrb

namespace nrb {
namespace tc {

// 32-bit instruction descriptor with explicit M (64 or 128).
__device__ __forceinline__ uint32_t make_idesc_m(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// General K-major x K-major chain: D[m, n] (+)= A[m, k_red] * B[n, k_red]^T, A / B canonical tiles with a_cols / b_cols
// columns (>= k_red), m in {64, 128}.  3xTF32.  One thread.
// (MN-major operands were measured to produce all-zero accumulators with kind::tf32 and SWIZZLE_NONE on B200 for every
//  LBO/SBO assignment, so the backward pass stages explicit transposes instead; see tools_tc_probe.py.)
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, int m, int n, uint32_t a_hi, uint32_t a_lo, int a_cols,
                                           uint32_t b_hi, uint32_t b_lo, int b_cols, int k_red, bool accumulate_first) {
  const uint32_t idesc = make_idesc_m(m, n, 0, 0);
  const uint32_t a_sbo = static_cast<uint32_t>(a_cols / 4) * 128u, b_sbo = static_cast<uint32_t>(b_cols / 4) * 128u;
  // the start-address field holds (address >> 4) in the low 14 bits: one k-step (8 tf32 = 256 bytes) adds 16
  uint64_t ah = make_desc(a_hi, 128, a_sbo), al = make_desc(a_lo, 128, a_sbo);
  uint64_t bh = make_desc(b_hi, 128, b_sbo), bl = make_desc(b_lo, 128, b_sbo);
  bool acc = accumulate_first;
  const int steps = k_red / 8;
#pragma unroll 4
  for (int k = 0; k < steps; ++k) {
    mma_tf32(d_tmem, al, bh, idesc, acc);  // small terms first
    mma_tf32(d_tmem, ah, bl, idesc, true);
    mma_tf32(d_tmem, ah, bh, idesc, true);
    acc = true;
    ah += 16;
    al += 16;
    bh += 16;
    bl += 16;
  }
}

// Stage the TRANSPOSE of a row-major weight matrix w[n_rows][k_cols] as a canonical tile with k_cols rows and n_pad
// columns (zero beyond n_rows), hi and lo halves.  Whole CTA.
__device__ __forceinline__ void stage_weight_transposed_split(const float* __restrict__ w, int n_rows, int n_pad,
                                                              int k_cols, char* hi_tile, char* lo_tile) {
  for (int e = threadIdx.x; e < k_cols * n_pad; e += blockDim.x) {
    const int r = e / n_pad, c = e - r * n_pad;  // tile row = input index, tile column = output index
    const float v = (c < n_rows) ? __ldg(w + c * k_cols + r) : 0.0f;
    const float h = tf32_hi(v);
    const uint32_t off = tile_offset(r, c >> 2, n_pad) + (c & 3) * 4;
    *reinterpret_cast<float*>(hi_tile + off) = h;
    *reinterpret_cast<float*>(lo_tile + off) = v - h;
  }
}

// Transpose a canonical [128 rows x C cols] tile into a canonical [C rows x 128 cols] tile.  kSplit: the source holds
// raw fp32 values that are split into hi / lo on the way; otherwise src_a / src_b are already the hi / lo halves and
// are transposed one to one.  Whole CTA (128 threads); callers synchronise around it.
template <bool kSplit>
__device__ __forceinline__ void transpose_tile(const char* src_a, const char* src_b, int C, char* dst_hi, char* dst_lo) {
  for (int e = threadIdx.x; e < C * 32; e += kRows) {
    const int j = (e & 7) + 8 * (e >> 8);  // feature (source column, destination row)
    const int c4 = (e >> 3) & 31;          // chunk of 4 samples (source rows 4*c4 .. 4*c4+3)
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t off = tile_offset(4 * c4 + i, j >> 2, C) + (j & 3) * 4;
      a[i] = *reinterpret_cast<const float*>(src_a + off);
      if constexpr (!kSplit) b[i] = *reinterpret_cast<const float*>(src_b + off);
    }
    float4 h, l;
    if constexpr (kSplit) {
      h = make_float4(tf32_hi(a[0]), tf32_hi(a[1]), tf32_hi(a[2]), tf32_hi(a[3]));
      l = make_float4(a[0] - h.x, a[1] - h.y, a[2] - h.z, a[3] - h.w);
    } else {
      h = make_float4(a[0], a[1], a[2], a[3]);
      l = make_float4(b[0], b[1], b[2], b[3]);
    }
    const uint32_t doff = tile_offset(j, c4, kRows);
    *reinterpret_cast<float4*>(dst_hi + doff) = h;
    *reinterpret_cast<float4*>(dst_lo + doff) = l;
  }
}

// Transposed staging straight from registers: thread t = 32*warp + lane holds v[0..31] = row t of a [128 x 32] matrix;
// the routine writes the TRANSPOSE as a canonical [32 rows x 128 cols] tile (hi / lo halves).  Each group of 4
// adjacent lanes transposes 4x4 blocks with two shuffle rounds, so that a lane ends up with 4 consecutive samples of
// one feature = one 16-byte chunk of the destination.  row_offset shifts the destination rows (for [emb | sh] etc).
template <int kGroups = 8>
__device__ __forceinline__ void store_rows_transposed_split(char* hi_tile, char* lo_tile, int warp, int lane,
                                                            const float (&v)[32], int row_offset = 0) {
  const int i = lane & 3;
  const int c4 = warp * 8 + (lane >> 2);  // 4-sample chunk index of this lane group
#pragma unroll
  for (int g = 0; g < kGroups; ++g) {  // kGroups * 4 destination rows
    float a0 = v[4 * g], a1 = v[4 * g + 1], a2 = v[4 * g + 2], a3 = v[4 * g + 3];
    // round 1 (xor 1): swap the off-diagonal elements of each 2x2 block
    {
      const bool odd = (i & 1) != 0;
      const float s01 = odd ? a0 : a1, s23 = odd ? a2 : a3;
      const float r01 = __shfl_xor_sync(kFull, s01, 1), r23 = __shfl_xor_sync(kFull, s23, 1);
      if (odd) {
        a0 = r01;
        a2 = r23;
      } else {
        a1 = r01;
        a3 = r23;
      }
    }
    // round 2 (xor 2): swap the off-diagonal 2x2 blocks
    {
      const bool up = (i & 2) != 0;
      const float s0 = up ? a0 : a2, s1 = up ? a1 : a3;
      const float r0 = __shfl_xor_sync(kFull, s0, 2), r1 = __shfl_xor_sync(kFull, s1, 2);
      if (up) {
        a0 = r0;
        a1 = r1;
      } else {
        a2 = r0;
        a3 = r1;
      }
    }
    // this lane now holds feature 4g+i of samples 4*c4 .. 4*c4+3
    float4 h = make_float4(tf32_hi(a0), tf32_hi(a1), tf32_hi(a2), tf32_hi(a3));
    float4 l = make_float4(a0 - h.x, a1 - h.y, a2 - h.z, a3 - h.w);
    const uint32_t off = tile_offset(row_offset + 4 * g + i, c4, kRows);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

// ---- coalesced I/O of row-major [M,32] matrices for "thread owns a row" kernels --------------------------------
// A thread reading its own 128-byte row makes every warp load touch 32 different lines (measured: 0.87 ms to stream
// 2 x 403 MB that way, 7x below the copy rate).  Instead the warp reads its 32 rows = 4 KB contiguous with 8
// coalesced 16-byte loads per lane and redistributes through a 4 KB per-warp bounce buffer whose 16-byte chunks are
// XOR-swizzled with the row index, which makes both the chunk-major and the row-major access conflict free.
__device__ __forceinline__ void warp_load_rows_coalesced(const float* __restrict__ g, int64_t row_base, int64_t M, int lane,
                                                         float4 (&pf)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int64_t r = min(row_base + 4 * q + (lane >> 3), M - 1);
    pf[q] = __ldg(reinterpret_cast<const float4*>(g + r * 32) + (lane & 7));
  }
}

__device__ __forceinline__ void warp_bounce_to_rows(char* bounce, int lane, const float4 (&pf)[8], float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int r = 4 * q + (lane >> 3), cc = lane & 7;
    *reinterpret_cast<float4*>(bounce + r * 128 + ((cc ^ (r & 7)) << 4)) = pf[q];
  }
  __syncwarp();
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    const float4 t = *reinterpret_cast<const float4*>(bounce + lane * 128 + ((cc ^ (lane & 7)) << 4));
    v[4 * cc] = t.x;
    v[4 * cc + 1] = t.y;
    v[4 * cc + 2] = t.z;
    v[4 * cc + 3] = t.w;
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_store_rows_coalesced(float* __restrict__ g, int64_t row_base, int64_t M, char* bounce,
                                                          int lane, const float (&v)[32]) {
#pragma unroll
  for (int cc = 0; cc < 8; ++cc)
    *reinterpret_cast<float4*>(bounce + lane * 128 + ((cc ^ (lane & 7)) << 4)) =
        make_float4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int r = 4 * q + (lane >> 3), cc = lane & 7;
    const float4 t = *reinterpret_cast<const float4*>(bounce + r * 128 + ((cc ^ (r & 7)) << 4));
    if (row_base + r < M) reinterpret_cast<float4*>(g + (row_base + r) * 32)[cc] = t;
  }
  __syncwarp();
}

// Write one row (K values) of a canonical tile without splitting.
template <int K>
__device__ __forceinline__ void store_row_raw(char* tile, int row, const float (&v)[K]) {
#pragma unroll
  for (int c = 0; c < K / 4; ++c)
    *reinterpret_cast<float4*>(tile + tile_offset(row, c, K)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// After the call lane j holds the sum over the warp's 32 lanes of v[j] (31 shuffles).
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const bool upper = (lane & s) != 0;
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(kFull, send, s);
    }
  }
  return v[0];
}


// ---- column-split variants: a 128-row tile is handled by 128 * S threads; thread (h, r) owns the W = 32 / S columns
// [h*W, (h+1)*W) of row r.  More warps per SM sub-partition hide the shuffle / shared-memory / mbarrier latencies that
// dominate the one-thread-per-row kernels, and each thread's serial staging work shrinks by S.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// W (8, 16 or 32) accumulator columns of this thread's row, starting at column col0.
template <int W>
__device__ __forceinline__ void tmem_load_cols(uint32_t tmem_base, int warp, int col0, float (&v)[W]) {
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(col0);
  if constexpr (W == 8) {
    tmem_ld8(taddr, v);
  } else {
#pragma unroll
    for (int c = 0; c < W / 16; ++c) {
      float t[16];
      tmem_ld16(taddr + c * 16, t);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c * 16 + i] = t[i];
    }
  }
  tmem_ld_wait();
}

// columns [col0, col0 + W) of one row of a canonical K-column tile, hi / lo halves
template <int W>
__device__ __forceinline__ void store_row_part_split(char* hi_tile, char* lo_tile, int row, int K, int col0,
                                                     const float (&v)[W]) {
#pragma unroll
  for (int c = 0; c < W / 4; ++c) {
    const float4 h = make_float4(tf32_hi(v[4 * c]), tf32_hi(v[4 * c + 1]), tf32_hi(v[4 * c + 2]), tf32_hi(v[4 * c + 3]));
    const float4 l = make_float4(v[4 * c] - h.x, v[4 * c + 1] - h.y, v[4 * c + 2] - h.z, v[4 * c + 3] - h.w);
    const uint32_t off = tile_offset(row, (col0 >> 2) + c, K);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

// Transposed staging of W columns: lane `lane` of row-quarter `quarter` holds v[0..W) = columns [col0, col0+W) of row
// 32 * quarter + lane; writes rows row_offset + [0, W) of the canonical [rows x 128] transposed tile (hi / lo).
template <int W>
__device__ __forceinline__ void store_part_transposed_split(char* hi_tile, char* lo_tile, int quarter, int lane,
                                                            const float (&v)[W], int row_offset) {
  const int i = lane & 3;
  const int c4 = quarter * 8 + (lane >> 2);
#pragma unroll
  for (int g = 0; g < W / 4; ++g) {
    float a0 = v[4 * g], a1 = v[4 * g + 1], a2 = v[4 * g + 2], a3 = v[4 * g + 3];
    {
      const bool odd = (i & 1) != 0;
      const float s01 = odd ? a0 : a1, s23 = odd ? a2 : a3;
      const float r01 = __shfl_xor_sync(kFull, s01, 1), r23 = __shfl_xor_sync(kFull, s23, 1);
      if (odd) {
        a0 = r01;
        a2 = r23;
      } else {
        a1 = r01;
        a3 = r23;
      }
    }
    {
      const bool up = (i & 2) != 0;
      const float s0 = up ? a0 : a2, s1 = up ? a1 : a3;
      const float r0 = __shfl_xor_sync(kFull, s0, 2), r1 = __shfl_xor_sync(kFull, s1, 2);
      if (up) {
        a0 = r0;
        a1 = r1;
      } else {
        a2 = r0;
        a3 = r1;
      }
    }
    const float4 h = make_float4(tf32_hi(a0), tf32_hi(a1), tf32_hi(a2), tf32_hi(a3));
    const float4 l = make_float4(a0 - h.x, a1 - h.y, a2 - h.z, a3 - h.w);
    const uint32_t off = tile_offset(row_offset + 4 * g + i, c4, kRows);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

// Coalesced I/O of columns [col0, col0 + W) of 32 consecutive rows of a row-major [M, 32] matrix by one warp:
// W / 4 lanes per row, W / 4 loads per lane, redistributed through a 32 * W * 4 byte bounce buffer whose 16-byte
// chunks are XOR-swizzled so that both the chunk-major and the row-major phase are bank-conflict free.
template <int W>
__device__ __forceinline__ int bounce_swizzle(int row) {
  constexpr int CPR = W / 4, RP = 8 / CPR;  // chunks per row, rows per 128 bytes
  return (row / RP) % CPR;
}

template <int W>
__device__ __forceinline__ void warp_load_part_coalesced(const float* __restrict__ g, int64_t row_base, int64_t M, int col0,
                                                         int lane, float4 (&pf)[W / 4]) {
  constexpr int CPR = W / 4, RPI = 32 / CPR;
#pragma unroll
  for (int q = 0; q < CPR; ++q) {
    const int64_t r = min(row_base + q * RPI + lane / CPR, M - 1);
    pf[q] = __ldg(reinterpret_cast<const float4*>(g + r * 32 + col0) + (lane % CPR));
  }
}

template <int W>
__device__ __forceinline__ void warp_bounce_part_to_rows(char* bounce, int lane, const float4 (&pf)[W / 4], float (&v)[W]) {
  constexpr int CPR = W / 4, RPI = 32 / CPR;
#pragma unroll
  for (int q = 0; q < CPR; ++q) {
    const int r = q * RPI + lane / CPR, cc = lane % CPR;
    *reinterpret_cast<float4*>(bounce + r * (W * 4) + ((cc ^ bounce_swizzle<W>(r)) << 4)) = pf[q];
  }
  __syncwarp();
#pragma unroll
  for (int cc = 0; cc < CPR; ++cc) {
    const float4 t = *reinterpret_cast<const float4*>(bounce + lane * (W * 4) + ((cc ^ bounce_swizzle<W>(lane)) << 4));
    v[4 * cc] = t.x;
    v[4 * cc + 1] = t.y;
    v[4 * cc + 2] = t.z;
    v[4 * cc + 3] = t.w;
  }
  __syncwarp();
}

template <int W>
__device__ __forceinline__ void warp_store_part_coalesced(float* __restrict__ g, int64_t row_base, int64_t M, int col0,
                                                          char* bounce, int lane, const float (&v)[W]) {
  constexpr int CPR = W / 4, RPI = 32 / CPR;
#pragma unroll
  for (int cc = 0; cc < CPR; ++cc)
    *reinterpret_cast<float4*>(bounce + lane * (W * 4) + ((cc ^ bounce_swizzle<W>(lane)) << 4)) =
        make_float4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < CPR; ++q) {
    const int r = q * RPI + lane / CPR, cc = lane % CPR;
    const float4 t = *reinterpret_cast<const float4*>(bounce + r * (W * 4) + ((cc ^ bounce_swizzle<W>(r)) << 4));
    if (row_base + r < M) reinterpret_cast<float4*>(g + (row_base + r) * 32 + col0)[cc] = t;
  }
  __syncwarp();
}


// ---- A operand from tensor memory (".ts" form): D[128, n] (+)= A[128, K] * B[n, K]^T where row r of A sits in TMEM lane
// r, one 32-bit column per K element.  Threads write their own rows with tcgen05.st; no shared-memory staging or
// shared-memory operand fetch for A.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// columns [col, col + W) of this thread's row, as tf32 hi / lo halves at hi_col / lo_col
template <int W>
__device__ __forceinline__ void tmem_store_row_split(uint32_t tmem_base, int warp, int hi_col, int lo_col, int col,
                                                     const float (&v)[W]) {
  const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll
  for (int c = 0; c < W / 8; ++c) {
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hi[i] = tf32_hi(v[8 * c + i]);
      lo[i] = v[8 * c + i] - hi[i];
    }
    tmem_st8(lane_base + static_cast<uint32_t>(hi_col + col + 8 * c), hi);
    tmem_st8(lane_base + static_cast<uint32_t>(lo_col + col + 8 * c), lo);
  }
  tmem_st_wait();
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}

// 3xTF32 chain with A (hi / lo) in TMEM columns and B (hi / lo) K-major canonical tiles of b_cols columns.  One thread.
// (Measured: these small instructions cost ~35-55 cycles each whatever the operand source and whether or not
//  consecutive ones share an accumulator - one accumulator per K step bought nothing and cost four TMEM loads.)
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, int n, uint32_t a_hi_tmem, uint32_t a_lo_tmem, uint32_t b_hi,
                                              uint32_t b_lo, int b_cols, int k_red) {
  const uint32_t idesc = make_idesc_m(128, n, 0, 0);
  const uint32_t b_sbo = static_cast<uint32_t>(b_cols / 4) * 128u;
  uint64_t bh = make_desc(b_hi, 128, b_sbo), bl = make_desc(b_lo, 128, b_sbo);
  bool acc = false;
#pragma unroll 4
  for (int k = 0; k < k_red / 8; ++k) {
    mma_tf32_ts(d_tmem, a_lo_tmem + 8 * k, bh, idesc, acc);  // small terms first
    mma_tf32_ts(d_tmem, a_hi_tmem + 8 * k, bl, idesc, true);
    mma_tf32_ts(d_tmem, a_hi_tmem + 8 * k, bh, idesc, true);
    acc = true;
    bh += 16;
    bl += 16;
  }
}

}  // namespace tc
}  // namespace nrb
