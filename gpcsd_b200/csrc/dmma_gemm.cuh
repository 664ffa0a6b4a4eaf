// FP64 tensor-core (DMMA.8x8x4) GEMM building blocks for sm_100a.
//
// B200 facts this is designed around (measured, profiles/r01_ubench_fp64.md):
//   * mma.sync.m8n8k4.f64 -> one DMMA.8x8x4 per 16 issue cycles per SM sub-partition = 64 FMA/clk/SM
//     = 37.0 TFLOP/s chip-wide; plain DFMA reaches the same rate, so the FP64 pipe is the roofline and
//     everything else (LDS, address math, barriers) must hide inside the 15 free issue slots.
//   * tcgen05 / TMEM have no FP64 kind: accumulators live in registers.
// Layout choices:
//   * K-major operand tiles are stored [rows][BK+4] doubles, N-major tiles [BK][BN+4]: a row stride
//     == 4 (mod 16) doubles makes the 16 lanes of each half-warp LDS.64 fragment read hit 16 distinct
//     8-byte bank pairs (lane = 4*g + q reads row g / k q, or k q / column g).
//   * global -> shared by 16-byte cp.async.cg (L2 only, no register staging) with zero-fill predication
//     for the M/N/K edges, STAGES-deep ring, one __syncthreads per 16-deep k-block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpcsd {

constexpr int BK = 16;            // k-block depth (doubles): 128 bytes per operand row
constexpr int KMAJ_LD = BK + 4;   // smem row stride of K-major tiles
constexpr int NTHREADS = 256;     // 8 warps per CTA

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4*g + q holds A[g][q], B[q][g], C[g][2q], C[g][2q+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int clamp_bytes(long remaining_elems) {
  return remaining_elems >= 2 ? 16 : (remaining_elems == 1 ? 8 : 0);
}

// ---- tile loaders ---------------------------------------------------------------------------------
// All per-thread addressing is hoisted out of the k loop: a loader object is built once per CTA tile and
// each k-block costs one 64-bit add + one LDGSTS per 16-byte chunk on the interior fast path (the
// branchy partial-chunk logic only runs for the last, ragged k-block).  Measured before this change: the
// loader's integer code stalled both warps of every SMSP for ~1.4k of every 5.7k cycles per k-block.

// K-major tile: ROWS x BK doubles from a row-major matrix (row stride ld, K contiguous).
// thread t copies chunks (row = (t>>3) + 32*i, ch = t&7), i < ROWS/32.
template <int ROWS>
struct KMajorLoader {
  static constexpr int NCH = ROWS / 32;
  const double* g;    // this thread's first chunk at k = 0 (row t>>3, chunk t&7)
  const double* safe; // any valid address (used with src-size 0)
  long row_step;      // 32 * ld
  int soff;           // smem offset of the first chunk
  int ch2;            // 2 * (t & 7)
  unsigned rowmask;   // bit i: row (t>>3) + 32 i is in range

  __device__ __forceinline__ void init(const double* base, long ld, int rows_valid, const double* safe_) {
    const int r0 = threadIdx.x >> 3;
    ch2 = 2 * (threadIdx.x & 7);
    g = base + (long)r0 * ld + ch2;
    safe = safe_;
    row_step = 32 * ld;
    soff = r0 * KMAJ_LD + ch2;
    rowmask = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (r0 + 32 * i < rows_valid) rowmask |= 1u << i;
  }
  // k_valid: number of in-range k from this block's first column on
  __device__ __forceinline__ void load(double* s, long k0, long k_valid) const {
    const int kb = (k_valid >= BK) ? 16 : clamp_bytes(k_valid - ch2);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int bytes = ((rowmask >> i) & 1u) ? kb : 0;
      const double* src = bytes ? (g + k0 + i * row_step) : safe;
      cp_async16(s + soff + i * 32 * KMAJ_LD, src, bytes);
    }
  }
};

// N-major tile: BK x COLS doubles from a row-major K x N matrix (row stride ld, N contiguous).
// thread t copies chunks (krow = t / (COLS/2) + RPI*i, ch = t % (COLS/2)).
template <int COLS>
struct NMajorLoader {
  static constexpr int CPR = COLS / 2;
  static constexpr int RPI = NTHREADS / CPR;   // k-rows covered per pass
  static constexpr int NCH = BK / RPI;
  static constexpr int LDS_ = COLS + 4;
  const double* g;
  const double* safe;
  long ld;
  int soff, krow0, nbytes;

  __device__ __forceinline__ void init(const double* base, long ld_, long n_valid, const double* safe_) {
    krow0 = threadIdx.x / CPR;
    const int ch = threadIdx.x % CPR;
    ld = ld_;
    g = base + (long)krow0 * ld + 2 * ch;
    safe = safe_;
    soff = krow0 * LDS_ + 2 * ch;
    nbytes = clamp_bytes(n_valid - 2 * ch);
  }
  __device__ __forceinline__ void load(double* s, long k0, long k_valid) const {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int bytes = (krow0 + RPI * i < k_valid) ? nbytes : 0;
      const double* src = bytes ? (g + (k0 + RPI * i) * ld) : safe;
      cp_async16(s + soff + i * RPI * LDS_, src, bytes);
    }
  }
};

// One k-block of MMAs for a warp tile WM x WN.  sA: K-major [.. rows][KMAJ_LD] positioned at the warp's
// first row; sB: either K-major (positioned at the warp's first column-row) or N-major (positioned at the
// warp's first column).  `ascale` multiplies the A fragments (segment weight of the weighted SYRK).
template <int WM, int WN, bool B_KMAJOR, int LDB_S, bool SCALE_A>
__device__ __forceinline__ void mma_kblock(const double* __restrict__ sA, const double* __restrict__ sB,
                                           double (&acc)[WM / 8][WN / 8][2], int g, int q, double ascale) {
#pragma unroll
  for (int kk = 0; kk < BK / 4; ++kk) {
    double af[WM / 8], bf[WN / 8];
#pragma unroll
    for (int i = 0; i < WM / 8; ++i) {
      af[i] = sA[(i * 8 + g) * KMAJ_LD + kk * 4 + q];
      if (SCALE_A) af[i] *= ascale;
    }
#pragma unroll
    for (int j = 0; j < WN / 8; ++j)
      bf[j] = B_KMAJOR ? sB[(j * 8 + g) * LDB_S + kk * 4 + q] : sB[(kk * 4 + q) * LDB_S + j * 8 + g];
#pragma unroll
    for (int i = 0; i < WM / 8; ++i)
#pragma unroll
      for (int j = 0; j < WN / 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic CTA-wide sum of one double per thread; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* red /* >= 8 doubles of smem */) {
  v = warp_sum(v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
  }
  return s;
}

}  // namespace gpcsd
