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

// K-major tile: ROWS x BK doubles from a row-major matrix (row stride ld, K contiguous).
//   rows_valid: number of in-range rows from `g` on; k_valid: in-range k from this block's k0 on.
template <int ROWS>
__device__ __forceinline__ void load_kmajor_tile(double* s, const double* g, long ld, int rows_valid, long k_valid,
                                                 const double* safe) {
  constexpr int CHUNKS = ROWS * (BK / 2);
#pragma unroll
  for (int c = threadIdx.x; c < CHUNKS; c += NTHREADS) {
    int row = c >> 3, ch = c & 7;
    int bytes = (row < rows_valid) ? clamp_bytes(k_valid - 2 * ch) : 0;
    const double* src = bytes ? (g + (long)row * ld + 2 * ch) : safe;
    cp_async16(s + row * KMAJ_LD + 2 * ch, src, bytes);
  }
}

// N-major tile: BK x COLS doubles from a row-major K x N matrix (row stride ld, N contiguous).
template <int COLS>
__device__ __forceinline__ void load_nmajor_tile(double* s, const double* g, long ld, long k_valid, long n_valid,
                                                 const double* safe) {
  constexpr int CPR = COLS / 2;  // 16-byte chunks per row
  constexpr int CHUNKS = BK * CPR;
  constexpr int LDS_ = COLS + 4;
#pragma unroll
  for (int c = threadIdx.x; c < CHUNKS; c += NTHREADS) {
    int row = c / CPR, ch = c % CPR;
    int bytes = (row < k_valid) ? clamp_bytes(n_valid - 2 * ch) : 0;
    const double* src = bytes ? (g + (long)row * ld + 2 * ch) : safe;
    cp_async16(s + row * LDS_ + 2 * ch, src, bytes);
  }
}

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
