// Persistent, warp-specialised FP64 DMMA GEMM / weighted SYRK for sm_100a with TMA-staged tiles.
//
//   warp 8      : producer -- one elected lane issues cp.async.bulk.tensor (TMA) loads into a ring of
//                 128B-swizzled shared-memory stages, signalled through mbarriers (full / empty)
//   warps 0..7  : consumers -- LDS + DMMA.8x8x4 on a 128x128 CTA tile (warp tile 64x32), no CTA-wide barrier
//                 in the main loop; epilogue of tile i overlaps the producer's prefetch of tile i+1
//   grid        : one CTA per SM, static round-robin over tiles (m-tiles fastest so the CTAs sharing a B
//                 column panel run together and the panel is fetched from HBM once)
//
// Shared-memory layout (SWIZZLE_128B, box inner extent = 16 doubles = 128 B):
//   K-major tile [rows][16 k]:  element (r, k) at  r*128 + (((k>>1) ^ (r&7)) << 4) + (k&1)*8
//   N-major tile = 8 boxes [16 k][16 n]: element (k, n) at (n>>4)*2048 + k*128 + ((((n&15)>>1) ^ (k&7)) << 4) + (n&1)*8
// Fragment mapping (lane = 4g + q), chosen so that every shared-memory wavefront is conflict-free:
//   * MMA k-slot q of k-step kk reads memory k = 8*(kk>>1) + 2q + (kk&1)   (any k permutation is legal as long
//     as A and B agree): a K-major operand then needs ONE LDS.128 per row-tile for two k-steps (kk even -> .x,
//     kk odd -> .y), an N-major operand one LDS.64 per (column-tile, k-step);
//   * MMA row/column index g of a K-major operand reads memory row 8*tile + perm(g), perm(g) = (g>>1) + 4*(g&1),
//     so the 8 lanes of each LDS.128 quarter-warp touch rows {0,4} / {1,5} / {2,6} / {3,7} (mod 8) -> 8 distinct
//     16-byte chunks after the XOR swizzle.
#include <cuda.h>

#include "common.h"
#include "dmma_gemm.cuh"

namespace gpcsd {

#ifndef GPCSD_TMA_CONSUMERS
#define GPCSD_TMA_CONSUMERS 8
#endif
constexpr int T_BM = 128, T_BN = 128, T_WN = 32;
constexpr int T_STAGES = 6;
constexpr int T_CONSUMERS = GPCSD_TMA_CONSUMERS;     // consumer warps: 8 (64x32 warp tiles) or 16 (32x32)
constexpr int T_WARPS_M = T_CONSUMERS / 4;           // warp grid T_WARPS_M x 4
constexpr int T_WM = T_BM / T_WARPS_M;
constexpr int T_MT = T_WM / 8;                       // 8-row MMA tiles per warp
constexpr int T_PRODUCER_WARPS = 4;                  // one warpgroup (setmaxnreg works per warpgroup); warp 0 issues TMA
constexpr int T_THREADS = 32 * (T_CONSUMERS + T_PRODUCER_WARPS);
constexpr int T_REGS_PRODUCER = 40, T_REGS_CONSUMER = 232;   // 128*40 + 256*232 = 64512 <= 65536 registers
constexpr int T_TILE_BYTES = T_BM * BK * 8;          // 16 KiB per operand per stage
constexpr int T_STAGE_BYTES = 2 * T_TILE_BYTES;
constexpr size_t T_SMEM_BYTES = (size_t)T_STAGES * T_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int perm8(int g) { return (g >> 1) + 4 * (g & 1); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }

// one 16-deep k-block of a 64x32 warp tile.
//   aBase/bBase: shared-memory byte addresses of the stage's A / B tiles
//   B_KMAJOR: B tile stored [n rows][16 k] (NT / SYRK) else 8 boxes [16 k][16 n] (NN)
// NTW = 8-column MMA tiles per warp (warp tile 64 x 8*NTW; CTA tile 128 x 32*NTW)
template <bool B_KMAJOR, bool SCALE_A, int NTW = 4, int MT = T_MT>
__device__ __forceinline__ void tma_mma_kblock(uint32_t aBase, uint32_t bBase, double (&acc)[MT][NTW][2], int wm, int wn, int g,
                                               int q, double ascale) {
  static_assert(MT % 4 == 0, "row tiles are processed four at a time");
  const int pg = perm8(g);
  // (row & 7) == pg for every row tile (rows advance by 8), so the swizzle XOR term is loop-invariant
  const uint32_t aRow = aBase + (wm * 8 * MT + pg) * 128;
  constexpr int WN = 8 * NTW;
  const uint32_t bRow = bBase + (wn * WN + pg) * 128;
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // k-step pairs (2h, 2h+1)
    const uint32_t kch = (uint32_t)(((4 * h + q) ^ pg) << 4);
    double b0[NTW], b1[NTW];
    if (B_KMAJOR) {
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const double2 t = lds128(bRow + j * 1024 + kch);
        b0[j] = t.x;
        b1[j] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const int n = wn * WN + 8 * j + g;
        const uint32_t box = bBase + (n >> 4) * 2048 + (n & 1) * 8;
        const int c = (n & 15) >> 1;
        const int k0 = 8 * h + 2 * q, k1 = k0 + 1;
        b0[j] = lds64(box + k0 * 128 + ((c ^ (k0 & 7)) << 4));
        b1[j] = lds64(box + k1 * 128 + ((c ^ (k1 & 7)) << 4));
      }
    }
#pragma unroll
    for (int ih = 0; ih < MT / 4; ++ih) {  // 4 row tiles at a time: keeps only 4 A fragments live
      double2 af[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        af[i] = lds128(aRow + (4 * ih + i) * 1024 + kch);
        if (SCALE_A) {
          af[i].x *= ascale;
          af[i].y *= ascale;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) dmma884(acc[4 * ih + i][j][0], acc[4 * ih + i][j][1], af[i].x, b0[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) dmma884(acc[4 * ih + i][j][0], acc[4 * ih + i][j][1], af[i].y, b1[j]);
    }
  }
}

struct TmaGemmArgs {
  double* C;
  long ldc, sC;
  int M, N, K, batch;
  int m_tiles, n_tiles;
  int a_batched;      // 0: A shared by all batches
  int a_div;          // A operand of batch b is slab b / a_div of the tensor map (restart-batched projection: a_div = nx)
  int grp, ngrp;      // EPI_QUAD: batches are reduced in ngrp groups of grp consecutive batches (one group per restart)
  const double* rD;   // EPI_QUAD
  long ldrd;
  double* partials;   // [ngrp][gridDim.x][2]
};

constexpr int TEPI_STORE = 0, TEPI_QUAD = 1;

__device__ __forceinline__ void pipeline_setup(uint8_t*& tiles, uint64_t*& full, uint64_t*& empty) {
  extern __shared__ uint8_t raw_smem[];
  const uintptr_t base = (reinterpret_cast<uintptr_t>(raw_smem) + 1023) & ~uintptr_t(1023);
  tiles = reinterpret_cast<uint8_t*>(base);
  full = reinterpret_cast<uint64_t*>(tiles + (size_t)T_STAGES * T_STAGE_BYTES);
  empty = full + T_STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < T_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, T_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
}

// C_b = A_b * op(B_b): persistent tiles, TMA producer + 8 DMMA consumer warps.  TBM = rows of the CTA tile: 128, or 64 for
// operands whose row count pads badly to 128 (the 192-order halves of a Neuropixels spatial factor: 256 -> 192 rows computed).
template <bool BT, int EPI, int NTW, int TBM>
__global__ void __launch_bounds__(T_THREADS, 1)
    tma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TmaGemmArgs p) {
  constexpr int BN = 32 * NTW, WN = 8 * NTW;     // CTA tile TBM x BN, warp tile TBM/2 x WN
  constexpr int WM = TBM / T_WARPS_M, MT = WM / 8;
  uint8_t* tiles;
  uint64_t *full, *empty;
  pipeline_setup(tiles, full, empty);
  const int lane = threadIdx.x & 31;
  const int nkb = (p.K + BK - 1) / BK;
  const long ntiles = (long)p.m_tiles * p.n_tiles * p.batch;

  if (threadIdx.x < 32 * T_PRODUCER_WARPS) {
    // ------------------------------------------------------------------ producer warpgroup
    reg_dealloc<T_REGS_PRODUCER>();
    if (threadIdx.x == 0) {
      uint32_t it = 0;
      for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int mt = (int)(t % p.m_tiles);
        const long r = t / p.m_tiles;
        const int nt_ = (int)(r % p.n_tiles), b = (int)(r / p.n_tiles);
        const int m0 = mt * TBM, n0 = nt_ * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % T_STAGES;
          mbar_wait(empty + s, ((it / T_STAGES) & 1) ^ 1);
          mbar_expect_tx(full + s, TBM * BK * 8 + BN * BK * 8);
          uint8_t* sa = tiles + (size_t)s * T_STAGE_BYTES;
          uint8_t* sb = sa + T_TILE_BYTES;
          tma_load_3d(sa, &tmA, full + s, kb * BK, m0, p.a_batched ? b / p.a_div : 0);
          if (BT) {
            tma_load_3d(sb, &tmB, full + s, kb * BK, n0, b);       // box 16 k x BN rows
          } else {
#pragma unroll
            for (int x = 0; x < BN / 16; ++x) tma_load_3d(sb + x * 2048, &tmB, full + s, n0 + 16 * x, kb * BK, b);
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers (2 warpgroups)
  reg_alloc<T_REGS_CONSUMER>();
  const int warp = (threadIdx.x >> 5) - T_PRODUCER_WARPS;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp % T_WARPS_M, wn = warp / T_WARPS_M;
  const int pg = perm8(g);
  double quad = 0.0, bsq = 0.0;
  uint32_t it = 0;
  // EPI_QUAD with several reduction groups (restarts): a CTA's tiles visit the groups in increasing order, so the running
  // sums are flushed into the group's slot whenever the group changes; slots of groups this CTA never visits stay zero.
  __shared__ double red[2][T_CONSUMERS];
  int cur_grp = -1;
  auto flush = [&](int grp_id) {
    const double q0 = warp_sum(quad), b0 = warp_sum(bsq);
    if (lane == 0) {
      red[0][warp] = q0;
      red[1][warp] = b0;
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * T_CONSUMERS) : "memory");
    if (warp == 0 && lane == 0) {
      double s0 = 0.0, s1 = 0.0;
      for (int w = 0; w < T_CONSUMERS; ++w) {
        s0 += red[0][w];
        s1 += red[1][w];
      }
      p.partials[2 * ((long)grp_id * gridDim.x + blockIdx.x)] = s0;
      p.partials[2 * ((long)grp_id * gridDim.x + blockIdx.x) + 1] = s1;
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * T_CONSUMERS) : "memory");
    quad = bsq = 0.0;
  };
  if (EPI == TEPI_QUAD) {
    for (int gi = threadIdx.x - 32 * T_PRODUCER_WARPS; gi < p.ngrp; gi += 32 * T_CONSUMERS) {
      p.partials[2 * ((long)gi * gridDim.x + blockIdx.x)] = 0.0;
      p.partials[2 * ((long)gi * gridDim.x + blockIdx.x) + 1] = 0.0;
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * T_CONSUMERS) : "memory");
  }
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int mt = (int)(t % p.m_tiles);
    const long r = t / p.m_tiles;
    const int nt_ = (int)(r % p.n_tiles), b = (int)(r / p.n_tiles);
    const int m0 = mt * TBM, n0 = nt_ * BN;
    if (EPI == TEPI_QUAD) {
      const int gnow = b / p.grp;
      if (gnow != cur_grp) {
        if (cur_grp >= 0) flush(cur_grp);
        cur_grp = gnow;
      }
    }
    double acc[MT][NTW][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTW; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it % T_STAGES;
      mbar_wait(full + s, (it / T_STAGES) & 1);
      const uint32_t sa = smem_u32(tiles + (size_t)s * T_STAGE_BYTES);
      tma_mma_kblock<BT, false, NTW, MT>(sa, sa + T_TILE_BYTES, acc, wm, wn, g, q, 1.0);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    // epilogue (overlaps the producer's prefetch of the next tile)
    double* C = p.C + (long)b * p.sC;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const int m = m0 + wm * WM + 8 * i + pg;
      if (m >= p.M) continue;
      double rr = 1.0;
      if (EPI == TEPI_QUAD) rr = __ldg(p.rD + (long)b * p.ldrd + m);
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        double v0 = acc[i][j][0], v1 = acc[i][j][1];
        if (EPI == TEPI_QUAD) {
          v0 *= rr;
          v1 *= rr;
        }
        if (BT) {
          const long na = (long)n0 + wn * WN + 8 * j + q, nb = na + 4;
          if (na < p.N) C[(long)m * p.ldc + na] = v0;
          if (nb < p.N) C[(long)m * p.ldc + nb] = v1;
        } else {
          const long n = (long)n0 + wn * WN + 8 * j + 2 * q;
          if (n >= p.N) continue;
          const bool two = (n + 1 < p.N);
          if (EPI == TEPI_QUAD) {
            quad += acc[i][j][0] * v0;
            bsq += v0 * v0;
            if (two) {
              quad += acc[i][j][1] * v1;
              bsq += v1 * v1;
            }
          }
          double* dst = C + (long)m * p.ldc + n;
          if (two)
            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
          else
            *dst = v0;
        }
      }
    }
  }
  if (EPI == TEPI_QUAD && cur_grp >= 0) flush(cur_grp);     // consumer-only reduction, fixed order -> deterministic
}

// ------------------------------------------------------------------------------------------------
// segment-weighted split-K SYRK (lower-triangular 128x128 tiles), X viewed as the 3-D tensor {k, row, seg}
// ------------------------------------------------------------------------------------------------
struct TmaSyrkArgs {
  const double* w;
  int M, nseg, seglen;
  int kbps;       // k-blocks per segment
  long total_kb;  // nseg * kbps
  int nsplit, tiles_1d;
  int seg_middle; // tensor-map dim order {k, seg, row} instead of {k, row, seg}
  double* ws;     // [restart][nsplit][ntiles][128*128] in fragment order
  long w_stride;  // restart strides (blockIdx.z = restart; the tensor map's 4th dimension)
  long ws_stride;
};

template <int TBM>     // tile edge: 128, or 64 where 128-tiles would compute much more than the lower triangle (syrk_plan)
__global__ void __launch_bounds__(T_THREADS, 1)
    tma_wsyrk_kernel(const __grid_constant__ CUtensorMap tmX, TmaSyrkArgs p) {
  constexpr int WM = TBM / T_WARPS_M, MT = WM / 8, NTW = TBM / 32;
  uint8_t* tiles;
  uint64_t *full, *empty;
  pipeline_setup(tiles, full, empty);
  const int lane = threadIdx.x & 31;
  int t = blockIdx.x, tm = 0;
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  const int tn = t - tm * (tm + 1) / 2;
  const bool diag = (tm == tn);
  const int split = blockIdx.y;
  const long f0 = p.total_kb * split / p.nsplit, f1 = p.total_kb * (split + 1) / p.nsplit;
  const int nkb = (int)(f1 - f0);

  if (threadIdx.x < 32 * T_PRODUCER_WARPS) {
    reg_dealloc<T_REGS_PRODUCER>();
    if (threadIdx.x == 0) {
      for (int it = 0; it < nkb; ++it) {
        const int s = it % T_STAGES;
        const long f = f0 + it;
        const int seg = (int)(f / p.kbps);
        const int k0 = (int)(f - (long)seg * p.kbps) * BK;
        mbar_wait(empty + s, ((it / T_STAGES) & 1) ^ 1);
        mbar_expect_tx(full + s, (diag ? 1 : 2) * TBM * BK * 8);
        uint8_t* sa = tiles + (size_t)s * T_STAGE_BYTES;
        const int rr = blockIdx.z;
        if (p.seg_middle) {
          tma_load_4d(sa, &tmX, full + s, k0, seg, tm * TBM, rr);
          if (!diag) tma_load_4d(sa + T_TILE_BYTES, &tmX, full + s, k0, seg, tn * TBM, rr);
        } else {
          tma_load_4d(sa, &tmX, full + s, k0, tm * TBM, seg, rr);
          if (!diag) tma_load_4d(sa + T_TILE_BYTES, &tmX, full + s, k0, tn * TBM, seg, rr);
        }
      }
    }
    return;
  }
  reg_alloc<T_REGS_CONSUMER>();
  const int warp = (threadIdx.x >> 5) - T_PRODUCER_WARPS;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp % T_WARPS_M, wn = warp / T_WARPS_M;
  double acc[MT][NTW][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int it = 0; it < nkb; ++it) {
    const int s = it % T_STAGES;
    const double wgt = p.w ? __ldg(p.w + (long)blockIdx.z * p.w_stride + (f0 + it) / p.kbps) : 1.0;
    mbar_wait(full + s, (it / T_STAGES) & 1);
    const uint32_t sa = smem_u32(tiles + (size_t)s * T_STAGE_BYTES);
    if (p.w)
      tma_mma_kblock<true, true, NTW, MT>(sa, diag ? sa : sa + T_TILE_BYTES, acc, wm, wn, g, q, wgt);
    else
      tma_mma_kblock<true, false, NTW, MT>(sa, diag ? sa : sa + T_TILE_BYTES, acc, wm, wn, g, q, 1.0);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
  // partial tile in fragment order: fully coalesced double2 stores
  const long ntiles = (long)p.tiles_1d * (p.tiles_1d + 1) / 2;
  double* out = p.ws + (long)blockIdx.z * p.ws_stride + ((long)split * ntiles + t) * (TBM * TBM);
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTW; ++j)
      *reinterpret_cast<double2*>(out + ((((warp * MT + i) * NTW + j) * 32 + lane) << 1)) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// sum the split-K partials in fixed order, un-permute the fragment layout, mirror the upper triangle.
// Block = 64 consecutive fragment-order elements x 4 split groups (group s sums splits s, s + 4, ...; the four group sums
// are added in order through shared memory): four times the loads in flight of a one-thread-per-element sum, which was
// latency-bound at 14 % of HBM bandwidth.
__global__ void __launch_bounds__(256) tma_wsyrk_reduce_kernel(const double* __restrict__ ws, int nsplit, int tiles_1d, int M,
                                                               double* __restrict__ C, long ldc, long ws_stride, long c_stride,
                                                               int bm) {
  const int te = bm * bm, mt_ = bm / (8 * T_WARPS_M), ntw = bm / 32;     // tile elements, row tiles / column tiles per warp
  __shared__ double part[4][64];
  ws += (long)blockIdx.z * ws_stride;
  C += (long)blockIdx.z * c_stride;
  const long ntiles = (long)tiles_1d * (tiles_1d + 1) / 2;
  const int t = blockIdx.y;
  int tm = 0;
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  const int tn = t - tm * (tm + 1) / 2;
  const int el = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + el;                   // fragment-order element index
  double s0 = 0.0, s1 = 0.0;
  int sp = grp;
  for (; sp + 4 < nsplit; sp += 8) {
    s0 += ws[((long)sp * ntiles + t) * te + e];
    s1 += ws[((long)(sp + 4) * ntiles + t) * te + e];
  }
  if (sp < nsplit) s0 += ws[((long)sp * ntiles + t) * te + e];
  part[grp][el] = s0 + s1;
  __syncthreads();
  if (grp != 0) return;
  const double s = (part[0][el] + part[1][el]) + (part[2][el] + part[3][el]);
  const int v = e & 1, lane = (e >> 1) & 31, j = (e >> 6) % ntw, i = ((e >> 6) / ntw) % mt_, warp = (e >> 6) / (ntw * mt_);
  const int g = lane >> 2, q = lane & 3, wm = warp % T_WARPS_M, wn = warp / T_WARPS_M;
  const int r = wm * 8 * mt_ + 8 * i + perm8(g), c = wn * 8 * ntw + 8 * j + q + 4 * v;
  const int m = tm * bm + r, n = tn * bm + c;
  if (m >= M || n >= M) return;
  C[(long)m * ldc + n] = s;
  if (tm != tn) C[(long)n * ldc + m] = s;
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map encoding through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D FP64 tensor map: dims {d0 (contiguous), d1, d2}, element strides {s1, s2}, box {b0, b1, b2}; 128B swizzle,
// zero fill for out-of-range box elements (this is what handles every M / N / K / segment edge).
static int make_map3(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                     uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return gp_fail("cuTensorMapEncodeTiled entry point not available");
  if (d1 == 0) d1 = 1;
  if (d2 == 0) d2 = 1;
  if (s1 == 0) s1 = (d0 + 1) & ~1ull;                       // unused stride (d1 == 1): any legal value
  if (s2 == 0) s2 = s1 * d1;                                 // unused stride (d2 == 1)
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 8, s2 * 8};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu strides %llu %llu", (int)r,
             (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1,
             (unsigned long long)s2);
    return gp_fail(buf);
  }
  return 0;
}

// 4-D variant (the SYRK's {k, row | seg, seg | row, restart} view of the trial data)
static int make_map4(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t s1, uint64_t s2,
                     uint64_t s3, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return gp_fail("cuTensorMapEncodeTiled entry point not available");
  if (d1 == 0) d1 = 1;
  if (d2 == 0) d2 = 1;
  if (d3 == 0) d3 = 1;
  if (s1 == 0) s1 = (d0 + 1) & ~1ull;
  if (s2 == 0) s2 = s1 * d1;
  if (s3 == 0) s3 = (s2 * d2 > s1 * d1 ? s2 * d2 : s1 * d1);
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 8, s2 * 8, s3 * 8};
  cuuint32_t box[4] = {b0, b1, b2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (4-D) failed (%d): dims %llu %llu %llu %llu", (int)r, (unsigned long long)d0,
             (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3);
    return gp_fail(buf);
  }
  return 0;
}

template <bool BT, int EPI, int NTW, int TBM>
static int launch_tma_gemm(const CUtensorMap& tA, const CUtensorMap& tB, TmaGemmArgs& p, cudaStream_t st) {
  auto kern = tma_gemm_kernel<BT, EPI, NTW, TBM>;
  static int attr[GP_MAX_DEVICES];
  if (gp_first_use_on_device(attr)) {
    GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES));
  }
  const long ntiles = (long)p.m_tiles * p.n_tiles * p.batch;
  const long grid = ntiles < gp_num_sms() ? ntiles : gp_num_sms();
  kern<<<(unsigned)grid, T_THREADS, T_SMEM_BYTES, st>>>(tA, tB, p);
  GP_CUDA(cudaGetLastError());
  return 0;
}

// Tile width: 128, 96 or 64 columns.  Persistent CTAs run ceil(tiles / SMs) rounds of cost ~BN each, so the
// narrower tiles win when they shave the ragged last round or the padded last column tile
// (configs[1]: N = 2000 -> 16 x 128 = 2048 columns in 11 rounds vs 21 x 96 = 2016 columns in 14 rounds: -4.5 %).
// Tile height: 64-row tiles where they cut the row padding by more than ~13 % (M = 192: 256 -> 192 rows; M = 250 pads to 256
// either way and keeps the more efficient 128-row tiles).
static int pick_bm(int M) {
  const long p128 = (long)((M + 127) / 128) * 128, p64 = (long)((M + 63) / 64) * 64;
  return (p128 * 100 > p64 * 115) ? 64 : 128;
}

static int pick_ntw(int M, int N, int batch) {
  const int bm = pick_bm(M);
  const long mt = (M + bm - 1) / bm, sms = gp_num_sms();
  int best = 4;
  double best_cost = 1e300;
  for (int ntw = 4; ntw >= 2; --ntw) {
    const int bn = 32 * ntw;
    const long tiles = mt * ((N + bn - 1) / bn) * batch;
    const long rounds = (tiles + sms - 1) / sms;
    const double cost = (double)rounds * (bn + 6.0);      // +6: per-tile prologue/epilogue overhead in column units
    if (cost < best_cost * 0.995) {
      best_cost = cost;
      best = ntw;
    }
  }
  return best;
}

int tma_gemm_ctas(int M, int N, int batch) {
  const int bn = 32 * pick_ntw(M, N, batch), bm = pick_bm(M);
  const long ntiles = (long)((M + bm - 1) / bm) * ((N + bn - 1) / bn) * batch;
  return (int)(ntiles < gp_num_sms() ? ntiles : gp_num_sms());
}

template <bool BT, int EPI>
static int launch_tma_gemm_ntw(int ntw, int bm, const CUtensorMap& tA, const CUtensorMap& tB, TmaGemmArgs& p, cudaStream_t st) {
  if (bm == 64) {
    if (ntw == 4) return launch_tma_gemm<BT, EPI, 4, 64>(tA, tB, p, st);
    if (ntw == 3) return launch_tma_gemm<BT, EPI, 3, 64>(tA, tB, p, st);
    return launch_tma_gemm<BT, EPI, 2, 64>(tA, tB, p, st);
  }
  if (ntw == 4) return launch_tma_gemm<BT, EPI, 4, 128>(tA, tB, p, st);
  if (ntw == 3) return launch_tma_gemm<BT, EPI, 3, 128>(tA, tB, p, st);
  return launch_tma_gemm<BT, EPI, 2, 128>(tA, tB, p, st);
}

// C_b = A_b op(B_b); epi_quad != 0 fuses the /D + quadratic-form epilogue.  Returns 0 on success.
int tma_gemm(int transB, int M, int N, int K, const double* A, long lda, long sA, const double* B, long ldb, long sB,
             double* C, long ldc, long sC, int batch, const double* rD, long ldrd, double* partials, int epi_quad,
             cudaStream_t st, int a_div, int grp) {
  CUtensorMap tA, tB;
  const bool a_batched = (sA != 0 && batch > 1);
  if (a_div < 1) a_div = 1;
  if (grp < 1) grp = batch;
  const int ntw = pick_ntw(M, N, batch), bn = 32 * ntw, bm = pick_bm(M);
  if (int e = make_map3(&tA, A, K, M, a_batched ? (batch + a_div - 1) / a_div : 1, lda, a_batched ? sA : 0, BK, bm, 1)) return e;
  const long sBe = (batch > 1) ? sB : 0;
  if (transB) {
    if (int e = make_map3(&tB, B, K, N, batch, ldb, sBe, BK, bn, 1)) return e;
  } else {
    if (int e = make_map3(&tB, B, N, K, batch, ldb, sBe, 16, BK, 1)) return e;
  }
  TmaGemmArgs p{};
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.m_tiles = (M + bm - 1) / bm;
  p.n_tiles = (N + bn - 1) / bn;
  p.a_batched = a_batched ? 1 : 0;
  p.a_div = a_div; p.grp = grp; p.ngrp = (batch + grp - 1) / grp;
  p.rD = rD; p.ldrd = ldrd; p.partials = partials;
  if (epi_quad) return launch_tma_gemm_ntw<false, TEPI_QUAD>(ntw, bm, tA, tB, p, st);
  if (transB) return launch_tma_gemm_ntw<true, TEPI_STORE>(ntw, bm, tA, tB, p, st);
  return launch_tma_gemm_ntw<false, TEPI_STORE>(ntw, bm, tA, tB, p, st);
}

template <int TBM>
static int launch_tma_wsyrk(const CUtensorMap& tX, const TmaSyrkArgs& p, dim3 grid, cudaStream_t st) {
  static int attr[GP_MAX_DEVICES];
  if (gp_first_use_on_device(attr)) {
    GP_CUDA(cudaFuncSetAttribute(tma_wsyrk_kernel<TBM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES));
  }
  tma_wsyrk_kernel<TBM><<<grid, T_THREADS, T_SMEM_BYTES, st>>>(tX, p);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int tma_wsyrk(int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, const double* w, double* C,
              long ldc, double* ws, int nsplit, int tiles_1d, int kbps, long total_kb, cudaStream_t st, int R, long x_stride,
              long w_stride, long c_stride, int bm) {
  CUtensorMap tX;
  if (R < 1) R = 1;
  // X[restart][m][seg][k] = X + restart*x_stride + m*row_stride + seg*seg_stride + k.  Outer tensor dims are ordered by
  // increasing stride (seg_middle: {k, seg, row, restart}; else {k, row, seg, restart}); the box is 16 k x 128 rows x 1 segment
  // x 1 restart either way, so the shared-memory image is always [128 rows][16 k].
  const bool seg_middle = (nseg > 1) && (seg_stride < row_stride);
  const long xs = R > 1 ? x_stride : 0;
  if (seg_middle) {
    if (int e = make_map4(&tX, X, seglen, nseg, M, R, seg_stride, row_stride, xs, BK, 1, bm)) return e;
  } else {
    if (int e = make_map4(&tX, X, seglen, M, nseg, R, row_stride, nseg > 1 ? seg_stride : 0, xs, BK, bm, 1)) return e;
  }
  TmaSyrkArgs p{};
  p.w = w; p.M = M; p.nseg = nseg; p.seglen = seglen;
  p.kbps = kbps; p.total_kb = total_kb; p.nsplit = nsplit; p.tiles_1d = tiles_1d; p.ws = ws;
  p.seg_middle = seg_middle ? 1 : 0;
  const long ntiles = (long)tiles_1d * (tiles_1d + 1) / 2;
  p.w_stride = w_stride;
  p.ws_stride = (long)nsplit * ntiles * ((long)bm * bm);
  const dim3 grid((unsigned)ntiles, (unsigned)nsplit, (unsigned)R);
  if (int e = (bm == 64) ? launch_tma_wsyrk<64>(tX, p, grid, st) : launch_tma_wsyrk<128>(tX, p, grid, st)) return e;
  tma_wsyrk_reduce_kernel<<<dim3(bm * bm / 64, (unsigned)ntiles, (unsigned)R), 256, 0, st>>>(ws, nsplit, tiles_1d, M, C, ldc,
                                                                                           p.ws_stride, c_stride, bm);
  GP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gpcsd
