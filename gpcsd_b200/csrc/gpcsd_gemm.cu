// ABI entry points of the GEMM-shaped stages (gpcsd_dgemm, gpcsd_project_quad, gpcsd_wsyrk) and the small-M
// kernels: cp.async DMMA GEMM for M <= 32 and the register-only small-M SYRK.  Everything with M > 32 is
// dispatched to the persistent TMA + mbarrier kernels in gpcsd_tma.cu.  See include/gpcsd_b200.h.
#include <stdlib.h>

#include "common.h"
#include "dmma_gemm.cuh"

namespace gpcsd {

// gpcsd_tma.cu: persistent TMA + mbarrier warp-specialised kernels (M > 32)
int tma_gemm(int transB, int M, int N, int K, const double* A, long lda, long sA, const double* B, long ldb, long sB,
             double* C, long ldc, long sC, int batch, const double* rD, long ldrd, double* partials, int epi_quad,
             cudaStream_t st, int a_div = 1, int grp = 0);
int tma_gemm_ctas(int M, int N, int batch);
int tma_wsyrk(int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, const double* w, double* C,
              long ldc, double* ws, int nsplit, int tiles_1d, int kbps, long total_kb, cudaStream_t st, int R, long x_stride,
              long w_stride, long c_stride, int bm);

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  long lda, ldb, ldc;
  long sA, sB, sC;  // batch strides
  int M, N, K;
  int m_tiles;
  int a_div = 1;     // A operand of batch b: A + (b / a_div) * sA   (restart-batched projection: a_div = nx)
  int grp = 0;       // quad epilogue: batches per reduction group (0: one group)
  // quad epilogue
  const double* rD;  // rD[batch*ldrd + m]
  long ldrd;
  double* partials;  // [num_ctas][2]
};

constexpr int EPI_STORE = 0;
constexpr int EPI_QUAD = 1;

template <int BM, int BN, bool BT, int STAGES>
struct SmemLayout {
  static constexpr int A_STAGE = BM * KMAJ_LD;
  static constexpr int LDB_S = BT ? KMAJ_LD : (BN + 4);
  static constexpr int B_STAGE = BT ? BN * KMAJ_LD : BK * (BN + 4);
  static constexpr size_t BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + 64;
};

// C_b = A_b * op(B_b) for M <= 32 (one 32-row tile).  grid.x = n_tiles, grid.y = batch.
template <int BM, int BN, int WM, int WN, int STAGES, bool BT, int EPI, int MINB>
__global__ void __launch_bounds__(NTHREADS, MINB) dmma_gemm_kernel(GemmArgs p) {
  using L = SmemLayout<BM, BN, BT, STAGES>;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * L::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  constexpr int WARPS_M = BM / WM;
  static_assert((BM / WM) * (BN / WN) == NTHREADS / 32, "warp tiling must use 8 warps");
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;

  const int mt = blockIdx.x % p.m_tiles, nt_ = blockIdx.x / p.m_tiles;
  const int m0 = mt * BM, n0 = nt_ * BN;
  const int b = blockIdx.y;
  const double* A = p.A + (long)(b / p.a_div) * p.sA + (long)m0 * p.lda;
  const double* B = p.B + (long)b * p.sB + (BT ? (long)n0 * p.ldb : (long)n0);
  const int rowsA = p.M - m0;
  const long colsB = (long)p.N - n0;
  const int nkb = (p.K + BK - 1) / BK;

  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  KMajorLoader<BM> ldA;
  ldA.init(A, p.lda, rowsA, p.A);
  KMajorLoader<BN> ldBk;
  NMajorLoader<BN> ldBn;
  if (BT)
    ldBk.init(B, p.ldb, (int)(colsB > BN ? BN : colsB), p.B);
  else
    ldBn.init(B, p.ldb, colsB, p.B);
  auto load_stage = [&](int st, int kb) {
    const long k0 = (long)kb * BK;
    ldA.load(sA + st * L::A_STAGE, k0, p.K - k0);
    if (BT)
      ldBk.load(sB + st * L::B_STAGE, k0, p.K - k0);
    else
      ldBn.load(sB + st * L::B_STAGE, k0, p.K - k0);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kb + STAGES - 1;
    if (nxt < nkb) load_stage(nxt % STAGES, nxt);
    cp_async_commit();
    const int st = kb % STAGES;
    const double* a = sA + st * L::A_STAGE + (wm * WM) * KMAJ_LD;
    const double* bb = sB + st * L::B_STAGE + (BT ? (wn * WN) * KMAJ_LD : (wn * WN));
    mma_kblock<WM, WN, BT, L::LDB_S, false>(a, bb, acc, g, q, 1.0);
  }
  cp_async_wait<0>();

  // ---- epilogue
  double* C = p.C + (long)b * p.sC;
  double quad = 0.0, bsq = 0.0;
#pragma unroll
  for (int i = 0; i < WM / 8; ++i) {
    const int m = m0 + wm * WM + i * 8 + g;
    if (m >= p.M) continue;
    double r = 1.0;
    if (EPI == EPI_QUAD) r = __ldg(p.rD + (long)b * p.ldrd + m);
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) {
      const long n = (long)n0 + wn * WN + j * 8 + 2 * q;
      if (n >= p.N) continue;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      const bool two = (n + 1 < p.N);
      if (EPI == EPI_QUAD) {
        const double u0 = v0 * r, u1 = v1 * r;
        quad += v0 * u0;
        bsq += u0 * u0;
        if (two) {
          quad += v1 * u1;
          bsq += u1 * u1;
        }
        v0 = u0;
        v1 = u1;
      }
      double* dst = C + (long)m * p.ldc + n;
      if (two)
        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      else
        *dst = v0;
    }
  }
  if (EPI == EPI_QUAD) {
    __shared__ double red[16];
    const double s0 = block_sum(quad, red);
    const double s1 = block_sum(bsq, red + 8);
    if (tid == 0) {
      const long cta = (long)blockIdx.y * gridDim.x + blockIdx.x;
      p.partials[2 * cta] = s0;
      p.partials[2 * cta + 1] = s1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// C_b = A_b (M x K) * B_b (K x N) for M <= 32 AND K <= 32 (Z = Qs^T Y of the 1-D model: 24 x 24 times 24 x (nt * trials),
// 3 flop per byte -> HBM-bound).  No shared memory: the whole A operand lives in DMMA fragments for the life of the warp
// (MT x KT doubles per lane); B fragments come straight from global memory with 16-byte loads and C leaves with 16-byte
// stores.  One warp step covers 16 columns: MMA column g of tile 0 / 1 is memory column 2g / 2g+1 (any column permutation is
// legal as long as loads and stores agree), so lane (g, q) loads B[4kk + q][n0 + 2g .. 2g+1] -- 8 lanes x 16 B = 128
// contiguous bytes per row -- and ends up holding C[8it + g][n0 + 4q .. 4q+3], 32 contiguous bytes.
// grid.x = CTAs striding over the 16-column steps, grid.y = batch.
// ------------------------------------------------------------------------------------------------
template <int MT, int KT>
__global__ void __launch_bounds__(NTHREADS) smallmk_gemm_kernel(GemmArgs p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int b = blockIdx.y;
  const double* A = p.A + (long)b * p.sA;
  const double* B = p.B + (long)b * p.sB;
  double* C = p.C + (long)b * p.sC;
  double a[MT][KT];
#pragma unroll
  for (int it = 0; it < MT; ++it)
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const int r = 8 * it + g, k = 4 * kk + q;
      a[it][kk] = (r < p.M && k < p.K) ? __ldg(A + (long)r * p.lda + k) : 0.0;
    }
  const long nsteps = ((long)p.N + 15) / 16;
  const long nwarps = (long)gridDim.x * (NTHREADS / 32), w0 = (long)blockIdx.x * (NTHREADS / 32) + warp;
  constexpr int U = (MT * KT > 18) ? 2 : 4;       // 16-column steps in flight per warp (register budget: 255)
  for (long s0 = w0 * U; s0 < nsteps; s0 += nwarps * U) {
    double2 bv[U][KT];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long n = (s0 + u) * 16 + 2 * g;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        const int k = 4 * kk + q;
        double2 t = make_double2(0.0, 0.0);
        if (s0 + u < nsteps && k < p.K) {
          if (n + 1 < p.N) t = __ldg(reinterpret_cast<const double2*>(B + (long)k * p.ldb + n));
          else if (n < p.N) t.x = __ldg(B + (long)k * p.ldb + n);
        }
        bv[u][kk] = t;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (s0 + u >= nsteps) break;
      double acc[MT][2][2];
#pragma unroll
      for (int it = 0; it < MT; ++it) acc[it][0][0] = acc[it][0][1] = acc[it][1][0] = acc[it][1][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int it = 0; it < MT; ++it) {
          dmma884(acc[it][0][0], acc[it][0][1], a[it][kk], bv[u][kk].x);     // memory columns n0 + 2g   -> MMA column g
          dmma884(acc[it][1][0], acc[it][1][1], a[it][kk], bv[u][kk].y);     // memory columns n0 + 2g+1
        }
      // lane (g, q): tile 0 holds MMA columns 2q, 2q+1 = memory 4q, 4q+2; tile 1 holds memory 4q+1, 4q+3
      const long n = (s0 + u) * 16 + 4 * q;
#pragma unroll
      for (int it = 0; it < MT; ++it) {
        const int r = 8 * it + g;
        if (r >= p.M) continue;
        double* dst = C + (long)r * p.ldc + n;
        if (n + 3 < p.N) {
          *reinterpret_cast<double2*>(dst) = make_double2(acc[it][0][0], acc[it][1][0]);
          *reinterpret_cast<double2*>(dst + 2) = make_double2(acc[it][0][1], acc[it][1][1]);
        } else {
          if (n < p.N) dst[0] = acc[it][0][0];
          if (n + 1 < p.N) dst[1] = acc[it][1][0];
          if (n + 2 < p.N) dst[2] = acc[it][0][1];
        }
      }
    }
  }
}

template <int MT>
static int launch_smallmk_kt(GemmArgs& p, int batch, cudaStream_t st) {
  const long nsteps = ((long)p.N + 15) / 16;
  long ctas = (nsteps + 4 * (NTHREADS / 32) - 1) / (4 * (NTHREADS / 32));
  const long cap = 4L * gp_num_sms();
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  dim3 grid((unsigned)ctas, (unsigned)batch);
  const int kt = (p.K + 3) / 4;
  switch (kt) {
    case 1: smallmk_gemm_kernel<MT, 1><<<grid, NTHREADS, 0, st>>>(p); break;
    case 2: smallmk_gemm_kernel<MT, 2><<<grid, NTHREADS, 0, st>>>(p); break;
    case 3: smallmk_gemm_kernel<MT, 3><<<grid, NTHREADS, 0, st>>>(p); break;
    case 4: smallmk_gemm_kernel<MT, 4><<<grid, NTHREADS, 0, st>>>(p); break;
    case 5: smallmk_gemm_kernel<MT, 5><<<grid, NTHREADS, 0, st>>>(p); break;
    case 6: smallmk_gemm_kernel<MT, 6><<<grid, NTHREADS, 0, st>>>(p); break;
    case 7: smallmk_gemm_kernel<MT, 7><<<grid, NTHREADS, 0, st>>>(p); break;
    default: smallmk_gemm_kernel<MT, 8><<<grid, NTHREADS, 0, st>>>(p); break;
  }
  GP_CUDA(cudaGetLastError());
  return 0;
}

static int launch_smallmk(GemmArgs& p, int batch, cudaStream_t st) {
  if (batch > 65535) return gp_fail("gemm batch too large");
  switch ((p.M + 7) / 8) {
    case 1: return launch_smallmk_kt<1>(p, batch, st);
    case 2: return launch_smallmk_kt<2>(p, batch, st);
    case 3: return launch_smallmk_kt<3>(p, batch, st);
    default: return launch_smallmk_kt<4>(p, batch, st);
  }
}

// fixed-order reduction of [n][2] partials -> out[2]
__global__ void reduce_pairs_kernel(const double* __restrict__ part, long n, double* __restrict__ out) {
  __shared__ double red[16];
  double a = 0.0, b = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    a += part[2 * i];
    b += part[2 * i + 1];
  }
  const double s0 = block_sum(a, red);
  const double s1 = block_sum(b, red + 8);
  if (threadIdx.x == 0) {
    out[0] = s0;
    out[1] = s1;
  }
}

// grouped variant: out[g * out_stride + {0,1}] = sum of partial pairs [g * n, (g+1) * n)   (one CTA per group / restart)
__global__ void reduce_pairs_grouped_kernel(const double* __restrict__ part, long n, double* __restrict__ out, long out_stride) {
  __shared__ double red[16];
  const double* pg = part + 2 * n * blockIdx.x;
  double a = 0.0, b = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    a += pg[2 * i];
    b += pg[2 * i + 1];
  }
  const double s0 = block_sum(a, red);
  const double s1 = block_sum(b, red + 8);
  if (threadIdx.x == 0) {
    out[(long)blockIdx.x * out_stride] = s0;
    out[(long)blockIdx.x * out_stride + 1] = s1;
  }
}

template <int BM, int BN, int WM, int WN, int STAGES, bool BT, int EPI, int MINB>
static int launch_gemm(GemmArgs& p, int batch, cudaStream_t st) {
  using L = SmemLayout<BM, BN, BT, STAGES>;
  auto kern = dmma_gemm_kernel<BM, BN, WM, WN, STAGES, BT, EPI, MINB>;
  static int attr_set[GP_MAX_DEVICES];
  if (gp_first_use_on_device(attr_set)) {
    GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
  }
  p.m_tiles = (p.M + BM - 1) / BM;
  const long n_tiles = ((long)p.N + BN - 1) / BN;
  const long gx = (long)p.m_tiles * n_tiles;
  if (gx > 2147483647L || batch > 65535) return gp_fail("gemm grid too large");
  dim3 grid((unsigned)gx, (unsigned)batch);
  kern<<<grid, NTHREADS, L::BYTES, st>>>(p);
  GP_CUDA(cudaGetLastError());
  return 0;
}

static int check_gemm_alignment(const GemmArgs& p) {
  if ((p.lda | p.ldb | p.ldc | p.sA | p.sB | p.sC) & 1L) return gp_fail("gemm: leading dimensions / batch strides must be even");
  if (((uintptr_t)p.A | (uintptr_t)p.B | (uintptr_t)p.C) & 15) return gp_fail("gemm: operand pointers must be 16-byte aligned");
  return 0;
}

template <int EPI>
static int dispatch_gemm(GemmArgs& p, int transB, int batch, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || batch <= 0) return 0;
  if (int e = check_gemm_alignment(p)) return e;
  if (EPI == EPI_STORE && !transB && p.M <= 32 && p.K <= 32 && p.K > 0) return launch_smallmk(p, batch, st);
  if (p.M <= 32) {
    if (transB) return launch_gemm<32, 128, 32, 16, 3, true, EPI, 2>(p, batch, st);
    return launch_gemm<32, 128, 32, 16, 3, false, EPI, 2>(p, batch, st);
  }
  if (batch > 65535) return gp_fail("gemm batch too large");
  if (EPI == EPI_STORE) {
    // Latency mode: a product whose 128 x 64 tiles would occupy less than half of the SMs (the 250-order rotations
    // Q C Q^T of the gradient, the eigenvector products of gpcsd_eigh_dc: 16 tiles, ~22 us each on 16 SMs) is cut into
    // 32 x 32 tiles instead -- 16 times the CTAs, each with a sixteenth of the k-loop work.
    const long big = (long)((p.M + 127) / 128) * ((p.N + 63) / 64) * batch;
    if (2 * big <= gp_num_sms()) {
      if (transB) return launch_gemm<32, 32, 8, 16, 4, true, EPI, 2>(p, batch, st);
      return launch_gemm<32, 32, 8, 16, 4, false, EPI, 2>(p, batch, st);
    }
  }
  return tma_gemm(transB, p.M, p.N, p.K, p.A, p.lda, p.sA, p.B, p.ldb, p.sB, p.C, p.ldc, p.sC, batch, p.rD, p.ldrd,
                  p.partials, EPI == EPI_QUAD, st, p.a_div, p.grp);
}

// ------------------------------------------------------------------------------------------------
// segment-weighted split-K SYRK
// ------------------------------------------------------------------------------------------------
struct SyrkArgs {
  const double* X;
  long row_stride, seg_stride;
  const double* w;
  int M, nseg, seglen;
  int kbps;        // k-blocks per segment
  long total_kb;   // nseg * kbps
  int nsplit;
  int tiles_1d;    // tiles per side
  double* ws;      // [nsplit][ntiles][BM*BN]
  long sX = 0, sW = 0, sWs = 0;   // restart strides of X, w and ws (small-M kernel: blockIdx.y = restart)
};

// ---- M <= 32: register-only variant ---------------------------------------------------------------
// One diagonal tile only, so the B fragment of MMA column-tile j IS the A fragment of row-tile j: lane (g,q)
// loads X[8i+g][k + 2q .. 2q+1] with one LDG.128 per row-tile and feeds two DMMAs (k-slots {0,2,4,6} and
// {1,3,5,7} -- any k permutation is legal as long as both operands use it).  Nothing is staged in shared
// memory; the kernel is HBM-bound (6 flop/byte) and keeps 4 chunks x MT LDG.128 in flight per lane.
// PAIR: additionally accumulate the UNWEIGHTED product in the same pass over X (Ms = sum_j lt_j B_j B_j^T and Ns = sum_j B_j B_j^T
// of the per-electrode-noise gradient read Bm once instead of twice); its partials follow the weighted ones in ws.
template <int MT, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 2) wsyrk_small_kernel(SyrkArgs p) {
  extern __shared__ __align__(16) double red[];  // [8 warps][32*32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  p.X += (long)blockIdx.y * p.sX;
  if (p.w) p.w += (long)blockIdx.y * p.sW;
  p.ws += (long)blockIdx.y * p.sWs;
  const long cps = (p.seglen + 7) / 8;
  const long total = (long)p.nseg * cps;
  const long nwarps = (long)gridDim.x * (NTHREADS / 32), wg = (long)blockIdx.x * (NTHREADS / 32) + warp;
  long c = total * wg / nwarps;
  const long c1 = total * (wg + 1) / nwarps;
  long seg = c / cps, kc = c - seg * cps;

  double acc[MT][MT][2];
  double accp[PAIR ? MT : 1][PAIR ? MT : 1][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      acc[i][j][0] = acc[i][j][1] = 0.0;
      if (PAIR) accp[PAIR ? i : 0][PAIR ? j : 0][0] = accp[PAIR ? i : 0][PAIR ? j : 0][1] = 0.0;
    }

  const double* rowp[MT];
  bool rowok[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    rowok[i] = (8 * i + g) < p.M;
    rowp[i] = p.X + (long)(rowok[i] ? 8 * i + g : 0) * p.row_stride + 2 * q;
  }
  constexpr int U = 4;
  while (c < c1) {
    double2 v[U][MT];
    double wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool live = (c + u) < c1;
      const long k = kc * 8 + 2 * q;
      const long off = seg * p.seg_stride + kc * 8;
      wv[u] = (live && p.w) ? __ldg(p.w + seg) : 1.0;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        double2 t = make_double2(0.0, 0.0);
        if (live && rowok[i]) {
          if (k + 1 < p.seglen)
            t = __ldg(reinterpret_cast<const double2*>(rowp[i] + off));
          else if (k < p.seglen)
            t.x = __ldg(rowp[i] + off);
        }
        v[u][i] = t;
      }
      if (++kc == cps) {
        kc = 0;
        ++seg;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const double a0 = v[u][i].x * wv[u], a1 = v[u][i].y * wv[u];
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          dmma884(acc[i][j][0], acc[i][j][1], a0, v[u][j].x);
          dmma884(acc[i][j][0], acc[i][j][1], a1, v[u][j].y);
          if (PAIR) {
            dmma884(accp[PAIR ? i : 0][PAIR ? j : 0][0], accp[PAIR ? i : 0][PAIR ? j : 0][1], v[u][i].x, v[u][j].x);
            dmma884(accp[PAIR ? i : 0][PAIR ? j : 0][0], accp[PAIR ? i : 0][PAIR ? j : 0][1], v[u][i].y, v[u][j].y);
          }
        }
      }
    }
    c += U;
  }
  // cross-warp reduction (fixed order) -> one 32x32 partial per CTA; only lower 8x8 tiles are meaningful
  double* mine = red + warp * 1024;
#pragma unroll
  for (int set = 0; set < (PAIR ? 2 : 1); ++set) {
    if (set) __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double2 o = make_double2(0.0, 0.0);
        if (i < MT && j <= i && j < MT) {
          const int ii = i < MT ? i : 0, jj = j < MT ? j : 0;
          o = (set == 0) ? make_double2(acc[ii][jj][0], acc[ii][jj][1])
                         : make_double2(accp[PAIR ? ii : 0][PAIR ? jj : 0][0], accp[PAIR ? ii : 0][PAIR ? jj : 0][1]);
        }
        *reinterpret_cast<double2*>(mine + (8 * i + g) * 32 + 8 * j + 2 * q) = o;
      }
    __syncthreads();
    double* out = p.ws + ((long)set * gridDim.x + blockIdx.x) * 1024;
    for (int e = tid; e < 1024; e += NTHREADS) {
      double sacc = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < NTHREADS / 32; ++w8) sacc += red[w8 * 1024 + e];
      out[e] = sacc;
    }
  }
}

// One WARP per output entry (128 CTAs instead of 4): lane l sums partials l, l + 32, ... and a shuffle tree closes the sum --
// a fixed order, so the result is deterministic.  grid.y selects the weighted / unweighted set of the PAIR variant.
__global__ void wsyrk_small_reduce_kernel(const double* __restrict__ ws, int nparts, int M, double* __restrict__ C0,
                                          double* __restrict__ C1, long ldc, long ws_stride, long c_stride) {
  ws += (long)blockIdx.z * ws_stride;
  C0 += (long)blockIdx.z * c_stride;
  if (C1) C1 += (long)blockIdx.z * c_stride;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= 1024) return;
  const int r = e >> 5, cc = e & 31;
  if (r >= M || cc >= M) return;
  const int es = ((r >> 3) >= (cc >> 3)) ? e : (cc * 32 + r);  // upper tiles: mirror of the computed lower tile
  const double* src = ws + (long)blockIdx.y * nparts * 1024;
  double s0 = 0.0;
  for (int pidx = lane; pidx < nparts; pidx += 32) s0 += src[(long)pidx * 1024 + es];
  s0 = warp_sum(s0);
  if (lane == 0) (blockIdx.y ? C1 : C0)[(long)r * ldc + cc] = s0;
}

static int syrk_small_ctas() { return 2 * gp_num_sms(); }

// tile edge of the TMA SYRK: 64 where the lower triangle in 64-tiles is at most 0.65 of the area 128-tiles would compute
// (M = 192: 6 x 64^2 against 3 x 128^2 = 0.50 -> 398 us becomes 208 us at the Neuropixels shape).  At M = 250 the ratio is
// 0.83, but 64-tiles read every row panel twice as often (16 x 64 rows against 4 x 128 rows per k-block) and measured 240 us
// against 156 us, so the larger tiles stay.  GPCSD_SYRK_TILE=128|64 overrides.
static int syrk_tile_edge(int M) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("GPCSD_SYRK_TILE");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 64 || forced == 128) return forced;
  const long t128 = (M + 127) / 128, t64 = (M + 63) / 64;
  const long a128 = t128 * (t128 + 1) / 2 * 128 * 128, a64 = t64 * (t64 + 1) / 2 * 64 * 64;
  return (a64 * 100 <= a128 * 65) ? 64 : 128;
}

static int syrk_plan(int M, int nseg, int seglen, int& bmn, int& tiles_1d, int& kbps, long& total_kb, int& nsplit) {
  bmn = (M <= 32) ? 32 : syrk_tile_edge(M);
  tiles_1d = (M + bmn - 1) / bmn;
  const long ntiles = (long)tiles_1d * (tiles_1d + 1) / 2;
  kbps = (seglen + BK - 1) / BK;
  total_kb = (long)nseg * kbps;
  // whole waves: ntiles * nsplit <= 2 CTAs-worth per SM (1 resident CTA/SM -> 2 full waves, no ragged tail); with very few
  // tiles (the half-order blocks of the folded bases: 3 tiles) ONE wave divides just as evenly and halves the partial tiles
  // the reduction kernel has to read back (38 MB -> 19 MB at M = 250)
  long ns = ((ntiles <= 4 ? 1L : 2L) * gp_num_sms()) / ntiles;
  const long min_kb = 8;  // at least 8 k-blocks per split
  if (ns > total_kb / min_kb) ns = total_kb / min_kb;
  if (ns < 1) ns = 1;
  if (ns > 65535) ns = 65535;
  nsplit = (int)ns;
  return 0;
}

// launch the small-M SYRK (+ its reduction); pair != nullptr selects the one-pass weighted + unweighted variant
static int syrk_small_ctas_batched(int R) {
  if (R <= 1) return syrk_small_ctas();
  const int c = syrk_small_ctas() / R;
  return c < 4 ? 4 : c;
}

static int wsyrk_small_launch(SyrkArgs& p, const SyrkArgs* pair, double* C0, double* C1, long ldc, cudaStream_t st, int R = 1,
                              long c_stride = 0) {
  const int mt = (p.M + 7) / 8, nct = syrk_small_ctas_batched(R);
  p.sWs = (long)(pair ? 2 : 1) * nct * 1024;
  const size_t bytes = (size_t)(NTHREADS / 32) * 1024 * sizeof(double);
  static int attr[GP_MAX_DEVICES];
  if (gp_first_use_on_device(attr)) {
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    GP_CUDA(cudaFuncSetAttribute(wsyrk_small_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  if (pair) {
    switch (mt) {
      case 1: wsyrk_small_kernel<1, true><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
      case 2: wsyrk_small_kernel<2, true><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
      case 3: wsyrk_small_kernel<3, true><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
      default: wsyrk_small_kernel<4, true><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
    }
  } else {
    switch (mt) {
      case 1: wsyrk_small_kernel<1, false><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
      case 2: wsyrk_small_kernel<2, false><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
      case 3: wsyrk_small_kernel<3, false><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
      default: wsyrk_small_kernel<4, false><<<dim3(nct, R), NTHREADS, bytes, st>>>(p); break;
    }
  }
  GP_CUDA(cudaGetLastError());
  wsyrk_small_reduce_kernel<<<dim3(1024 * 32 / 256, pair ? 2 : 1, R), 256, 0, st>>>(p.ws, nct, p.M, C0, C1, ldc, p.sWs, c_stride);
  GP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gpcsd

using namespace gpcsd;

extern "C" {

int gpcsd_dgemm(int transB, int M, int N, int K, const double* A, long lda, long strideA, const double* B, long ldb,
                long strideB, double* C, long ldc, long strideC, int batch, void* stream) {
  GemmArgs p{};
  p.A = A; p.B = B; p.C = C;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = strideA; p.sB = strideB; p.sC = strideC;
  p.M = M; p.N = N; p.K = K;
  return dispatch_gemm<EPI_STORE>(p, transB, batch, (cudaStream_t)stream);
}

static long project_quad_ctas(int nx, int nt, int ntrials) {
  if (nt > 32) return tma_gemm_ctas(nt, ntrials, nx);   // persistent kernel: one partial pair per CTA
  const long ntl = ((long)ntrials + 127) / 128;
  return ntl * nx;
}

long gpcsd_project_quad_ws_doubles(int nx, int nt, int ntrials) { return 2 * project_quad_ctas(nx, nt, ntrials) + 2; }

int gpcsd_project_quad(int nx, int nt, int ntrials, const double* QtT, long ldq, const double* Z, long ldn,
                       const double* rD, long ldrd, double* Bout, double* partials, double* out2, void* stream) {
  GemmArgs p{};
  p.A = QtT; p.B = Z; p.C = Bout;
  p.lda = ldq; p.ldb = ldn; p.ldc = ldn;
  p.sA = 0; p.sB = (long)nt * ldn; p.sC = (long)nt * ldn;
  p.M = nt; p.N = ntrials; p.K = nt;
  p.rD = rD; p.ldrd = ldrd; p.partials = partials;
  if (int e = dispatch_gemm<EPI_QUAD>(p, 0, nx, (cudaStream_t)stream)) return e;
  reduce_pairs_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, project_quad_ctas(nx, nt, ntrials), out2);
  GP_CUDA(cudaGetLastError());
  return 0;
}

// Same contraction for a sub-block of the temporal eigenbasis (centrosymmetric fold): order m, operands are views into the
// parent [nx][nt][ldn] arrays whose per-electrode stride is bstride.
int gpcsd_project_quad_strided(int nx, int m, int ntrials, const double* AT, long lda, const double* Z, long ldn, long bstride,
                               const double* rD, long ldrd, double* Bout, double* partials, double* out2, void* stream) {
  GemmArgs p{};
  p.A = AT; p.B = Z; p.C = Bout;
  p.lda = lda; p.ldb = ldn; p.ldc = ldn;
  p.sA = 0; p.sB = bstride; p.sC = bstride;
  p.M = m; p.N = ntrials; p.K = m;
  p.rD = rD; p.ldrd = ldrd; p.partials = partials;
  if (int e = dispatch_gemm<EPI_QUAD>(p, 0, nx, (cudaStream_t)stream)) return e;
  reduce_pairs_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, project_quad_ctas(nx, m, ntrials), out2);
  GP_CUDA(cudaGetLastError());
  return 0;
}

/* Restart-batched form of gpcsd_project_quad_strided: R restarts x nx spatial eigen-indices in ONE launch.  Restart r uses the
 * m x m block AT + r*strideA and the parent arrays Z / Bout / rD of restart r, which must follow each other contiguously
 * (Z_r = Z + r*nx*bstride, rD_r = rD + r*nx*ldrd); out2[r*out_stride + {0,1}] = (quad, bsq) of restart r.
 * partials: gpcsd_project_quad_batched_ws_doubles(R, nx, m, ntrials) doubles. */
long gpcsd_project_quad_batched_ws_doubles(int R, int nx, int m, int ntrials) {
  if (m > 32) return 2L * R * tma_gemm_ctas(m, ntrials, nx * R) + 2;
  return 2L * R * nx * (((long)ntrials + 127) / 128) + 2;
}

int gpcsd_project_quad_batched(int R, int nx, int m, int ntrials, const double* AT, long lda, long strideA, const double* Z, long ldn,
                               long bstride, const double* rD, long ldrd, double* Bout, double* partials, double* out2,
                               long out_stride, void* stream) {
  if (R <= 0 || nx <= 0) return 0;
  GemmArgs p{};
  p.A = AT; p.B = Z; p.C = Bout;
  p.lda = lda; p.ldb = ldn; p.ldc = ldn;
  p.sA = strideA; p.sB = bstride; p.sC = bstride;
  p.M = m; p.N = ntrials; p.K = m;
  p.a_div = nx; p.grp = nx;
  p.rD = rD; p.ldrd = ldrd; p.partials = partials;
  if ((long)R * nx > 65535) return gp_fail("project_quad_batched: R * nx too large");
  if (int e = dispatch_gemm<EPI_QUAD>(p, 0, R * nx, (cudaStream_t)stream)) return e;
  const long per = (m > 32) ? tma_gemm_ctas(m, ntrials, nx * R) : (long)nx * (((long)ntrials + 127) / 128);
  reduce_pairs_grouped_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(partials, per, out2, out_stride);
  GP_CUDA(cudaGetLastError());
  return 0;
}

long gpcsd_wsyrk_ws_doubles(int M, int nseg, int seglen) {
  int bmn, t1, kbps, nsplit;
  long total;
  syrk_plan(M, nseg, seglen, bmn, t1, kbps, total, nsplit);
  if (bmn == 32) return (long)syrk_small_ctas() * 1024;
  return (long)nsplit * ((long)t1 * (t1 + 1) / 2) * bmn * bmn;
}

int gpcsd_wsyrk(int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, const double* w,
                double* C, long ldc, double* ws, void* stream) {
  if (M <= 0) return 0;
  if ((row_stride | seg_stride) & 1L) return gp_fail("wsyrk: strides must be even");
  if (((uintptr_t)X | (uintptr_t)ws) & 15) return gp_fail("wsyrk: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  SyrkArgs p{};
  int bmn;
  syrk_plan(M, nseg, seglen, bmn, p.tiles_1d, p.kbps, p.total_kb, p.nsplit);
  p.X = X; p.row_stride = row_stride; p.seg_stride = seg_stride; p.w = w;
  p.M = M; p.nseg = nseg; p.seglen = seglen; p.ws = ws;
  const long ntiles = (long)p.tiles_1d * (p.tiles_1d + 1) / 2;
  dim3 grid((unsigned)ntiles, (unsigned)p.nsplit);
  if (bmn == 32) return wsyrk_small_launch(p, nullptr, C, nullptr, ldc, st);
  return tma_wsyrk(M, nseg, seglen, X, row_stride, seg_stride, w, C, ldc, ws, p.nsplit, p.tiles_1d, p.kbps, p.total_kb, st, 1, 0, 0,
                   0, bmn);
}

/* Restart-batched SYRK: for r < R,  Cw_r = sum_seg w_r[seg] X_r,seg X_r,seg^T  (and Cp_r, the unweighted product, when Cp != NULL)
 * with X_r = X + r*strideX, w_r = w + r*strideW, C_r = C + r*strideC -- all restarts of a multi-start batch in one launch
 * (blockIdx.z / the tensor map's 4th dimension).  ws: gpcsd_wsyrk_batched_ws_doubles(R, M, nseg, seglen, Cp != NULL). */
static void syrk_plan_batched(int R, int M, int nseg, int seglen, int& bmn, int& t1, int& kbps, long& total, int& nsplit) {
  syrk_plan(M, nseg, seglen, bmn, t1, kbps, total, nsplit);
  if (R > 1) {
    nsplit = nsplit / R;
    if (nsplit < 1) nsplit = 1;
  }
}

long gpcsd_wsyrk_batched_ws_doubles(int R, int M, int nseg, int seglen, int pair) {
  int bmn, t1, kbps, nsplit;
  long total;
  syrk_plan_batched(R, M, nseg, seglen, bmn, t1, kbps, total, nsplit);
  if (bmn == 32) return (long)R * (pair ? 2 : 1) * syrk_small_ctas_batched(R) * 1024;
  return (long)R * nsplit * ((long)t1 * (t1 + 1) / 2) * bmn * bmn;      // (the pair runs two passes through the same scratch)
}

int gpcsd_wsyrk_batched(int R, int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, long strideX,
                        const double* w, long strideW, double* Cw, double* Cp, long ldc, long strideC, double* ws, void* stream) {
  if (M <= 0 || R <= 0) return 0;
  if ((row_stride | seg_stride | strideX) & 1L) return gp_fail("wsyrk: strides must be even");
  if (((uintptr_t)X | (uintptr_t)ws) & 15) return gp_fail("wsyrk: pointers must be 16-byte aligned");
  if (R > 65535) return gp_fail("wsyrk_batched: too many restarts");
  cudaStream_t st = (cudaStream_t)stream;
  SyrkArgs p{};
  int bmn;
  syrk_plan_batched(R, M, nseg, seglen, bmn, p.tiles_1d, p.kbps, p.total_kb, p.nsplit);
  p.X = X; p.row_stride = row_stride; p.seg_stride = seg_stride; p.w = w;
  p.M = M; p.nseg = nseg; p.seglen = seglen; p.ws = ws;
  p.sX = strideX; p.sW = strideW;
  if (bmn == 32) return wsyrk_small_launch(p, Cp ? &p : nullptr, Cw, Cp, ldc, st, R, strideC);
  if (int e = tma_wsyrk(M, nseg, seglen, X, row_stride, seg_stride, w, Cw, ldc, ws, p.nsplit, p.tiles_1d, p.kbps, p.total_kb, st, R,
                        strideX, strideW, strideC, bmn))
    return e;
  if (Cp)
    return tma_wsyrk(M, nseg, seglen, X, row_stride, seg_stride, nullptr, Cp, ldc, ws, p.nsplit, p.tiles_1d, p.kbps, p.total_kb, st, R,
                     strideX, 0, strideC, bmn);
  return 0;
}

/* Cw = sum_seg w[seg] X_seg X_seg^T and Cp = sum_seg X_seg X_seg^T in ONE pass over X (M <= 32: the Ms / Ns pair of the
 * per-electrode-noise gradient); larger M: two passes.  ws: 2 * gpcsd_wsyrk_ws_doubles(M, nseg, seglen). */
int gpcsd_wsyrk_pair(int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, const double* w,
                     double* Cw, double* Cp, long ldc, double* ws, void* stream) {
  if (M <= 0) return 0;
  if (M > 32) {
    if (int e = gpcsd_wsyrk(M, nseg, seglen, X, row_stride, seg_stride, w, Cw, ldc, ws, stream)) return e;
    return gpcsd_wsyrk(M, nseg, seglen, X, row_stride, seg_stride, nullptr, Cp, ldc, ws, stream);
  }
  if ((row_stride | seg_stride) & 1L) return gp_fail("wsyrk: strides must be even");
  if (((uintptr_t)X | (uintptr_t)ws) & 15) return gp_fail("wsyrk: pointers must be 16-byte aligned");
  SyrkArgs p{};
  int bmn;
  syrk_plan(M, nseg, seglen, bmn, p.tiles_1d, p.kbps, p.total_kb, p.nsplit);
  p.X = X; p.row_stride = row_stride; p.seg_stride = seg_stride; p.w = w;
  p.M = M; p.nseg = nseg; p.seglen = seglen; p.ws = ws;
  return wsyrk_small_launch(p, &p, Cw, Cp, ldc, (cudaStream_t)stream);
}

}  // extern "C"
