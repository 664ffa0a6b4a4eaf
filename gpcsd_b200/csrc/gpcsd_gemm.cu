// DMMA GEMM kernels + launchers: generic strided-batched GEMM, fused projection+quadratic-form,
// segment-weighted split-K SYRK.  See include/gpcsd_b200.h for the ABI contract.
#include "common.h"
#include "dmma_gemm.cuh"

namespace gpcsd {

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  long lda, ldb, ldc;
  long sA, sB, sC;  // batch strides
  int M, N, K;
  int m_tiles;
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

// C_b = A_b * op(B_b).  grid.x = m_tiles * n_tiles (m fastest so CTAs sharing a B column panel are
// co-resident and the panel is fetched from HBM once), grid.y = batch.
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
  const double* A = p.A + (long)b * p.sA + (long)m0 * p.lda;
  const double* B = p.B + (long)b * p.sB + (BT ? (long)n0 * p.ldb : (long)n0);
  const int rowsA = p.M - m0;
  const long colsB = (long)p.N - n0;
  const int nkb = (p.K + BK - 1) / BK;

  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int st, int kb) {
    const long k0 = (long)kb * BK;
    load_kmajor_tile<BM>(sA + st * L::A_STAGE, A + k0, p.lda, rowsA, p.K - k0, p.A);
    if (BT)
      load_kmajor_tile<BN>(sB + st * L::B_STAGE, B + k0, p.ldb, (int)(colsB > BN ? BN : colsB), p.K - k0, p.B);
    else
      load_nmajor_tile<BN>(sB + st * L::B_STAGE, B + k0 * p.ldb, p.ldb, p.K - k0, colsB, p.B);
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

template <int BM, int BN, int WM, int WN, int STAGES, bool BT, int EPI, int MINB>
static int launch_gemm(GemmArgs& p, int batch, cudaStream_t st) {
  using L = SmemLayout<BM, BN, BT, STAGES>;
  auto kern = dmma_gemm_kernel<BM, BN, WM, WN, STAGES, BT, EPI, MINB>;
  static bool attr_set = false;
  if (!attr_set) {
    GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
    attr_set = true;
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
  if (p.M <= 32) {
    if (transB) return launch_gemm<32, 128, 32, 16, 3, true, EPI, 2>(p, batch, st);
    return launch_gemm<32, 128, 32, 16, 3, false, EPI, 2>(p, batch, st);
  }
  if (transB) return launch_gemm<128, 128, 64, 32, 4, true, EPI, 1>(p, batch, st);
  return launch_gemm<128, 128, 64, 32, 4, false, EPI, 1>(p, batch, st);
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
};

template <int BMN, int WM, int WN, int STAGES, int MINB>
__global__ void __launch_bounds__(NTHREADS, MINB) wsyrk_kernel(SyrkArgs p) {
  constexpr int STAGE = BMN * KMAJ_LD;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  constexpr int WARPS_M = BMN / WM;
  static_assert((BMN / WM) * (BMN / WN) == NTHREADS / 32, "warp tiling must use 8 warps");
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;

  // lower-triangular tile index -> (tm, tn), tn <= tm
  int t = blockIdx.x, tm = 0;
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  const int tn = t - tm * (tm + 1) / 2;
  const bool diag = (tm == tn);
  const int split = blockIdx.y;
  const long f0 = p.total_kb * split / p.nsplit, f1 = p.total_kb * (split + 1) / p.nsplit;
  const int nkb = (int)(f1 - f0);

  const double* XA = p.X + (long)tm * BMN * p.row_stride;
  const double* XB = p.X + (long)tn * BMN * p.row_stride;
  const int rowsA = p.M - tm * BMN, rowsB = p.M - tn * BMN;

  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int st, long f) {
    const int seg = (int)(f / p.kbps);
    const long k0 = (long)(f - (long)seg * p.kbps) * BK;
    const long off = (long)seg * p.seg_stride + k0;
    load_kmajor_tile<BMN>(sA + st * STAGE, XA + off, p.row_stride, rowsA, p.seglen - k0, p.X);
    if (!diag) load_kmajor_tile<BMN>(sB + st * STAGE, XB + off, p.row_stride, rowsB, p.seglen - k0, p.X);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, f0 + s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kb + STAGES - 1;
    if (nxt < nkb) load_stage(nxt % STAGES, f0 + nxt);
    cp_async_commit();
    const int st = kb % STAGES;
    const double wgt = p.w ? __ldg(p.w + (f0 + kb) / p.kbps) : 1.0;
    const double* a = sA + st * STAGE + (wm * WM) * KMAJ_LD;
    const double* bb = (diag ? sA : sB) + st * STAGE + (wn * WN) * KMAJ_LD;
    if (p.w)
      mma_kblock<WM, WN, true, KMAJ_LD, true>(a, bb, acc, g, q, wgt);
    else
      mma_kblock<WM, WN, true, KMAJ_LD, false>(a, bb, acc, g, q, 1.0);
  }
  cp_async_wait<0>();

  const long ntiles = (long)p.tiles_1d * (p.tiles_1d + 1) / 2;
  double* out = p.ws + ((long)split * ntiles + t) * (BMN * BMN);
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) {
      const int r = wm * WM + i * 8 + g, c = wn * WN + j * 8 + 2 * q;
      *reinterpret_cast<double2*>(out + r * BMN + c) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

template <int BMN>
__global__ void wsyrk_reduce_kernel(const double* __restrict__ ws, int nsplit, int tiles_1d, int M,
                                    double* __restrict__ C, long ldc) {
  const long ntiles = (long)tiles_1d * (tiles_1d + 1) / 2;
  const int t = blockIdx.y;
  int tm = 0;
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  const int tn = t - tm * (tm + 1) / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= BMN * BMN) return;
  const int r = e / BMN, c = e % BMN;
  const int m = tm * BMN + r, n = tn * BMN + c;
  if (m >= M || n >= M) return;
  double s = 0.0;
  for (int sp = 0; sp < nsplit; ++sp) s += ws[((long)sp * ntiles + t) * (BMN * BMN) + e];
  if (tm == tn) {
    C[(long)m * ldc + n] = s;
  } else {
    C[(long)m * ldc + n] = s;
    C[(long)n * ldc + m] = s;
  }
}

static int syrk_plan(int M, int nseg, int seglen, int& bmn, int& tiles_1d, int& kbps, long& total_kb, int& nsplit) {
  bmn = (M <= 32) ? 32 : 128;
  tiles_1d = (M + bmn - 1) / bmn;
  const long ntiles = (long)tiles_1d * (tiles_1d + 1) / 2;
  kbps = (seglen + BK - 1) / BK;
  total_kb = (long)nseg * kbps;
  const int sms = gp_num_sms();
  const long ctas_target = (bmn == 32) ? 6L * sms : 2L * sms;
  long ns = (ctas_target + ntiles - 1) / ntiles;
  const long min_kb = 8;  // at least 8 k-blocks per split
  if (ns > total_kb / min_kb) ns = total_kb / min_kb;
  if (ns < 1) ns = 1;
  if (ns > 65535) ns = 65535;
  nsplit = (int)ns;
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
  const int bm = (nt <= 32) ? 32 : 128;
  const long mt = (nt + bm - 1) / bm, ntl = ((long)ntrials + 127) / 128;
  return mt * ntl * nx;
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

long gpcsd_wsyrk_ws_doubles(int M, int nseg, int seglen) {
  int bmn, t1, kbps, nsplit;
  long total;
  syrk_plan(M, nseg, seglen, bmn, t1, kbps, total, nsplit);
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
  if (bmn == 32) {
    constexpr int STAGES = 6;
    const size_t bytes = (size_t)2 * STAGES * 32 * KMAJ_LD * sizeof(double);
    auto kern = wsyrk_kernel<32, 8, 16, STAGES, 4>;
    static bool attr = false;
    if (!attr) { GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); attr = true; }
    kern<<<grid, NTHREADS, bytes, st>>>(p);
    GP_CUDA(cudaGetLastError());
    wsyrk_reduce_kernel<32><<<dim3(4, (unsigned)ntiles), 256, 0, st>>>(ws, p.nsplit, p.tiles_1d, M, C, ldc);
  } else {
    constexpr int STAGES = 4;
    const size_t bytes = (size_t)2 * STAGES * 128 * KMAJ_LD * sizeof(double);
    auto kern = wsyrk_kernel<128, 64, 32, STAGES, 1>;
    static bool attr = false;
    if (!attr) { GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); attr = true; }
    kern<<<grid, NTHREADS, bytes, st>>>(p);
    GP_CUDA(cudaGetLastError());
    wsyrk_reduce_kernel<128><<<dim3(64, (unsigned)ntiles), 256, 0, st>>>(ws, p.nsplit, p.tiles_1d, M, C, ldc);
  }
  GP_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
