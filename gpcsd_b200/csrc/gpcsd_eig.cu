// Symmetric eigensolver for orders 130..256 built around cuSOLVER's one-CTA-per-matrix batched path (order <= 128).
//
// cuSOLVER syevd spends ~1.2 ms + 8 us per column in a latency-bound tridiagonalisation (profiles/r01_eigh_options.md);
// for the factor orders that dominate GPCSD evaluations (the 250-order halves of a 500-point temporal factor, the
// 192-order halves of a Neuropixels spatial factor) this file replaces it by one level of Cuppen's divide and conquer:
//
//   1. tridiag_cluster_kernel   Householder tridiagonalisation M = H T H^T.  One 8-CTA thread-block cluster per matrix:
//                               the matrix lives in the cluster's shared memory (row-cyclic slabs, <= 64 KB per CTA), the
//                               reflector and the symv result are exchanged through distributed shared memory, two
//                               cluster barriers per column (~1 us per column instead of ~8).
//   2. split                    T = blockdiag(T1', T2') + rho u u^T, rho = |e_mid|; T1', T2' (order n/2 <= 128) go through
//                               cusolverDnXsyevBatched (gpcsd_eigh_batched).
//   3. dc_merge_kernel          eigen-decomposition of D + rho z z^T: LAPACK-style deflation (tiny z_i, close d_i), secular
//                               roots in pole-shifted coordinates (bracketed geometric/arithmetic bisection, one thread per
//                               root), Gu-Eisenstat recomputation of z for numerically orthogonal eigenvectors.
//   4. one GEMM (gpcsd_dgemm)   eigenvectors of T = S^T * (rotated block eigenvectors)
//   5. backtransform_kernel     apply the Householder reflectors to every eigenvector (one thread per vector).
//
// All matrices row-major; eigenvectors are stored as ROWS (Q^T), the convention of gpcsd_eigh.
#include <cooperative_groups.h>

#include "common.h"
#include "dmma_gemm.cuh"

namespace cg = cooperative_groups;

namespace gpcsd {

constexpr int TRD_CLUSTER = 8;        // CTAs per matrix
constexpr int TRD_THREADS = 256;
constexpr int TRD_MAXN = 256;
constexpr int TRD_ROWS = TRD_MAXN / TRD_CLUSTER;   // local rows per CTA (row i lives in CTA i % 8, slot i / 8)
constexpr int TRD_LD = TRD_MAXN + 2;               // smem row stride (even, +2 keeps rows 16-byte aligned and de-phased)

struct TridiagSmem {
  double A[TRD_ROWS][TRD_LD];   // local row slab (full rows, both triangles kept current)
  double v[2][TRD_MAXN];        // reflector, double buffered by column parity (written by the owner into every CTA)
  double p[2][TRD_MAXN];        // tau * A22 * v, all-gathered (every CTA writes its rows into every CTA)
  double part[2][TRD_CLUSTER];  // partial p.v per CTA
  double hdr[2][4];             // tau, beta broadcast by the owner
  double red[TRD_THREADS / 32];
};

__device__ __forceinline__ double cta_sum(double x, double* red) {
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = x;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < TRD_THREADS / 32; ++i) s += red[i];
  return s;
}

// grid = 8 * nmat CTAs, cluster (8,1,1).  M: [nmat][n][ldm] symmetric.  Out: d[nmat][n], e[nmat][n] (e[k] = T[k+1][k]),
// V[nmat][n][ldv] (row k holds reflector k: V[k][j] for j > k+1, implicit 1 at j = k+1), tau[nmat][n].
__global__ void __cluster_dims__(TRD_CLUSTER, 1, 1) __launch_bounds__(TRD_THREADS, 1)
    tridiag_cluster_kernel(int n, const double* __restrict__ M, long ldm, double* __restrict__ d, double* __restrict__ e,
                           double* __restrict__ V, long ldv, double* __restrict__ tau) {
  extern __shared__ __align__(16) unsigned char trd_raw[];
  TridiagSmem& S = *reinterpret_cast<TridiagSmem*>(trd_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int mat = blockIdx.x / TRD_CLUSTER;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  M += (long)mat * n * ldm;
  d += (long)mat * n;
  e += (long)mat * n;
  tau += (long)mat * n;
  V += (long)mat * n * ldv;

  // load the local rows
  const int nloc = (n - rank + TRD_CLUSTER - 1) / TRD_CLUSTER;   // rows rank, rank+8, ...
  for (int idx = tid; idx < nloc * n; idx += TRD_THREADS) {
    const int s = idx / n, j = idx % n;
    S.A[s][j] = M[(long)(rank + s * TRD_CLUSTER) * ldm + j];
  }
  TridiagSmem* peers[TRD_CLUSTER];
#pragma unroll
  for (int r = 0; r < TRD_CLUSTER; ++r) peers[r] = cluster.map_shared_rank(&S, r);
  cluster.sync();

  for (int k = 0; k < n - 2; ++k) {
    const int buf = k & 1;
    const int len = n - k - 1;          // trailing order; reflector acts on indices k+1 .. n-1
    // ---- (a) the owner of row k builds the reflector from its row (== column k by symmetry) and broadcasts it
    if (rank == k % TRD_CLUSTER) {
      const double* row = S.A[k / TRD_CLUSTER];
      double ss = 0.0;
      for (int j = k + 2 + tid; j < n; j += TRD_THREADS) ss += row[j] * row[j];
      const double xnorm2 = cta_sum(ss, S.red);
      const double alpha = row[k + 1];
      double t = 0.0, beta = alpha, scal = 0.0;
      if (xnorm2 > 0.0) {
        beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
        t = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
      }
      for (int j = k + 1 + tid; j < n; j += TRD_THREADS) {
        const double vj = (j == k + 1) ? 1.0 : row[j] * scal;
#pragma unroll
        for (int r = 0; r < TRD_CLUSTER; ++r) peers[r]->v[buf][j] = vj;
        if (j > k + 1) V[(long)k * ldv + j] = vj;
      }
      if (tid == 0) {
#pragma unroll
        for (int r = 0; r < TRD_CLUSTER; ++r) {
          peers[r]->hdr[buf][0] = t;
          peers[r]->hdr[buf][1] = beta;
        }
        d[k] = row[k];
        e[k] = beta;
        tau[k] = t;
      }
    }
    cluster.sync();
    const double t = S.hdr[buf][0];
    // ---- (b) p_i = tau * sum_j A_ij v_j for the local rows i > k, all-gathered; partial p.v
    double pv = 0.0;
    for (int s = warp; s < nloc; s += TRD_THREADS / 32) {
      const int i = rank + s * TRD_CLUSTER;
      if (i <= k) continue;
      double acc = 0.0;
      for (int j = k + 1 + lane; j < n; j += 32) acc += S.A[s][j] * S.v[buf][j];
      acc = warp_sum(acc) * t;
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < TRD_CLUSTER; ++r) peers[r]->p[buf][i] = acc;
        pv += acc * S.v[buf][i];
      }
    }
    const double pvsum = cta_sum(pv, S.red);
    if (tid == 0) {
#pragma unroll
      for (int r = 0; r < TRD_CLUSTER; ++r) peers[r]->part[buf][rank] = pvsum;
    }
    cluster.sync();
    // ---- (c) w = p - (tau/2)(p.v) v ;  A22 -= v w^T + w v^T on the local rows
    double dot = 0.0;
#pragma unroll
    for (int r = 0; r < TRD_CLUSTER; ++r) dot += S.part[buf][r];
    const double c = 0.5 * t * dot;
    if (t != 0.0) {
      for (int s = warp; s < nloc; s += TRD_THREADS / 32) {
        const int i = rank + s * TRD_CLUSTER;
        if (i <= k) continue;
        const double vi = S.v[buf][i], wi = S.p[buf][i] - c * vi;
        for (int j = k + 1 + lane; j < n; j += 32) {
          const double vj = S.v[buf][j], wj = S.p[buf][j] - c * vj;
          S.A[s][j] -= vi * wj + wi * vj;
        }
      }
    }
    __syncthreads();   // the next owner reads its own (just updated) row; v/p buffers alternate, so no cluster barrier here
    (void)len;
  }
  // last 2x2 block
  cluster.sync();
  if (rank == (n - 2) % TRD_CLUSTER && tid == 0) {
    const double* row = S.A[(n - 2) / TRD_CLUSTER];
    d[n - 2] = row[n - 2];
    e[n - 2] = row[n - 1];
    tau[n - 2] = 0.0;
  }
  if (rank == (n - 1) % TRD_CLUSTER && tid == 0) {
    d[n - 1] = S.A[(n - 1) / TRD_CLUSTER][n - 1];
    e[n - 1] = 0.0;
    tau[n - 1] = 0.0;
  }
  cluster.sync();   // keep every CTA's shared memory alive until all remote accesses are done
}

// Apply H = H_0 H_1 ... H_{n-3} to every eigenvector: row x of XT (eigenvector of T) -> H x (H_k symmetric), reflectors
// applied from k = n-3 down to 0.  One WARP per vector: lane l keeps x[l + 32 m] (m < 8) in registers, the reflectors are
// read from global memory (L1/L2 hits: every warp streams the same V).
constexpr int BT_WARPS = 8;
__global__ void __launch_bounds__(32 * BT_WARPS) backtransform_kernel(int n, const double* __restrict__ V, long ldv,
                                                                      const double* __restrict__ tau,
                                                                      double* __restrict__ XT, long ldx) {
  const int mat = blockIdx.y;
  V += (long)mat * n * ldv;
  tau += (long)mat * n;
  XT += (long)mat * n * ldx;
  const int lane = threadIdx.x & 31;
  const int vec = blockIdx.x * BT_WARPS + (threadIdx.x >> 5);
  if (vec >= n) return;
  double x[TRD_MAXN / 32];
#pragma unroll
  for (int m = 0; m < TRD_MAXN / 32; ++m) {
    const int j = lane + 32 * m;
    x[m] = (j < n) ? XT[(long)vec * ldx + j] : 0.0;
  }
  for (int k = n - 3; k >= 0; --k) {
    const double t = __ldg(tau + k);
    if (t == 0.0) continue;
    const double* vk = V + (long)k * ldv;
    double vv[TRD_MAXN / 32];
    double dot = 0.0;
#pragma unroll
    for (int m = 0; m < TRD_MAXN / 32; ++m) {
      const int j = lane + 32 * m;
      vv[m] = (j > k + 1 && j < n) ? __ldg(vk + j) : ((j == k + 1) ? 1.0 : 0.0);
      dot += vv[m] * x[m];
    }
    dot = warp_sum(dot) * t;
#pragma unroll
    for (int m = 0; m < TRD_MAXN / 32; ++m) x[m] -= dot * vv[m];
  }
#pragma unroll
  for (int m = 0; m < TRD_MAXN / 32; ++m) {
    const int j = lane + 32 * m;
    if (j < n) XT[(long)vec * ldx + j] = x[m];
  }
}

}  // namespace gpcsd

using namespace gpcsd;

extern "C" {

// Householder tridiagonalisation of `nmat` symmetric matrices of order n (3 <= n <= 256) on 8-CTA clusters.
int gpcsd_tridiag(int n, int nmat, const double* M, long ldm, double* d, double* e, double* V, long ldv, double* tau,
                  void* stream) {
  if (n < 3 || n > TRD_MAXN) return gp_fail("gpcsd_tridiag: order must be in 3..256");
  static bool attr = false;
  if (!attr) {
    GP_CUDA(cudaFuncSetAttribute(tridiag_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TridiagSmem)));
    attr = true;
  }
  tridiag_cluster_kernel<<<TRD_CLUSTER * nmat, TRD_THREADS, sizeof(TridiagSmem), (cudaStream_t)stream>>>(n, M, ldm, d, e, V, ldv, tau);
  GP_CUDA(cudaGetLastError());
  return 0;
}

// XT (rows = eigenvectors of the tridiagonal matrices) -> rows = eigenvectors of the original matrices.
int gpcsd_backtransform(int n, int nmat, const double* V, long ldv, const double* tau, double* XT, long ldx, void* stream) {
  if (n < 3 || n > TRD_MAXN) return gp_fail("gpcsd_backtransform: order must be in 3..256");
  dim3 grid((n + BT_WARPS - 1) / BT_WARPS, nmat);
  backtransform_kernel<<<grid, 32 * BT_WARPS, 0, (cudaStream_t)stream>>>(n, V, ldv, tau, XT, ldx);
  GP_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
