// In-house symmetric eigensolver for orders 3..256 (FP64, full eigen-decomposition with vectors), one 8-CTA thread-block
// cluster per matrix, batched over matrices.
//
// cuSOLVER syevd spends ~1.2 ms + 8 us per column in a latency-bound tridiagonalisation (profiles/r01_eigh_options.md) and
// it is the Amdahl term of every GPCSD evaluation (the 250-order halves of a 500-point temporal factor, the 192-order halves
// of a Neuropixels spatial factor, comp_eig_D utility_functions.py:44-64).  Three kernels replace it:
//
//   1. tridiag_cluster_kernel   Householder tridiagonalisation M = H T H^T.  The matrix lives in the cluster's shared memory
//                               (row-cyclic slabs, <= 64 KB per CTA); the reflector and the symv result are exchanged through
//                               distributed shared memory, two cluster barriers per column.
//   2. dc_cluster_kernel        Cuppen divide and conquer on T down to 1x1 leaves (ceil(log2 n) merge levels): per merge
//                               LAPACK-style deflation, secular roots by bracketed two-pole rational iteration (one warp per
//                               root, roots spread over the 128 warps of the cluster), Gu-Eisenstat recomputation of z for
//                               numerically orthogonal eigenvectors, eigenvector update as a shared-memory tiled GEMM whose
//                               A operand (z_j / (d_j - lambda_i)) is generated on the fly.  Small per-merge vectors are
//                               broadcast through distributed shared memory, the eigenvector matrices ping-pong through L2.
//                               The numerical core (dc_core.h) also compiles for the host: tests/test_dc_host.py checks it
//                               against LAPACK without a GPU.
//   3. backtransform_kernel     apply the Householder reflectors to every eigenvector (one warp per vector).
//
// All matrices row-major; eigenvectors are stored as ROWS (Q^T), the convention of gpcsd_eigh.
#include <cooperative_groups.h>
#include <stddef.h>

#include "common.h"
#include "dc_core.h"
#include "dmma_gemm.cuh"

namespace cg = cooperative_groups;

namespace gpcsd {

// shared::cluster address of `ptr` (a shared-memory object of this CTA) in CTA `rank` of the cluster, and a remote store
__device__ __forceinline__ uint32_t dsmem_addr(const void* ptr, int rank) {
  uint32_t local = (uint32_t)__cvta_generic_to_shared(ptr), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local), "r"(rank));
  return remote;
}
__device__ __forceinline__ void dsmem_store(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}
// remote store that also signals 8 transaction bytes on an mbarrier of the SAME remote CTA: data and "it has arrived" travel
// together, so the consumer needs neither a cluster barrier nor a fence
__device__ __forceinline__ void dsmem_store_signal(uint32_t addr, double v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];\n" ::"r"(addr), "d"(v), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

constexpr int TRD_CLUSTER = 8;        // CTAs per matrix
constexpr int TRD_MAXN = 256;
constexpr int TRD_WARPS = 16;                      // warps per CTA
constexpr int TRD_RPW = TRD_MAXN / (TRD_CLUSTER * TRD_WARPS);   // rows per warp (2): row i lives in CTA i % 8, slot i / 8
constexpr int TRD_NR = TRD_MAXN / 32;              // row elements per lane

struct TridiagSmem {
  double v[2][TRD_MAXN];   // reflector, double buffered by column parity (written by the owner warp into every CTA)
  double p[2][TRD_MAXN];   // tau * A22 * v, all-gathered (every row warp writes its entries into every CTA)
  double hdr[2][2];        // tau broadcast by the owner
  uint64_t bar_v, bar_p[2];   // transaction barriers: "reflector k has arrived", "all p_i of column k have arrived" (by parity)
};

// Register-resident Householder tridiagonalisation.  grid = 8 * nmat CTAs, cluster (8,1,1), up to 16 warps per CTA:
// every matrix row lives in the REGISTERS of one warp (slot s = i / 8 belongs to warp s % 16; lane l holds columns l, l+32,
// ...; both triangles are kept current), so the symv p = tau A v and the rank-2 update A -= v w^T + w v^T touch shared
// memory only for the two exchanged vectors.  Per column: the warp that owns row k+1 builds the next reflector straight
// from its registers right after updating them and stores it into every CTA of the cluster (distributed shared memory);
// every row warp stores its p_i into every CTA.  Both exchanges are st.async stores that complete transaction bytes on an
// mbarrier in the receiving CTA, so the per-column synchronisation is two mbarrier waits (~DSMEM latency) -- no cluster
// barrier, no fence (a cluster barrier costs a MEMBAR.ALL.GPU that also waits for the reflector stores to global memory),
// no CTA-wide barrier, no shared-memory reduction.  The exchanged vectors are double buffered by column parity; the data
// dependencies of the algorithm (reflector k+1 needs every p of column k, p of column k+1 needs reflector k+1) guarantee
// that no CTA can run more than one exchange ahead of the slowest one (the p barrier is doubled by column parity so that
// a fast CTA's p of column k+1 can never be counted against a slow CTA's still-open column k).
// M: [nmat][n][ldm] symmetric.  Out: d[nmat][n], e[nmat][n] (e[k] = T[k+1][k]), V[nmat][n][ldv] (row k holds reflector k:
// V[k][j] for j > k+1, implicit 1 at j = k+1), tau[nmat][n].
__global__ void __cluster_dims__(TRD_CLUSTER, 1, 1) __launch_bounds__(32 * TRD_WARPS, 1)
    tridiag_cluster_kernel(int n, const double* __restrict__ M, long ldm, double* __restrict__ d, double* __restrict__ e,
                           double* __restrict__ V, long ldv, double* __restrict__ tau) {
  __shared__ __align__(16) TridiagSmem S;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int mat = blockIdx.x / TRD_CLUSTER;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  M += (long)mat * n * ldm;
  d += (long)mat * n;
  e += (long)mat * n;
  tau += (long)mat * n;
  V += (long)mat * n * ldv;

  int row[TRD_RPW];
  double a[TRD_RPW][TRD_NR];
#pragma unroll
  for (int q = 0; q < TRD_RPW; ++q) {
    row[q] = rank + (warp + q * TRD_WARPS) * TRD_CLUSTER;
#pragma unroll
    for (int m = 0; m < TRD_NR; ++m) {
      const int j = lane + 32 * m;
      a[q][m] = (row[q] < n && j < n) ? M[(long)row[q] * ldm + j] : 0.0;
    }
  }
  const int peer_rank = lane & (TRD_CLUSTER - 1);
  if (threadIdx.x == 0) {
    mbar_init(&S.bar_v, 1);
    mbar_init(&S.bar_p[0], 1);
    mbar_init(&S.bar_p[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  cluster.sync();                       // every CTA's shared memory and barriers are live before remote stores
  for (int k = -1; k < n - 2; ++k) {    // k = -1: only builds reflector 0
    const int buf = k & 1;
    const int kn = k + 1;
    if (threadIdx.x == 0) {             // arm this column's transaction counts (early arrivals are fine)
      if (k >= 0) mbar_expect_tx(&S.bar_p[buf], 8u * (uint32_t)(n - k - 1));
      if (kn < n - 2) mbar_expect_tx(&S.bar_v, 8u * (uint32_t)(n - kn));
    }
    if (k >= 0) {
      const double t = S.hdr[buf][0];
      // ---- p_i = tau * sum_{j > k} A_ij v_j, all-gathered
      double acc[TRD_RPW];
#pragma unroll
      for (int q = 0; q < TRD_RPW; ++q) acc[q] = 0.0;
#pragma unroll
      for (int m = 0; m < TRD_NR; ++m) {
        const int j = lane + 32 * m;
        const double vj = (j > k && j < n) ? S.v[buf][j] : 0.0;
#pragma unroll
        for (int q = 0; q < TRD_RPW; ++q) acc[q] += a[q][m] * vj;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < TRD_RPW; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
#pragma unroll
      for (int q = 0; q < TRD_RPW; ++q)
        if (row[q] > k && row[q] < n && lane < TRD_CLUSTER)
          dsmem_store_signal(dsmem_addr(&S.p[buf][row[q]], peer_rank), acc[q] * t, dsmem_addr(&S.bar_p[buf], peer_rank));
      mbar_wait_cluster(&S.bar_p[buf], (uint32_t)((k >> 1) & 1));
      // ---- w = p - (tau/2)(p.v) v ;  A22 -= v w^T + w v^T on the own rows
      double pv = 0.0;
#pragma unroll
      for (int m = 0; m < TRD_NR; ++m) {
        const int j = lane + 32 * m;
        if (j > k && j < n) pv += S.v[buf][j] * S.p[buf][j];
      }
      const double c = 0.5 * t * warp_sum(pv);
#pragma unroll
      for (int q = 0; q < TRD_RPW; ++q) {
        if (row[q] > k && row[q] < n) {
          const double vi = S.v[buf][row[q]], wi = S.p[buf][row[q]] - c * vi;
#pragma unroll
          for (int m = 0; m < TRD_NR; ++m) {
            const int j = lane + 32 * m;
            if (j > k && j < n) {
              const double vj = S.v[buf][j];
              a[q][m] -= vi * (S.p[buf][j] - c * vj) + wi * vj;
            }
          }
        }
      }
    }
    // ---- the owner of row k+1 builds reflector k+1 (== column k+1 by symmetry) straight from the registers it has just
    //      updated, and stores it into every CTA
    if (kn < n - 2) {
#pragma unroll
      for (int q = 0; q < TRD_RPW; ++q) {
        if (row[q] != kn) continue;                // warp-uniform
        const int nb = kn & 1, j1 = kn + 1;
        double sel = 0.0, dia = 0.0, ss = 0.0;
#pragma unroll
        for (int m = 0; m < TRD_NR; ++m) {
          const int j = lane + 32 * m;
          if (j == j1) sel = a[q][m];
          if (j == kn) dia = a[q][m];
          if (j > j1 && j < n) ss += a[q][m] * a[q][m];
        }
        const double alpha = __shfl_sync(0xffffffffu, sel, j1 & 31);
        const double akk = __shfl_sync(0xffffffffu, dia, kn & 31);
        const double xnorm2 = warp_sum(ss);
        double t = 0.0, beta = alpha, scal = 0.0;
        if (xnorm2 > 0.0) {
          beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
          t = (beta - alpha) / beta;
          scal = 1.0 / (alpha - beta);
        }
        if (lane < TRD_CLUSTER) dsmem_store_signal(dsmem_addr(&S.hdr[nb][0], peer_rank), t, dsmem_addr(&S.bar_v, peer_rank));
#pragma unroll
        for (int m = 0; m < TRD_NR; ++m) {
          const int j = lane + 32 * m;
          if (j >= j1 && j < n) {
            const double vj = (j == j1) ? 1.0 : a[q][m] * scal;
#pragma unroll
            for (int r = 0; r < TRD_CLUSTER; ++r) dsmem_store_signal(dsmem_addr(&S.v[nb][j], r), vj, dsmem_addr(&S.bar_v, r));
            if (j > j1) V[(long)kn * ldv + j] = vj;
          }
        }
        if (lane == 0) {
          d[kn] = akk;
          e[kn] = beta;
          tau[kn] = t;
        }
      }
      mbar_wait_cluster(&S.bar_v, (uint32_t)(kn & 1));
    }
  }
  // last 2x2 block
#pragma unroll
  for (int q = 0; q < TRD_RPW; ++q) {
    if (row[q] == n - 2 || row[q] == n - 1) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int m = 0; m < TRD_NR; ++m) {
        const int j = lane + 32 * m;
        if (j == row[q]) s0 = a[q][m];
        if (j == n - 1) s1 = a[q][m];
      }
      const double dd = __shfl_sync(0xffffffffu, s0, row[q] & 31), ee = __shfl_sync(0xffffffffu, s1, (n - 1) & 31);
      if (lane == 0) {
        d[row[q]] = dd;
        e[row[q]] = (row[q] == n - 2) ? ee : 0.0;
        tau[row[q]] = 0.0;
      }
    }
  }
  cluster.sync();   // keep every CTA's shared memory alive until all remote accesses are done
}

// ---------------------------------------------------------------------------------------------------------------------------
// Divide and conquer on the tridiagonal matrix
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int DC_CLUSTER = 8;
constexpr int DC_THREADS = 512;
constexpr int DC_WARPS = DC_THREADS / 32;
constexpr int DC_MAXN = 256;
constexpr int DC_MAXNODES = DC_MAXN / 2;
constexpr int DC_TM = 64, DC_TN = 64, DC_TK = 16;     // eigenvector-update tile (rows = new eigenvectors, cols = components)
constexpr int DC_DIRECT_MAX = 32;                     // merge nodes up to this size use the one-warp-per-row update

struct DcSmem {
  // replicated in every CTA (each CTA runs the O(n) bookkeeping redundantly, bit-identically)
  double d[DC_MAXN], dn[DC_MAXN], z[DC_MAXN], dl[DC_MAXN], w[DC_MAXN], rot_c[DC_MAXN], rot_s[DC_MAXN];
  int srt[DC_MAXN], row[DC_MAXN], rot_p[DC_MAXN], rot_n[DC_MAXN], pid[DC_MAXN];
  int na[DC_MAXNODES], nc[DC_MAXNODES], nb[DC_MAXNODES], nk[DC_MAXNODES], nrot[DC_MAXNODES];
  double nrho[DC_MAXNODES];
  unsigned long long ndmax[DC_MAXNODES], nzmax[DC_MAXNODES];
  // broadcast by the computing warp into every CTA (distributed shared memory)
  double mu[DC_MAXN], zh[DC_MAXN];
  int org[DC_MAXN];
  // eigenvector-update tiles
  double As[DC_TK][DC_TM + 2], Bs[DC_TK][DC_TN + 2];
  double nrm[DC_THREADS / DC_TM][DC_TM];
  double inv[DC_TM];
  double red[DC_WARPS];
};

// grid = 8 * nmat CTAs, cluster (8,1,1).  d_in[nmat][n], e_in[nmat][n] with e_in[k] = T[k+1][k].  Qa, Qb: [nmat][n][ldq]
// workspaces.  Out: W[nmat][n] ascending, XT[nmat][n][ldx] rows = eigenvectors of T in the order of W.
__global__ void __cluster_dims__(DC_CLUSTER, 1, 1) __launch_bounds__(DC_THREADS, 1)
    dc_cluster_kernel(int n, const double* __restrict__ d_in, const double* __restrict__ e_in, double* Qa, double* Qb, long ldq,
                      double* __restrict__ W, double* XT, long ldx) {
  extern __shared__ __align__(16) unsigned char dc_raw[];
  DcSmem& S = *reinterpret_cast<DcSmem*>(dc_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int mat = blockIdx.x / DC_CLUSTER;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gw = rank * DC_WARPS + warp;                 // warp index within the cluster
  constexpr int GW = DC_CLUSTER * DC_WARPS;
  d_in += (long)mat * n;
  e_in += (long)mat * n;
  Qa += (long)mat * n * ldq;
  Qb += (long)mat * n * ldq;
  W += (long)mat * n;
  XT += (long)mat * n * ldx;
  DcSmem* peer = cluster.map_shared_rank(&S, lane & (DC_CLUSTER - 1));

  // ---- scale to unit max-norm, tear every off-diagonal (leaves of size 1), Q = I
  double v = 0.0;
  for (int i = tid; i < n; i += DC_THREADS) v = fmax(v, fmax(fabs(d_in[i]), (i < n - 1) ? fabs(e_in[i]) : 0.0));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) S.red[warp] = v;
  __syncthreads();
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < DC_WARPS; ++i) scale = fmax(scale, S.red[i]);
  if (!(scale > 0.0)) scale = 1.0;
  if (tid < n) {
    const double el = (tid > 0) ? fabs(e_in[tid - 1] / scale) : 0.0, er = (tid < n - 1) ? fabs(e_in[tid] / scale) : 0.0;
    S.d[tid] = d_in[tid] / scale - el - er;
  }
  for (int idx = rank * DC_THREADS + tid; idx < n * n; idx += DC_CLUSTER * DC_THREADS) {
    const int r = idx / n, c = idx - r * n;
    Qa[(long)r * ldq + c] = (r == c) ? 1.0 : 0.0;
    Qb[(long)r * ldq + c] = 0.0;      // entries outside the diagonal blocks of a level are read as zeros by the next one
  }
  double* Qold = Qa;
  double* Qnew = Qb;
  cluster.sync();

  const int levels = dc::num_levels(n);
  for (int L = 1; L <= levels; ++L) {
    const int nodes = 1 << (levels - L);
    const int mmax = (n + nodes - 1) / nodes;
    // ---- node table
    if (tid < nodes) {
      const int a = dc::node_start(n, 2 * nodes, 2 * tid), c = dc::node_start(n, 2 * nodes, 2 * tid + 1),
                b = dc::node_start(n, 2 * nodes, 2 * tid + 2);
      S.na[tid] = a;
      S.nc[tid] = c;
      S.nb[tid] = b;
      S.nrho[tid] = (c == a || c == b) ? 0.0 : 2.0 * fabs(e_in[c - 1] / scale);
      S.ndmax[tid] = 0ull;
      S.nzmax[tid] = 0ull;
    }
    if (tid < n) S.pid[tid] = dc::node_of(n, nodes, tid);
    __syncthreads();
    // ---- z = [last component of the left child's eigenvectors; sign(e_c) * first component of the right child's] / sqrt 2
    if (tid < n) {
      const int p = S.pid[tid], c = S.nc[p];
      double zz = 0.0;
      if (S.nrho[p] != 0.0) {
        const double q = __ldcg(Qold + (long)tid * ldq + ((tid < c) ? c - 1 : c));
        zz = ((tid < c || e_in[c - 1] >= 0.0) ? q : -q) * 0.70710678118654752440;
      }
      S.z[tid] = zz;
      atomicMax(&S.ndmax[p], (unsigned long long)__double_as_longlong(fabs(S.d[tid])));
      atomicMax(&S.nzmax[p], (unsigned long long)__double_as_longlong(fabs(zz)));
    }
    __syncthreads();
    // ---- counting sort of the node's eigenvalues
    if (tid < n) {
      const int p = S.pid[tid], a = S.na[p], b = S.nb[p];
      const double dg = S.d[tid];
      int r = 0;
      for (int h = a; h < b; ++h) {
        const double dh = S.d[h];
        r += (dh < dg) || (dh == dg && h < tid);
      }
      S.srt[a + r] = tid;
    }
    __syncthreads();
    // ---- deflation, one thread per node
    if (tid < nodes) {
      const int a = S.na[tid], m = S.nb[tid] - a;
      int k = 0, nrot = 0;
      if (m > 0)
        dc::deflate(a, m, S.srt, S.d, S.z, S.nrho[tid], __longlong_as_double((long long)S.ndmax[tid]),
                    __longlong_as_double((long long)S.nzmax[tid]), S.row, S.dl, S.w, S.rot_p, S.rot_n, S.rot_c, S.rot_s, k,
                    nrot);
      S.nk[tid] = k;
      S.nrot[tid] = nrot;
    }
    __syncthreads();
    // ---- deflating Givens rotations on the rows of Qold: one thread per column, the running row stays in a register
    if (warp == 0) {
      const int col = rank * 32 + lane;
      if (col < n) {
        const int p = S.pid[col], a = S.na[p], nr = S.nrot[p];
        if (nr > 0) {
          double* Q = Qold + col;
          int crow = -1;
          double cval = 0.0;
          double vn_next = __ldcg(Q + (long)S.rot_n[a] * ldq);
          for (int q = 0; q < nr; ++q) {
            const int rp = S.rot_p[a + q], rn = S.rot_n[a + q];
            const double c = S.rot_c[a + q], s = S.rot_s[a + q];
            const double vn = vn_next;
            if (q + 1 < nr) vn_next = __ldcg(Q + (long)S.rot_n[a + q + 1] * ldq);
            double vp;
            if (rp == crow) {
              vp = cval;
            } else {
              vp = __ldcg(Q + (long)rp * ldq);
              if (crow >= 0) Q[(long)crow * ldq] = cval;
            }
            Q[(long)rp * ldq] = c * vp + s * vn;
            crow = rn;
            cval = c * vn - s * vp;
          }
          Q[(long)crow * ldq] = cval;
        }
      }
    }
    cluster.sync();
    // ---- secular roots: one warp per root, broadcast (mu, org) to every CTA
    for (int g = gw; g < n; g += GW) {
      const int p = S.pid[g], a = S.na[p], r = g - a, k = S.nk[p];
      if (r < k) {
        double mu;
        int org;
        dc::secular_root<dc::WarpLanes>(k, r, &S.dl[a], &S.w[a], S.nrho[p], mu, org);
        if (lane < DC_CLUSTER) {
          peer->mu[g] = mu;
          peer->org[g] = org;
        }
      }
    }
    cluster.sync();
    // ---- Gu-Eisenstat z
    for (int g = gw; g < n; g += GW) {
      const int p = S.pid[g], a = S.na[p], r = g - a, k = S.nk[p];
      if (r < k) {
        const double zh = dc::zhat_component<dc::WarpLanes>(k, r, &S.dl[a], &S.w[a], &S.mu[a], &S.org[a]);
        if (lane < DC_CLUSTER) peer->zh[g] = zh;
      }
    }
    cluster.sync();
    // ---- new eigenvalues (every CTA, identical)
    if (tid < n) {
      const int p = S.pid[tid], a = S.na[p], r = tid - a;
      S.dn[tid] = (r < S.nk[p]) ? S.dl[a + S.org[tid]] + S.mu[tid] : S.dl[tid];
    }
    // ---- eigenvector update: new row a+i = sum_j zh_j / (dl_j - lambda_i) / |.| * old row row[a+j]; deflated rows are copied
    if (mmax <= DC_DIRECT_MAX) {
      for (int g = gw; g < n; g += GW) {
        const int p = S.pid[g], a = S.na[p], m = S.nb[p] - a, r = g - a, k = S.nk[p];
        const bool active = lane < m;
        const int col = a + (active ? lane : 0);
        double out;
        if (r < k) {
          const int org_i = S.org[g];
          const double mu_i = S.mu[g];
          double acc = 0.0, nrm = 0.0;
          for (int j = 0; j < k; ++j) {
            const double cf = S.zh[a + j] / dc::delta_ji(&S.dl[a], j, org_i, mu_i);
            nrm += cf * cf;
            acc += cf * __ldcg(Qold + (long)S.row[a + j] * ldq + col);
          }
          out = acc / sqrt(nrm);
        } else {
          out = __ldcg(Qold + (long)S.row[g] * ldq + col);
        }
        if (active) Qnew[(long)g * ldq + col] = out;
      }
    } else {
      const int tr = (mmax + DC_TM - 1) / DC_TM, tc = (mmax + DC_TN - 1) / DC_TN;
      const int ntiles = nodes * tr * tc;
      const int ty = tid >> 4, tx = tid & 15;            // compute mapping: rows 2ty, 2ty+1; cols 4tx .. 4tx+3
      const int ii = tid & (DC_TM - 1), kq = tid >> 6;    // A-generation mapping: row ii, k-slots kq and kq + 8
      for (int T = rank; T < ntiles; T += DC_CLUSTER) {
        const int p = T / (tr * tc), rem = T - p * (tr * tc);
        const int i0 = (rem / tc) * DC_TM, c0 = (rem % tc) * DC_TN;
        const int a = S.na[p], m = S.nb[p] - a, k = S.nk[p];
        if (i0 >= m || c0 >= m) continue;                 // uniform over the CTA
        const bool vi = (i0 + ii) < k;
        const int org_i = vi ? S.org[a + i0 + ii] : 0;
        const double mu_i = vi ? S.mu[a + i0 + ii] : 1.0;
        double nrm_part = 0.0;
        double acc[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[h][q] = 0.0;
        const int kend = (i0 < k) ? k : 0;                // tiles made of deflated rows only: no GEMM
        for (int j0 = 0; j0 < kend; j0 += DC_TK) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int kk = kq + 8 * h, j = j0 + kk;
            double cf = 0.0;
            if (vi && j < k) cf = S.zh[a + j] / dc::delta_ji(&S.dl[a], j, org_i, mu_i);
            S.As[kk][ii] = cf;
            nrm_part += cf * cf;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int e = tid + DC_THREADS * h, kk = e >> 6, cc = e & 63, j = j0 + kk, col = c0 + cc;
            double bv = 0.0;
            if (j < k && col < m) bv = __ldcg(Qold + (long)S.row[a + j] * ldq + a + col);
            S.Bs[kk][cc] = bv;
          }
          __syncthreads();
#pragma unroll
          for (int kk = 0; kk < DC_TK; ++kk) {
            const double2 av = *reinterpret_cast<const double2*>(&S.As[kk][2 * ty]);
            const double2 b0 = *reinterpret_cast<const double2*>(&S.Bs[kk][4 * tx]);
            const double2 b1 = *reinterpret_cast<const double2*>(&S.Bs[kk][4 * tx + 2]);
            acc[0][0] += av.x * b0.x; acc[0][1] += av.x * b0.y; acc[0][2] += av.x * b1.x; acc[0][3] += av.x * b1.y;
            acc[1][0] += av.y * b0.x; acc[1][1] += av.y * b0.y; acc[1][2] += av.y * b1.x; acc[1][3] += av.y * b1.y;
          }
          __syncthreads();
        }
        S.nrm[kq][ii] = nrm_part;
        __syncthreads();
        if (tid < DC_TM) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < DC_THREADS / DC_TM; ++q) s += S.nrm[q][tid];
          S.inv[tid] = (s > 0.0) ? 1.0 / sqrt(s) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = i0 + 2 * ty + h;
          if (r >= m) continue;
          const double sc = S.inv[2 * ty + h];
          const long src = (long)S.row[a + r] * ldq + a;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = c0 + 4 * tx + q;
            if (col < m) Qnew[(long)(a + r) * ldq + a + col] = (r < k) ? acc[h][q] * sc : __ldcg(Qold + src + col);
          }
        }
        __syncthreads();
      }
    }
    cluster.sync();
    if (tid < n) S.d[tid] = S.dn[tid];
    double* t = Qold;
    Qold = Qnew;
    Qnew = t;
    __syncthreads();
  }

  // ---- ascending order, undo the scaling, permute the rows into XT
  if (tid < n) {
    const double dg = S.d[tid];
    int r = 0;
    for (int h = 0; h < n; ++h) {
      const double dh = S.d[h];
      r += (dh < dg) || (dh == dg && h < tid);
    }
    S.srt[tid] = r;
    if (rank == 0) W[r] = dg * scale;
  }
  __syncthreads();
  for (int g = gw; g < n; g += GW) {
    const long dst = (long)S.srt[g] * ldx;
    for (int col = lane; col < n; col += 32) XT[dst + col] = __ldcg(Qold + (long)g * ldq + col);
  }
  cluster.sync();   // no CTA may exit while peers can still address its shared memory
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
  const int slots = (n + TRD_CLUSTER - 1) / TRD_CLUSTER;               // rows per CTA
  const int threads = 32 * (slots < TRD_WARPS ? slots : TRD_WARPS);
  tridiag_cluster_kernel<<<TRD_CLUSTER * nmat, threads, 0, (cudaStream_t)stream>>>(n, M, ldm, d, e, V, ldv, tau);
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

// Eigen-decomposition of `nmat` symmetric tridiagonal matrices (d[nmat][n], e[nmat][n], e[k] = T[k+1][k]) by divide and
// conquer on 8-CTA clusters: W[nmat][n] ascending, XT[nmat][n][ldx] rows = eigenvectors.  ws: 2*nmat*n*ldx doubles.
long gpcsd_tridiag_eig_ws_doubles(int n, long ldx, int nmat) { return 2L * nmat * n * ldx; }

int gpcsd_tridiag_eig(int n, int nmat, const double* d, const double* e, double* W, double* XT, long ldx, double* ws,
                      long ws_doubles, void* stream) {
  if (n < 2 || n > DC_MAXN) return gp_fail("gpcsd_tridiag_eig: order must be in 2..256");
  if (ws_doubles < gpcsd_tridiag_eig_ws_doubles(n, ldx, nmat)) return gp_fail("gpcsd_tridiag_eig: workspace too small");
  static bool attr = false;
  if (!attr) {
    GP_CUDA(cudaFuncSetAttribute(dc_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DcSmem)));
    attr = true;
  }
  double* Qa = ws;
  double* Qb = ws + (long)nmat * n * ldx;
  dc_cluster_kernel<<<DC_CLUSTER * nmat, DC_THREADS, sizeof(DcSmem), (cudaStream_t)stream>>>(n, d, e, Qa, Qb, ldx, W, XT, ldx);
  GP_CUDA(cudaGetLastError());
  return 0;
}

// Full symmetric eigen-decomposition of `nmat` stacked matrices M[nmat][n][ldm] (3 <= n <= 256): QT[nmat][n][ldq] rows =
// eigenvectors, W[nmat][n] ascending.  Replaces np.linalg.eigh of utility_functions.py:58-59 for the factor orders that
// dominate GPCSD evaluations.  M is not modified.  ws: gpcsd_eigh_dc_ws_doubles(n, ldq, nmat) doubles.
long gpcsd_eigh_dc_ws_doubles(int n, long ldq, int nmat) { return (long)nmat * (3L * n * ldq + 3L * n); }

int gpcsd_eigh_dc(int n, int nmat, const double* M, long ldm, double* QT, long ldq, double* W, double* ws, long ws_doubles,
                  void* stream) {
  if (n < 3 || n > DC_MAXN) return gp_fail("gpcsd_eigh_dc: order must be in 3..256");
  if (ws_doubles < gpcsd_eigh_dc_ws_doubles(n, ldq, nmat)) return gp_fail("gpcsd_eigh_dc: workspace too small");
  double* V = ws;                                  // reflectors [nmat][n][ldq]
  double* Qab = V + (long)nmat * n * ldq;          // D&C ping-pong, 2 x [nmat][n][ldq]
  double* d = Qab + 2L * nmat * n * ldq;
  double* e = d + (long)nmat * n;
  double* tau = e + (long)nmat * n;
  if (gpcsd_tridiag(n, nmat, M, ldm, d, e, V, ldq, tau, stream)) return 1;
  if (gpcsd_tridiag_eig(n, nmat, d, e, W, QT, ldq, Qab, 2L * nmat * n * ldq, stream)) return 1;
  return gpcsd_backtransform(n, nmat, V, ldq, tau, QT, ldq, stream);
}

}  // extern "C"
