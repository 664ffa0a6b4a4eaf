// In-house symmetric eigensolvers for orders 1..256 (FP64, full eigen-decomposition with vectors), batched over matrices.
//
// cuSOLVER syevd spends ~1.2 ms + 8 us per column in a latency-bound tridiagonalisation (profiles/r01_eigh_options.md) and
// it is the Amdahl term of every GPCSD evaluation (the 250-order halves of a 500-point temporal factor, the 192-order halves
// of a Neuropixels spatial factor, comp_eig_D utility_functions.py:44-64).
//
// Orders <= 32: jacobi_small_kernel, one CTA per matrix -- parallel cyclic Jacobi in shared memory after a pre-rotation by
// the DCT-II or DCT-IV basis (whichever leaves the smaller off-diagonal norm: the two halves of a centrosymmetric Toeplitz
// factor are nearly diagonal in exactly these bases).  ~80-140 us whatever the batch (profiles/r02i_small_eigh.md).
//
// Orders 33..256: one 8-CTA thread-block cluster per matrix, three kernels (profiles/r02n_eigensolver.md):
//
//   1. tridiag_cluster_kernel<NR>  Householder tridiagonalisation M = H T H^T.  Every matrix row lives in the registers of
//                               one warp (row i in CTA i mod 8, up to 4 rows per warp); ONE exchange per column through
//                               distributed shared memory in which no warp is special: every row warp sends the raw pair
//                               ((A x)_i, a_{i,k+1}) with 16-byte st.async stores that complete transaction bytes on an
//                               mbarrier of the receiver -- no cluster barrier in the loop.  The reflector stays
//                               unnormalised, so its rsqrt / reciprocal chain runs next to the next column's reduction;
//                               the body is specialised on the register-block count and the first live block.
//                               ~1.08 us per column at n = 250, 0.8 us at n <= 64.
//   2. dc_cluster_kernel        Cuppen divide and conquer on T down to 1x1 leaves (ceil(log2 n) merge levels): per merge
//                               LAPACK-style deflation, secular roots by a bracketed two-pole rational iteration (1 to 32
//                               lanes per root depending on the merge size), Gu-Eisenstat recomputation of z for numerically
//                               orthogonal eigenvectors, eigenvector update as a shared-memory tiled DMMA GEMM whose A
//                               operand (z_j / (d_j - lambda_i)) is generated on the fly.  Bottom levels (at least one
//                               merge node per CTA) run CTA-locally; above, small per-merge vectors are broadcast through
//                               distributed shared memory and the eigenvector matrices ping-pong through L2.  The numerical
//                               core (dc_core.h) also compiles for the host: tests/test_dc_host.py checks it against LAPACK
//                               without a GPU.
//   3. back-transformation      n < 97: backtransform_kernel applies the reflectors to every eigenvector (one warp per
//                               vector).  n >= 97: the top-level merge and the back-transformation are two full-GPU DMMA
//                               GEMMs, QT = (C * Q_below) * H^T, with H^T formed by backtransform_kernel on a side stream
//                               while the divide-and-conquer kernel runs (gpcsd_eigh_dc below).
//
// All matrices row-major; eigenvectors are stored as ROWS (Q^T), the convention of gpcsd_eigh.
#include <cooperative_groups.h>
#include <stddef.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "common.h"
#include "dc_core.h"
#include "dmma_gemm.cuh"

namespace cg = cooperative_groups;

namespace gpcsd {

// shared::cluster address of `ptr` (a shared-memory object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* ptr, int rank) {
  uint32_t local = (uint32_t)__cvta_generic_to_shared(ptr), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local), "r"(rank));
  return remote;
}
// remote store that also signals 8 transaction bytes on an mbarrier of the SAME remote CTA: data and "it has arrived" travel
// together, so the consumer needs neither a cluster barrier nor a fence
__device__ __forceinline__ void dsmem_store_signal(uint32_t addr, double v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];\n" ::"r"(addr), "d"(v), "r"(mbar)
               : "memory");
}
// one bulk copy (the TMA engine; size a multiple of 16 B) from this CTA's shared memory into a peer's, completing `bytes`
// on an mbarrier of the peer -- a single instruction instead of bytes/8 st.async requests (measured: ~17 cycles of issue per
// 8-byte st.async request, 2200 cycles for a 16-entry row to 8 peers)
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
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
constexpr int TRD_WARPS = 8;                       // warps per CTA
constexpr int TRD_RPW = TRD_MAXN / (TRD_CLUSTER * TRD_WARPS);   // rows per warp (up to 4): row i lives in CTA i % 8, slot i / 8, warp slot % 8
constexpr int TRD_NR = TRD_MAXN / 32;              // row elements per lane

// -DGPCSD_EIG_PROF: clock64 phase timers of the cluster kernels (developer builds only; scripts/eig_prof_tridiag.py, scripts/eig_prof_dc.py, scripts/eig_trace_tridiag.py)
#ifdef GPCSD_EIG_PROF
__device__ long long g_eig_prof[64];
#define EIG_PROF_DECL long long prof_t[32] = {0}; long long prof_last = clock64();
__device__ long long g_eig_trace[8 * 16 * 12];
#define EIG_PROF(i) { const long long t_ = clock64(); prof_t[i] += t_ - prof_last; prof_last = t_; \
    if (rank == 0 && k >= 4 && k < 12 && lane == 0) g_eig_trace[((k - 4) * 16 + warp) * 12 + i] = t_; }
__device__ long long g_eig_lp[16 * 16];     // [level][phase] of the divide-and-conquer kernel
#define EIG_PROF2(i) { const long long t_ = clock64(); prof_t[i] += t_ - prof_last; \
    if (rank == 0 && tid == 0 && mat == 0 && prof_level >= 0) g_eig_lp[prof_level * 16 + (i)] = t_ - prof_last; prof_last = t_; }
#define EIG_PROF_DUMP(cond, cnt) if ((cond) && (threadIdx.x & 31) == 0) { for (int i_ = 0; i_ < (cnt); ++i_) g_eig_prof[i_] = prof_t[i_]; }
#else
#define EIG_PROF_DECL
#define EIG_PROF(i)
#define EIG_PROF2(i)
#define EIG_PROF_DUMP(cond, cnt)
#endif

struct TridiagSmem {
  double2 pc[2][TRD_MAXN];        // the per-column exchange, by parity: .x = (A x)_i (p_i = tau (.y + scal .x) is completed by
                                  // the receiver), .y = a_{i,k+1}, row i's element of the NEXT pivot column (every row warp
                                  // stores its pair into every CTA)
  double vs[TRD_WARPS][TRD_MAXN]; // per-warp private copy of the current (unnormalised) reflector (single-element reads without shuffles)
  double ael[TRD_WARPS][TRD_RPW]; // per warp: a_{i,k+1} of its rows, handed from the lane that holds column k+1 to the sending lanes
  uint64_t bar[2];                // transaction barriers of the exchange, by parity
};

// Warp totals of eight values with 9 shuffles instead of 40: every stage halves the number of values a lane carries (the
// upper half-warp keeps v[4..7] and hands v[0..3] to its partner, ...), so lane L ends up with the total of v[L >> 2].
// Partners add the same two numbers, so the four lanes of a group (and every warp given the same inputs) agree bit for bit.
__device__ __forceinline__ double warp_reduce8_transposed(const double (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  double u[4], t[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double keep = b4 ? v[i + 4] : v[i], give = b4 ? v[i] : v[i + 4];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double keep = b3 ? u[i + 2] : u[i], give = b3 ? u[i] : u[i + 2];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
  }
  const double keep = b2 ? t[1] : t[0], give = b2 ? t[0] : t[1];
  double r = keep + __shfl_xor_sync(0xffffffffu, give, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// 16-byte remote store that also signals 16 transaction bytes on an mbarrier of the same remote CTA
__device__ __forceinline__ void dsmem_store2_signal(uint32_t addr, double v0, double v1, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];\n" ::"r"(addr), "d"(v0),
               "d"(v1), "r"(mbar)
               : "memory");
}

// 1/sqrt(s) and 1/x to full double precision from the single-precision hardware approximations plus two Newton steps: the
// library sqrt and division are ~30 dependent FP64 instructions each (~1000 cycles for sqrt + div on this part, measured
// with clock64 inside the kernel) and they sit on the per-column critical path.  Arguments outside the float range take
// the library path.
__device__ __forceinline__ double fast_rsqrt(double s) {
  if (!(s > 1e-30 && s < 1e30)) return 1.0 / sqrt(s);
  const double y = (double)rsqrtf((float)s);          // ~2^-22
  const double r = fma(-s * y, y, 1.0);               // 1 - s y^2
  return fma(y * r, fma(r, 0.375, 0.5), y);           // y (1 + r/2 + 3 r^2/8): third order -> ~2^-62 (+ rounding)
}
// one third-order step from a single-precision seed y0 ~ 1/x
__device__ __forceinline__ double refine_recip(double x, double y0) {
  const double r = fma(-x, y0, 1.0);
  return fma(y0 * r, 1.0 + r, y0);                    // y (1 + r + r^2)
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// Register-resident, single-exchange Householder tridiagonalisation.  grid = 8 * nmat CTAs, cluster (8,1,1), up to 16 warps
// per CTA.  Every matrix row lives in the REGISTERS of one warp (slot s = i / 8 belongs to warp s % 16 of CTA i % 8; lane l
// holds columns l, l+32, ...; both triangles are kept current), so the symv p = tau A v and the rank-2 update
// A -= v w^T + w v^T never touch shared memory for the matrix.
//
// The classical algorithm needs two cluster-wide exchanges per column (broadcast the reflector, all-gather p).  Here there
// is ONE, and no warp is special in it: with its p_i of column k every row warp also sends a_{i,k+1}, its element of the
// next pivot column (= row k+1 by symmetry), so that after the exchange every warp holds p AND the next pivot row and
// redundantly (bit-identically) finishes column k (w = p - tau/2 (p.v) v), applies that rank-2 update to the received row
// to obtain column k+1 of the current matrix and builds reflector k+1 from it.  Per column:
//     wait -> { S1, S2 (warp reduction) || rsqrt/reciprocal of the previous reflector } -> x -> { |x|^2, A x, w.x, v.x }
//     (ONE merged warp reduction) -> send:
//  * the symv runs on the UNNORMALISED pivot column x (A v = a_{:,k+1} + scal * A x_tail), so it shares the reduction
//    with the norm instead of waiting for the reflector;
//  * it runs on the rows as they stand BEFORE the previous reflector's update (A' x = A x - v (w.x) - w (v.x), two more sums
//    in the same reduction); that rank-2 update of the own rows, the normalised reflector and its store to global memory
//    are done AFTER the send, under the latency of the exchange;
//  * what is sent is the RAW pair ((A' x)_i, a'_{i,k+1}) and the reflector is kept UNNORMALISED (x, not v = scal x): every
//    use of tau and scal = 1/(alpha - beta) becomes a uniform coefficient that is first needed at the END of the next
//    column's p.v reduction, so the ~20 dependent FP64 operations of the rsqrt / reciprocal chain are issued in the shadow
//    of that reduction instead of between the merged reduction and the send.
// With every warp of the SM executing this in lock step the loop is bound by instruction ISSUE, not by latency (clock64
// trace: every phase advances at the same pace in all warps; a first version with 16 warps x 2 rows and per-block
// predicates issued 916 instructions per warp and column), so
//  * the body is specialised at compile time on NR = ceil(n / 32) register blocks per row and on M0 = k / 32, the first
//    live block: no per-block predicates, no work on dead blocks; only blocks M0 and M0+1 carry lane masks.  Out-of-range
//    columns (j >= n) need no masks either: the rows and the exchange buffers are zero there and stay zero;
//  * 8 warps hold up to 4 rows each (the redundant per-warp part -- w, x, the reflector -- is issued 8 times per SM, not 16);
//  * the seven sums of the merged reduction go through ONE transposed butterfly (warp_reduce8_transposed);
//  * the per-row send arithmetic is done by the 8 lanes that send that row, not replicated for every row in every lane.
// Exchanges are st.async stores that complete transaction bytes on an mbarrier of the receiving CTA (double buffered by
// parity): no cluster barrier (it costs a MEMBAR.ALL.GPU that also waits for the reflector stores to global memory), no
// fence, no CTA-wide barrier, no shared-memory reduction.  A CTA cannot run more than one exchange ahead of the slowest one
// because every p_i of exchange k+1 needs all of exchange k.
// M: [nmat][n][ldm] symmetric.  Out: d[nmat][n], e[nmat][n] (e[k] = T[k+1][k]), V[nmat][n][ldv] (row k holds reflector k:
// V[k][j] for j > k+1, implicit 1 at j = k+1), tau[nmat][n].
template <int NR>
__global__ void __cluster_dims__(TRD_CLUSTER, 1, 1) __launch_bounds__(32 * TRD_WARPS, 1)
    tridiag_cluster_kernel(int n, const double* __restrict__ M, long ldm, double* __restrict__ d, double* __restrict__ e,
                           double* __restrict__ V, long ldv, double* __restrict__ tau) {
  constexpr int RPW = (NR + 1) / 2;     // row slots per warp that can be occupied at this order (n <= 32 NR)
  static_assert(RPW <= TRD_RPW && RPW + 3 <= 8, "row sums + three shared sums go through one 8-value reduction");
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

  int row[RPW];
  double a[RPW][NR];
#pragma unroll
  for (int q = 0; q < RPW; ++q) {
    row[q] = rank + (warp + q * TRD_WARPS) * TRD_CLUSTER;
#pragma unroll
    for (int m = 0; m < NR; ++m) {
      const int j = lane + 32 * m;
      a[q][m] = (row[q] < n && j < n) ? M[(long)row[q] * ldm + j] : 0.0;
    }
  }
  double* vs = S.vs[warp];
#pragma unroll
  for (int m = 0; m < TRD_NR; ++m) vs[lane + 32 * m] = 0.0;
  for (int i = threadIdx.x; i < 2 * TRD_MAXN; i += blockDim.x) (&S.pc[0][0])[i] = make_double2(0.0, 0.0);
  if (threadIdx.x == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  cluster.sync();                       // every CTA's shared memory and barriers are live before remote stores

  // lanes 8q .. 8q+7 send the pair of row slot q to CTAs 0-7
  const int peer_rank = lane & (TRD_CLUSTER - 1), myq = lane >> 3;
  const int send_row = rank + (warp + myq * TRD_WARPS) * TRD_CLUSTER;       // < 256
  const bool sender = myq < RPW && send_row < n;
  // (the peer's TridiagSmem sits at the same offset of its shared-memory window: one mapa, constant offsets)
  const uint32_t remote = dsmem_addr(&S, peer_rank) + 16u * (uint32_t)send_row;
  const uint32_t remote_bar = dsmem_addr(&S.bar[0], peer_rank);
  constexpr uint32_t PC1 = (uint32_t)sizeof(double2) * TRD_MAXN;

  // A warp takes part in column k while it owns a row >= k; afterwards it leaves the loop (it would neither send nor be
  // waited for).  The warp that owns the CTA's LAST row stays longest and arms the CTA's barriers.
  int last_row = -1;
#pragma unroll
  for (int q = 0; q < RPW; ++q)
    if (row[q] < n) last_row = row[q];
  const bool armer = lane == 0 && last_row >= 0 && last_row + TRD_CLUSTER >= n;

  // exchange 0: column 0 of every row (p = 0)
  if (armer) mbar_expect_tx(&S.bar[0], 16u * (uint32_t)n);
  {
    double e0 = 0.0;
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      const double eq = __shfl_sync(0xffffffffu, a[q][0], 0);
      if (myq == q) e0 = eq;
    }
    if (sender) dsmem_store2_signal(remote, 0.0, e0, remote_bar);
  }

  double xp[NR];                        // UNNORMALISED reflector k-1: the pivot column beyond the subdiagonal (zero for j <= k)
#pragma unroll
  for (int m = 0; m < NR; ++m) xp[m] = 0.0;
  double alpha_p = 0.0, xn2_p = 0.0;    // its alpha and |x|^2: tau, beta, 1/(alpha - beta) are formed one exchange later
  bool live = true;
  int k = 0;

  EIG_PROF_DECL
  // one column, specialised on the first live register block M0 = k / 32; returns false when this warp is done
  auto column = [&](auto M0c) -> bool {
    constexpr int M0 = decltype(M0c)::value;
    constexpr bool HAS1 = M0 + 1 < NR;  // block M0+1 exists
    constexpr int M1 = HAS1 ? M0 + 1 : M0;
    constexpr int Q0 = M0 / 2;          // row slots below Q0 (rows < 64 Q0 <= k) are finished in every warp
    EIG_PROF(0)
    const int pb = k & 1;               // exchange k: raw pairs of column k-1 / row k
    const bool build = k < n - 2;       // k = n-2: only finishes column n-3
    // (a warp whose last row is exactly k has nothing left to send either -- nobody waits for it, so it must not wait on
    //  barriers that may run ahead of it; the owner of row n-2 stays for the final update, after which nothing runs ahead)
    if (last_row < k || (last_row == k && build)) return false;
    if (armer && build) mbar_expect_tx(&S.bar[pb ^ 1], 16u * (uint32_t)(n - k - 1));     // exchange k+1: rows i > k
    const int kl = k & 31;
    const bool wrap = HAS1 && kl == 31; // column k+1 is lane 0 of block M0+1
    const int l1 = (kl + 1) & 31;       // lane of column k+1
    // a_{i,k+1} of the own rows (as the registers hold them: before the update by reflector k-1) -> the sending lanes
    if (lane == l1) {
#pragma unroll
      for (int q = Q0; q < RPW; ++q) S.ael[warp][q] = wrap ? a[q][M1] : a[q][M0];
    }
    mbar_wait_cluster(&S.bar[pb], (uint32_t)((k >> 1) & 1));
    EIG_PROF(1)
    // ---- the received raw pairs (s_j, a_j): p_j = tau (a_j + scal s_j) with the scalars of reflector k-1, v_k = 1 and
    //      v_j = scal x_j beyond, so  p.v = tau (a_k + scal (s_k + S1 + scal S2)),  S1 = sum a_j x_j, S2 = sum s_j x_j:
    //      the two sums do not need the scalars, whose rsqrt / reciprocal chain (~20 dependent operations) is issued next to
    //      the reduction instead of in front of it.  (x_j = 0 for j <= k: stale entries of finished rows drop out.)
    //      Every load from the exchange buffer happens BEFORE this warp's send: once all warps have sent, the peers may
    //      run on and overwrite this parity with exchange k+2.
    double sp[NR], x[NR];
    double S1 = 0.0, S2 = 0.0, S1b = 0.0, S2b = 0.0;
#pragma unroll
    for (int m = M0; m < NR; ++m) {
      const double2 t2 = S.pc[pb][lane + 32 * m];
      sp[m] = t2.x;
      x[m] = t2.y;
      if ((m - M0) & 1) {
        S1b += t2.y * xp[m];
        S2b += t2.x * xp[m];
      } else {
        S1 += t2.y * xp[m];
        S2 += t2.x * xp[m];
      }
    }
    S1 += S1b;
    S2 += S2b;
    const double2 ek = S.pc[pb][k], ek1 = S.pc[pb][k + 1];   // (s_k, a_k), (s_{k+1}, a_{k+1}): uniform loads
    const double xk1 = vs[k + 1];                            // x^{k-1}_{k+1}
    double vi[RPW], wi[RPW];
    double2 pr[RPW];
    bool act[RPW];
#pragma unroll
    for (int q = Q0; q < RPW; ++q) {
      act[q] = row[q] > k && row[q] < n;                     // warp-uniform; row k itself is final
      const int r = act[q] ? row[q] : k;
      vi[q] = vs[r];                                         // x_i for now
      pr[q] = S.pc[pb][r];
    }
    double vim = vs[send_row];                               // the same for the row this lane sends
    const double2 prm = S.pc[pb][send_row];
    __syncwarp();
    const double aem = S.ael[warp][myq & (TRD_RPW - 1)];
    // reflector k-1: tau = 1 + |alpha| / |beta|, beta = -sign(alpha) sqrt(alpha^2 + |x|^2), scal = 1 / (alpha - beta);
    // fast path without branches (MUFU seeds + Newton, see fast_rsqrt / refine_recip), discarded when out of range
    const double s2 = fma(alpha_p, alpha_p, xn2_p), aa = fabs(alpha_p);
    const bool nz = xn2_p > 0.0, fastp = s2 > 1e-30 && s2 < 1e30;
    double tt, abv, rden;
    {
      // (single MUFU instructions without the denormal / range fix-up branches of rsqrtf, sqrtf and __frcp_rn: the chain
      //  stays in ONE basic block with the reduction below, so the scheduler interleaves the two)
      const float sf = (float)s2;
      float yf, y0f;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(sf));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(fmaf(sf, yf, (float)aa)));      // ~ 1 / (|alpha| + sqrt(s2))
      const double y = (double)yf;
      const double r = fma(-s2 * y, y, 1.0);
      const double rn = fma(y * r, fma(r, 0.375, 0.5), y);   // 1 / |beta|
      abv = s2 * rn;                                         // |beta|
      tt = fma(aa, rn, 1.0);
      const double den = aa + abv;                           // |alpha - beta|
      const double y0 = (double)y0f;
      const double r2 = fma(-den, y0, 1.0);
      rden = fma(y0 * r2, 1.0 + r2, y0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      S1 += __shfl_xor_sync(0xffffffffu, S1, o);
      S2 += __shfl_xor_sync(0xffffffffu, S2, o);
    }
    if (!fastp && nz) {                 // rare: outside the single-precision range
      const double rn = 1.0 / sqrt(s2);
      abv = s2 * rn;
      tt = fma(aa, rn, 1.0);
      rden = 1.0 / (aa + abv);
    }
    const double t = nz ? tt : 0.0, beta = nz ? -copysign(abv, alpha_p) : alpha_p, scal = nz ? copysign(rden, alpha_p) : 0.0;
    EIG_PROF(2)
    const double pv = t * (ek.y + scal * ((S1 + ek.x) + scal * S2));
    const double c = 0.5 * t * pv;
    const double ts = t * scal, cs = c * scal;
    const double wk = t * fma(scal, ek.x, ek.y) - c;         // w_k = p_k - c v_k, v_k = 1
    const double vk1 = scal * xk1;
    const double wk1 = t * fma(scal, ek1.x, ek1.y) - c * vk1;
    const double akk = ek.y - 2.0 * wk;
    const double alpha = ek1.y - wk1 - wk * vk1;
    const double wks = wk * scal;
    // ---- finish column k-1: w_j = p_j - c v_j = tau a_j + (tau scal) s_j - (c scal) x_j   (j > k; column k is dead)
    double w[NR];
#pragma unroll
    for (int m = M0; m < NR; ++m) w[m] = fma(-cs, xp[m], fma(ts, sp[m], t * x[m]));
    if (lane < kl) w[M0] = 0.0;         // columns < k: stale pairs of finished rows
#pragma unroll
    for (int q = Q0; q < RPW; ++q) {
      vi[q] *= scal;
      wi[q] = t * fma(scal, pr[q].x, pr[q].y) - c * vi[q];
      if (!act[q]) vi[q] = wi[q] = 0.0;                      // finished rows: the update below leaves them alone
    }
    vim *= scal;
    const double wim = t * fma(scal, prm.x, prm.y) - c * vim;
    bool recorder = false;              // one warp of the cluster (an active one: the owner of row k+1) records
#pragma unroll
    for (int q = Q0; q < RPW; ++q) recorder |= row[q] == k + 1;
    auto record_prev = [&]() {          // column k-1: reflector, tau, subdiagonal
      if (k > 0) {
#pragma unroll
        for (int m = M0; m < NR; ++m) {
          const int j = lane + 32 * m;
          if (j > k && j < n) V[(long)(k - 1) * ldv + j] = xp[m] * scal;
        }
        if (lane == 0) {
          e[k - 1] = beta;
          tau[k - 1] = t;
        }
      }
    };
    if (!build) {                       // last pass: rows n-2 (= k) and n-1 by reflector n-3; here column k is alive (v_k = 1)
      double v[NR];
#pragma unroll
      for (int m = M0; m < NR; ++m) v[m] = xp[m] * scal;
      if (lane == kl) {
        v[M0] = 1.0;
        w[M0] = wk;
      }
#pragma unroll
      for (int q = Q0; q < RPW; ++q) {
        if (row[q] >= k && row[q] < n) {
          const double vq = (row[q] == k) ? 1.0 : vi[q], wq = (row[q] == k) ? wk : wi[q];
#pragma unroll
          for (int m = M0; m < NR; ++m) a[q][m] = fma(-wq, v[m], fma(-vq, w[m], a[q][m]));
        }
      }
      if (recorder) record_prev();
      return false;
    }
    // ---- row k of the current matrix (== column k) beyond the subdiagonal, rebuilt from the received elements:
    //      x_j = a_j - w_j - w_k v_j
#pragma unroll
    for (int m = M0; m < NR; ++m) x[m] = (x[m] - w[m]) - wks * xp[m];
    if (lane <= kl + 1) x[M0] = 0.0;    // columns <= k+1
    if (wrap && lane == 0) x[M1] = 0.0;
    double red[8], redb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) red[i] = redb[i] = 0.0;
#pragma unroll
    for (int m = M0; m < NR; ++m) {
      if ((m - M0) & 1) {
#pragma unroll
        for (int q = Q0; q < RPW; ++q) redb[q] += a[q][m] * x[m];
        redb[4] += w[m] * x[m];
        redb[5] += xp[m] * x[m];
        redb[6] += x[m] * x[m];
      } else {
#pragma unroll
        for (int q = Q0; q < RPW; ++q) red[q] += a[q][m] * x[m];      // (A x)_i, rows as held
        red[4] += w[m] * x[m];
        red[5] += xp[m] * x[m];
        red[6] += x[m] * x[m];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[i] += redb[i];
    EIG_PROF(3)
    // ---- ONE reduction for the norm, the two correction sums and the row sums of A x
    const double tot = warp_reduce8_transposed(red, lane);             // lane L: total of red[L >> 2]
    const double sxm = __shfl_sync(0xffffffffu, tot, (4 * myq) & 31);  // (A x)_i of the row this lane sends
    const double wx = __shfl_sync(0xffffffffu, tot, 16), vx = scal * __shfl_sync(0xffffffffu, tot, 20);
    const double xnorm2 = __shfl_sync(0xffffffffu, tot, 24);
    EIG_PROF(4)
    // ---- send the raw pair of the row this lane looks after: (A' x)_i = (A x)_i - v_i (w.x) - w_i (v.x) and a'_{i,k+1},
    //      A' = the rows after the update by reflector k-1
    {
      const double an = aem - (vim * wk1 + wim * vk1);
      const double spn = (sxm - vim * wx) - wim * vx;
      if (sender && send_row > k) dsmem_store2_signal(remote + (pb ? 0u : PC1), spn, an, remote_bar + (pb ? 0u : 8u));
    }
    EIG_PROF(5)
    // ---- under the latency of the exchange: column k-1 is recorded, the own rows catch up with reflector k-1
    //      (columns j > k: a_ij -= v_i w_j + w_i scal x_j), the new unnormalised reflector is kept
    if (recorder) {
      record_prev();
      if (lane == 0) d[k] = akk;
    }
    EIG_PROF(6)
    if (k > 0) {
#pragma unroll
      for (int q = Q0; q < RPW; ++q) {
        const double wis = wi[q] * scal;
#pragma unroll
        for (int m = M0; m < NR; ++m) a[q][m] = fma(-wis, xp[m], fma(-vi[q], w[m], a[q][m]));
      }
    }
    EIG_PROF(7)
    __syncwarp();     // every lane has read vs[] (x of column k-1) above
#pragma unroll
    for (int m = M0; m < NR; ++m) {
      xp[m] = x[m];
      vs[lane + 32 * m] = x[m];
    }
    alpha_p = alpha;
    xn2_p = xnorm2;
    EIG_PROF(8)
    __syncwarp();     // vs[] written above is read (by other lanes of this warp) in the next column
    return true;
  };
  static_for<0, NR>([&](auto M0c) {
    constexpr int M0 = decltype(M0c)::value;
    const int kend = min(32 * (M0 + 1), n - 1);            // columns k = 0 .. n-2
    for (; live && k < kend; ++k) live = column(M0c);
  });
  EIG_PROF_DUMP(last_row == n - 1, 9)
  // last 2x2 block (rows n-2 and n-1 are final now)
#pragma unroll
  for (int q = 0; q < RPW; ++q) {
    if (row[q] == n - 2 || row[q] == n - 1) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int m = 0; m < NR; ++m) {
        if (lane + 32 * m == row[q]) s0 = a[q][m];
        if (lane + 32 * m == n - 1) s1 = a[q][m];
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
constexpr int DC_DIRECT_MAX = 16;                     // merge nodes up to this size use the one-thread-per-element update

struct DcSmem {
  // replicated in every CTA (each CTA runs the O(n) bookkeeping redundantly, bit-identically)
  double d[DC_MAXN], dn[DC_MAXN], z[DC_MAXN], dS[DC_MAXN], zS[DC_MAXN], dl[DC_MAXN], w[DC_MAXN], rot_c[DC_MAXN], rot_s[DC_MAXN];
  int srt[DC_MAXN], row[DC_MAXN], rot_p[DC_MAXN], rot_n[DC_MAXN], pid[DC_MAXN];
  int na[DC_MAXNODES], nc[DC_MAXNODES], nb[DC_MAXNODES], nk[DC_MAXNODES], nrot[DC_MAXNODES];
  double nrho[DC_MAXNODES];
  unsigned long long ndmax[DC_MAXNODES], nzmax[DC_MAXNODES];
  // broadcast by the computing warp into every CTA (distributed shared memory)
  double mu[DC_MAXN], zh[DC_MAXN];
  int org[DC_MAXN];
  // eigenvector-update tiles
  double As[DC_TK][DC_TM + 4], Bs[DC_TK][DC_TN + 4];   // row stride 68: conflict-free DMMA fragment reads (lane = 4 g + q reads [k0 + q][g])
  double nrm[DC_THREADS / DC_TM][DC_TM];
  double inv[DC_TM];
  double red[DC_WARPS];
};

// G adjacent lanes cooperate on one secular root / one z component: G = 32 for the large merges, 4 and 1 for the small ones
// at the bottom of the tree, where a 5-stage warp reduction per sum would cost more than the sums themselves
template <int G>
struct GroupLanes {
  static constexpr int L = G;
  __device__ __forceinline__ static int lane() { return threadIdx.x & (G - 1); }
  __device__ __forceinline__ static unsigned mask() {
    return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  }
  __device__ __forceinline__ static double sum(double x) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask(), x, o);
    return x;
  }
  __device__ __forceinline__ static double prod(double x) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x *= __shfl_xor_sync(mask(), x, o);
    return x;
  }
};

// Secular roots of the merge nodes covering positions [glo, ghi), spread over lane groups of G lanes.  `nshare` CTAs share
// the range (this one is number `me` of them); with nshare == 1 the results stay in this CTA's shared memory, otherwise
// (mu, org) are stored into every CTA of the cluster.
template <int G>
__device__ __forceinline__ void dc_secular_phase(DcSmem& S, cg::cluster_group& cluster, int glo, int ghi, int me, int nshare) {
  const int gid = (me * DC_THREADS + (int)threadIdx.x) / G, ngroups = nshare * DC_THREADS / G;
  const int gl = threadIdx.x & (G - 1);
  for (int g = glo + gid; g < ghi; g += ngroups) {
    const int p = S.pid[g], a = S.na[p], r = g - a, k = S.nk[p];
    if (r < k) {
      double mu;
      int org;
      dc::secular_root<GroupLanes<G>>(k, r, &S.dl[a], &S.w[a], S.nrho[p], mu, org);
      if (nshare == 1) {
        if (gl == 0) {
          S.mu[g] = mu;
          S.org[g] = org;
        }
      } else {
        for (int c = gl; c < DC_CLUSTER; c += G) {
          DcSmem* pr = cluster.map_shared_rank(&S, c);
          pr->mu[g] = mu;
          pr->org[g] = org;
        }
      }
    }
  }
}

// Gu-Eisenstat z of the same nodes
template <int G>
__device__ __forceinline__ void dc_zhat_phase(DcSmem& S, cg::cluster_group& cluster, int glo, int ghi, int me, int nshare) {
  const int gid = (me * DC_THREADS + (int)threadIdx.x) / G, ngroups = nshare * DC_THREADS / G;
  const int gl = threadIdx.x & (G - 1);
  for (int g = glo + gid; g < ghi; g += ngroups) {
    const int p = S.pid[g], a = S.na[p], r = g - a, k = S.nk[p];
    if (r < k) {
      const double zh = dc::zhat_component<GroupLanes<G>>(k, r, &S.dl[a], &S.w[a], &S.mu[a], &S.org[a]);
      if (nshare == 1) {
        if (gl == 0) S.zh[g] = zh;
      } else {
        for (int c = gl; c < DC_CLUSTER; c += G) cluster.map_shared_rank(&S, c)->zh[g] = zh;
      }
    }
  }
}

// grid = 8 * nmat CTAs, cluster (8,1,1).  d_in[nmat][n], e_in[nmat][n] with e_in[k] = T[k+1][k].  Qa, Qb: [nmat][n][ldq]
// workspaces.  Out: W[nmat][n] ascending, XT[nmat][n][ldx] rows = eigenvectors of T in the order of W.
// With Cmat != nullptr the LAST merge is left to the caller as a full-GPU GEMM: the kernel stops after the secular solve of
// the top level and writes its coefficient matrix Cmat[nmat][n][ldc] (rows already in ascending eigenvalue order, deflated
// rows = unit vectors) so that XT = Cmat * Q, Q = the level below (in Qa if ceil(log2 n) is odd, else Qb).
__global__ void __cluster_dims__(DC_CLUSTER, 1, 1) __launch_bounds__(DC_THREADS, 1)
    dc_cluster_kernel(int n, const double* __restrict__ d_in, const double* __restrict__ e_in, double* Qa, double* Qb, long ldq,
                      double* __restrict__ W, double* XT, long ldx, double* __restrict__ Cmat, long ldc,
                      int* __restrict__ info) {
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
  if (Cmat) Cmat += (long)mat * n * ldc;

  // ---- non-finite input (the reference lets NaN/inf flow and numpy.linalg.eigh then raises LinAlgError): every index
  //      computation below assumes a total order, so report it (info = 1, NaN outputs) and stop.  All CTAs agree.
  {
    int bad = 0;
    for (int i = tid; i < n; i += DC_THREADS) bad |= !isfinite(d_in[i]) || (i < n - 1 && !isfinite(e_in[i]));
    bad = __syncthreads_or(bad);
    if (info && rank == 0 && tid == 0) info[mat] = bad ? 1 : 0;
    if (bad) {
      const double qnan = __longlong_as_double(0x7ff8000000000000ll);
      double* out = Cmat ? Cmat : XT;
      const long ldo = Cmat ? ldc : ldx;
      for (int idx = rank * DC_THREADS + tid; idx < n * n; idx += DC_CLUSTER * DC_THREADS) {
        const int r = idx / n, c = idx - r * n;
        out[(long)r * ldo + c] = qnan;
      }
      if (rank == 0 && tid < n) W[tid] = qnan;
      return;
    }
  }
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
  EIG_PROF_DECL
#ifdef GPCSD_EIG_PROF
  int prof_level = 0;
#endif
  cluster.sync();
  EIG_PROF2(0)

  const int levels = dc::num_levels(n);
  bool shared_mode = false;
  int glo_prev = 0, ghi_prev = 0;
  for (int L = 1; L <= levels; ++L) {
#ifdef GPCSD_EIG_PROF
    const long long lvl_t0 = clock64();
    prof_level = L;
#endif
    const int nodes = 1 << (levels - L);
    const int mmax = (n + nodes - 1) / nodes;
    // While the level has at least one node per CTA, every CTA merges its own contiguous share of the nodes with no
    // communication at all (CTA barriers only); the levels above are shared by the whole cluster.
    const bool local = nodes >= DC_CLUSTER;
    if (!local && !shared_mode) {
      // first shared level: every CTA publishes the eigenvalues of its share (the eigenvector blocks are in global memory)
      if (tid < ghi_prev - glo_prev) {
        const double dv = S.d[glo_prev + tid];
        for (int c = 0; c < DC_CLUSTER; ++c) cluster.map_shared_rank(&S, c)->d[glo_prev + tid] = dv;
      }
      cluster.sync();
      shared_mode = true;
    }
    const int plo = local ? rank * (nodes / DC_CLUSTER) : 0, phi = local ? plo + nodes / DC_CLUSTER : nodes;
    const int glo = dc::node_start(n, nodes, plo), ghi = dc::node_start(n, nodes, phi);
    const int me = local ? 0 : rank, nshare = local ? 1 : DC_CLUSTER;
    glo_prev = glo;
    ghi_prev = ghi;
    const int g_own = glo + tid;                 // the position this thread looks after in the per-position phases
    const bool has_g = g_own < ghi;
    // ---- node table
    if (tid < phi - plo) {
      const int p = plo + tid;
      const int a = dc::node_start(n, 2 * nodes, 2 * p), c = dc::node_start(n, 2 * nodes, 2 * p + 1),
                b = dc::node_start(n, 2 * nodes, 2 * p + 2);
      S.na[p] = a;
      S.nc[p] = c;
      S.nb[p] = b;
      S.nrho[p] = (c == a || c == b) ? 0.0 : 2.0 * fabs(e_in[c - 1] / scale);
      S.ndmax[p] = 0ull;
      S.nzmax[p] = 0ull;
    }
    if (has_g) S.pid[g_own] = dc::node_of(n, nodes, g_own);
    __syncthreads();
    // ---- z = [last component of the left child's eigenvectors; sign(e_c) * first component of the right child's] / sqrt 2
    if (has_g) {
      const int p = S.pid[g_own], c = S.nc[p];
      double zz = 0.0;
      if (S.nrho[p] != 0.0) {
        const double q = __ldcg(Qold + (long)g_own * ldq + ((g_own < c) ? c - 1 : c));
        zz = ((g_own < c || e_in[c - 1] >= 0.0) ? q : -q) * 0.70710678118654752440;
      }
      S.z[g_own] = zz;
      atomicMax(&S.ndmax[p], (unsigned long long)__double_as_longlong(fabs(S.d[g_own])));
      atomicMax(&S.nzmax[p], (unsigned long long)__double_as_longlong(fabs(zz)));
    }
    __syncthreads();
    EIG_PROF2(1)
    // ---- counting sort of the node's eigenvalues
    if (has_g) {
      const int p = S.pid[g_own], a = S.na[p], b = S.nb[p];
      const double dg = S.d[g_own];
      int r = 0;
      for (int h = a; h < b; ++h) {
        const double dh = S.d[h];
        r += (dh < dg) || (dh == dg && h < g_own);
      }
      S.srt[a + r] = g_own;
      S.dS[a + r] = dg;
      S.zS[a + r] = S.z[g_own];
    }
    __syncthreads();
    EIG_PROF2(2)
    // ---- deflation, one thread per node
    if (tid < phi - plo) {
      const int p = plo + tid;
      const int a = S.na[p], m = S.nb[p] - a;
      int k = 0, nrot = 0;
      if (m > 0)
        dc::deflate(a, m, S.srt, S.dS, S.zS, S.nrho[p], __longlong_as_double((long long)S.ndmax[p]),
                    __longlong_as_double((long long)S.nzmax[p]), S.row, S.dl, S.w, S.rot_p, S.rot_n, S.rot_c, S.rot_s, k,
                    nrot);
      S.nk[p] = k;
      S.nrot[p] = nrot;
    }
    __syncthreads();
    EIG_PROF2(3)
    // ---- deflating Givens rotations on the rows of Qold: one thread per column, the running row stays in a register
    {
      // shared level: 32 columns per CTA (n <= 256); local level: the CTA's own columns
      const int col = local ? g_own : ((warp == 0) ? rank * 32 + lane : n);
      if (col < (local ? ghi : n)) {
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
    if (local) __syncthreads(); else cluster.sync();
    EIG_PROF2(4)
    // ---- secular roots
    if (local) {
      if (mmax <= 2) dc_secular_phase<1>(S, cluster, glo, ghi, me, nshare);
      else if (mmax <= 8) dc_secular_phase<4>(S, cluster, glo, ghi, me, nshare);
      else dc_secular_phase<16>(S, cluster, glo, ghi, me, nshare);
    } else {
      if (mmax <= 4) dc_secular_phase<1>(S, cluster, glo, ghi, me, nshare);
      else if (mmax <= 32) dc_secular_phase<4>(S, cluster, glo, ghi, me, nshare);
      else dc_secular_phase<32>(S, cluster, glo, ghi, me, nshare);
    }
    if (local) __syncthreads(); else cluster.sync();
    EIG_PROF2(5)
    // ---- Gu-Eisenstat z
    if (local) {
      if (mmax <= 2) dc_zhat_phase<1>(S, cluster, glo, ghi, me, nshare);
      else if (mmax <= 8) dc_zhat_phase<4>(S, cluster, glo, ghi, me, nshare);
      else dc_zhat_phase<16>(S, cluster, glo, ghi, me, nshare);
    } else {
      if (mmax <= 4) dc_zhat_phase<1>(S, cluster, glo, ghi, me, nshare);
      else if (mmax <= 32) dc_zhat_phase<4>(S, cluster, glo, ghi, me, nshare);
      else dc_zhat_phase<32>(S, cluster, glo, ghi, me, nshare);
    }
    if (local) __syncthreads(); else cluster.sync();
    EIG_PROF2(6)
    // ---- new eigenvalues (every CTA, identical)
    if (has_g) {
      const int p = S.pid[g_own], a = S.na[p], r = g_own - a;
      S.dn[g_own] = (r < S.nk[p]) ? S.dl[a + S.org[g_own]] + S.mu[g_own] : S.dl[g_own];
    }
    if (Cmat != nullptr && L == levels) {
      // ---- top level handed to the caller: final order, eigenvalues, coefficient matrix (one warp per output row)
      __syncthreads();
      if (tid < n) {
        const double dg = S.dn[tid];
        int r = 0;
        for (int h = 0; h < n; ++h) {
          const double dh = S.dn[h];
          r += (dh < dg) || (dh == dg && h < tid);
        }
        S.srt[tid] = r;                       // final position of merged row tid
        S.pid[S.row[tid]] = tid;              // old row -> its position in the merge (first k: kept, then deflated)
        if (rank == 0) W[r] = dg * scale;
      }
      __syncthreads();
      const int k = S.nk[0];
      for (int g = gw; g < n; g += GW) {
        double* out = Cmat + (long)S.srt[g] * ldc;
        if (g < k) {
          const int org_i = S.org[g];
          const double mu_i = S.mu[g];
          double nrm = 0.0;
          for (int j = lane; j < k; j += 32) {
            const double cf = S.zh[j] * dc::rcp(dc::delta_ji(S.dl, j, org_i, mu_i));
            nrm += cf * cf;
          }
          const double inv = rsqrt(warp_sum(nrm));
          for (int c = lane; c < n; c += 32) {
            const int pos = S.pid[c];
            out[c] = (pos < k) ? S.zh[pos] * dc::rcp(dc::delta_ji(S.dl, pos, org_i, mu_i)) * inv : 0.0;
          }
        } else {
          for (int c = lane; c < n; c += 32) out[c] = (S.pid[c] == g) ? 1.0 : 0.0;
        }
      }
      break;
    }
    // ---- eigenvector update: new row a+i = sum_j zh_j / (dl_j - lambda_i) / |.| * old row row[a+j]; deflated rows are copied
    if (mmax <= DC_DIRECT_MAX) {
      // small nodes: one thread per output element
      for (int idx = me * DC_THREADS + tid; idx < (ghi - glo) * mmax; idx += nshare * DC_THREADS) {
        const int g = glo + idx / mmax, cc = idx - (g - glo) * mmax;
        const int p = S.pid[g], a = S.na[p], m = S.nb[p] - a, r = g - a, k = S.nk[p];
        if (cc >= m) continue;
        const int col = a + cc;
        double out;
        if (r < k) {
          const int org_i = S.org[g];
          const double mu_i = S.mu[g];
          double acc = 0.0, nrm = 0.0;
          for (int j = 0; j < k; ++j) {
            const double cf = S.zh[a + j] * dc::rcp(dc::delta_ji(&S.dl[a], j, org_i, mu_i));
            nrm += cf * cf;
            acc += cf * __ldcg(Qold + (long)S.row[a + j] * ldq + col);
          }
          out = acc * rsqrt(nrm);
        } else {
          out = __ldcg(Qold + (long)S.row[g] * ldq + col);
        }
        Qnew[(long)g * ldq + col] = out;
      }
    } else {
      // 64 x 64 output tiles on the FP64 tensor path: 16 warps x (16 x 16) = 2 x 2 DMMA.8x8x4 tiles per warp.  (A first
      // version with a 2 x 4 register tile per thread read 48 bytes of shared memory per 8 FMAs and was bound by
      // shared-memory bandwidth: 6.8 k cycles per 16-deep k-step.)
      const int tr = (mmax + DC_TM - 1) / DC_TM, tc = (mmax + DC_TN - 1) / DC_TN;
      const int ntiles = nodes * tr * tc;
      const int wy = warp >> 2, wx = warp & 3;            // warp tile: rows 16 wy .., cols 16 wx ..
      const int fg = lane >> 2, fq = lane & 3;            // fragment coordinates
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
        double acc[2][2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < 2; ++q) acc[h][q][0] = acc[h][q][1] = 0.0;
        const int kend = (i0 < k) ? k : 0;                // tiles made of deflated rows only: no GEMM
        for (int j0 = 0; j0 < kend; j0 += DC_TK) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int kk = kq + 8 * h, j = j0 + kk;
            double cf = 0.0;
            if (vi && j < k) cf = S.zh[a + j] * dc::rcp(dc::delta_ji(&S.dl[a], j, org_i, mu_i));
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
          for (int ks = 0; ks < DC_TK / 4; ++ks) {
            double af[2], bf[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              af[h] = S.As[4 * ks + fq][16 * wy + 8 * h + fg];
              bf[h] = S.Bs[4 * ks + fq][16 * wx + 8 * h + fg];
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int q = 0; q < 2; ++q) dmma884(acc[h][q][0], acc[h][q][1], af[h], bf[q]);
          }
          __syncthreads();
        }
        S.nrm[kq][ii] = nrm_part;
        __syncthreads();
        if (tid < DC_TM) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < DC_THREADS / DC_TM; ++q) s += S.nrm[q][tid];
          S.inv[tid] = (s > 0.0) ? rsqrt(s) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rl = 16 * wy + 8 * h + fg, r = i0 + rl;
          if (r >= m) continue;
          const double sc = S.inv[rl];
          const long src = (long)S.row[a + r] * ldq + a;
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const int col = c0 + 16 * wx + 8 * q + 2 * fq + v;
              if (col < m) Qnew[(long)(a + r) * ldq + a + col] = (r < k) ? acc[h][q][v] * sc : __ldcg(Qold + src + col);
            }
        }
        __syncthreads();
      }
    }
    if (local) __syncthreads(); else cluster.sync();
    if (has_g) S.d[g_own] = S.dn[g_own];
    double* t = Qold;
    Qold = Qnew;
    Qnew = t;
    __syncthreads();
    EIG_PROF2(7)
#ifdef GPCSD_EIG_PROF
    prof_t[16 + L] = clock64() - lvl_t0;
#endif
  }

  if (Cmat != nullptr) {
    cluster.sync();
    return;
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
  EIG_PROF2(8)
  EIG_PROF_DUMP(rank == 0 && tid == 0 && mat == 0, 32)
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
  // The reflector rows are read two steps ahead of their use (registers): one step is a dot product, a 5-stage warp
  // reduction and an axpy (~300 cycles), an L2 round trip for the row is longer than that, and nothing but the loads is
  // independent of the previous step.  (Entries j <= k+1 of row k hold no reflector data: masked at use.)
  constexpr int NRB = TRD_MAXN / 32;
  auto load_row = [&](int k, double (&r)[NRB], double& t) {
    if (k < 0) return;
    const double* vk = V + (long)k * ldv;
    t = __ldg(tau + k);
#pragma unroll
    for (int m = 0; m < NRB; ++m) {
      const int j = lane + 32 * m;
      r[m] = (j < n) ? __ldg(vk + j) : 0.0;
    }
  };
  double ra[NRB], rb[NRB], rc[NRB], ta = 0.0, tb = 0.0, tc = 0.0;
#pragma unroll
  for (int m = 0; m < NRB; ++m) ra[m] = rb[m] = rc[m] = 0.0;
  load_row(n - 3, ra, ta);
  load_row(n - 4, rb, tb);
  for (int k = n - 3; k >= 0; --k) {
    load_row(k - 2, rc, tc);
    if (ta != 0.0) {
      double vv[NRB];
      double dot0 = 0.0, dot1 = 0.0;
#pragma unroll
      for (int m = 0; m < NRB; ++m) {
        const int j = lane + 32 * m;
        vv[m] = (j > k + 1) ? ra[m] : ((j == k + 1) ? 1.0 : 0.0);
        if (m & 1) dot1 += vv[m] * x[m]; else dot0 += vv[m] * x[m];
      }
      const double dot = warp_sum(dot0 + dot1) * ta;
#pragma unroll
      for (int m = 0; m < NRB; ++m) x[m] -= dot * vv[m];
    }
#pragma unroll
    for (int m = 0; m < NRB; ++m) {
      ra[m] = rb[m];
      rb[m] = rc[m];
    }
    ta = tb;
    tb = tc;
  }
#pragma unroll
  for (int m = 0; m < TRD_MAXN / 32; ++m) {
    const int j = lane + 32 * m;
    if (j < n) XT[(long)vec * ldx + j] = x[m];
  }
}

// stacked identity matrices [nmat][n][ld] (padding columns zero)
__global__ void identity_rows_kernel(int n, long ld, long total, double* __restrict__ R) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long c = idx % ld, r = (idx / ld) % n;
  R[idx] = (c == r) ? 1.0 : 0.0;
}

// ------------------------------------------------------------------------------------------------------------------------
// Orders 1..32: two-sided parallel Jacobi, ONE CTA per matrix (matrix and eigenvector matrix in shared memory).
// The cluster solver above spends ~155 us on a 24-order matrix (5 merge levels and 24 exchange round trips of fixed latency)
// and occupies 8 SMs per matrix, so the 64 + 128 factors of a 64-restart batch at 24 x 50 run in ~10 waves; here every matrix
// gets its own CTA and all of them run at once.  Round-robin ordering (circle method): each round applies ceil(n/2) rotations on
// disjoint index pairs -- rows first, then columns (and the columns of V) -- with thread (k, j) handling pair k, column / row
// j.  A rotation is skipped when |a_pq| <= max(8 eps sqrt(|a_pp a_qq|), eps/8 max|a_ij|): the absolute floor is the accuracy
// LAPACK and the cluster solver deliver (eps |K|); chasing the purely relative criterion inside the numerically degenerate
// jitter-level cluster of a GP factor costs 22 sweeps instead of ~8 for digits nothing downstream can use.  A sweep without
// rotations ends the iteration.
// Eigenvalues ascending, eigenvectors as ROWS of QT like the other solvers; info = 1 (NaN outputs) on non-finite input.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int JAC_MAXN = 32;

__global__ void __launch_bounds__(JAC_MAXN* JAC_MAXN / 2) jacobi_small_kernel(int n, const double* __restrict__ M, long ldm,
                                                                              double* __restrict__ QT, long ldq,
                                                                              double* __restrict__ W, int* __restrict__ info) {
  __shared__ double A[JAC_MAXN][JAC_MAXN + 1], V[JAC_MAXN][JAC_MAXN + 1], P[JAC_MAXN][JAC_MAXN + 1];
  __shared__ double cs[JAC_MAXN / 2], sn[JAC_MAXN / 2];
  __shared__ int pp[JAC_MAXN / 2], qq[JAC_MAXN / 2];
  __shared__ int rotated, bad;
  __shared__ double fro;
  const int mat = blockIdx.x;
  M += (long)mat * n * ldm;
  QT += (long)mat * n * ldq;
  W += (long)mat * n;
  const int m = (n + 1) / 2, np = 2 * m;            // pairs per round, padded order (index n is a dummy when n is odd)
  const int tid = threadIdx.x, k = tid / np, j = tid - k * np;   // launched with m * np threads
  if (tid == 0) {
    rotated = 0;
    bad = 0;
    fro = 0.0;
  }
  __syncthreads();
  for (int e = tid; e < np * np; e += blockDim.x) {
    const int r = e / np, c = e - r * np;
    double v = 0.0;
    if (r < n && c < n) {
      v = 0.5 * (M[(long)r * ldm + c] + M[(long)c * ldm + r]);
      if (!isfinite(v)) bad = 1;
    }
    A[r][c] = v;
    V[r][c] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (!bad && n >= 4) {
    // Pre-rotation by an orthonormal cosine basis C (DCT-II: C[i][k] = sqrt((k ? 2 : 1) / n) cos(pi (2i+1) k / (2n))): it
    // nearly diagonalises the (near-)Toeplitz covariance factors this solver is fed, so Jacobi starts close to diagonal
    // (5 sweeps instead of 9-12 on the GP spatial and temporal factors); for any other matrix it is a harmless orthogonal
    // similarity.  A <- C^T A C, V <- C; T = A C goes through registers.
    // Two candidates: DCT-II, and DCT-IV (C[i][k] = sqrt(2/n) cos(pi (2i+1)(2k+1) / (4n))).  The eigenvectors of a symmetric
    // Toeplitz matrix of order 2n are close to the DCT-II basis of order 2n, whose even members restricted to the first half
    // are the DCT-II basis of order n and whose odd members are the DCT-IV basis of order n: the symmetric half of the
    // centrosymmetric split is pre-rotated best by the first, the skew half by the second.  The kernel is not told which
    // matrix it holds; it keeps the candidate with the larger sum of squared diagonal entries of C^T A C (= the smaller
    // off-diagonal norm), formed in a fixed order so the choice is reproducible.
    double score[2];
    for (int cand = 0; cand < 2; ++cand) {
      for (int e = tid; e < n * n; e += blockDim.x) {
        const int i = e / n, c = e - i * n;
        V[i][c] = cand ? sqrt(2.0 / n) * cospi((double)((2 * i + 1) * (2 * c + 1)) / (double)(4 * n))
                       : sqrt((c ? 2.0 : 1.0) / n) * cospi((double)((2 * i + 1) * c) / (double)(2 * n));
      }
      __syncthreads();
      for (int e = tid; e < n * n; e += blockDim.x) {       // P[i][c] = C[i][c] * (A C)[i][c]
        const int i = e / n, c = e - i * n;
        double a0 = 0.0;
        for (int l = 0; l < n; ++l) a0 += A[i][l] * V[l][c];
        P[i][c] = a0 * V[i][c];
      }
      __syncthreads();
      if (tid < n) {                                        // diagonal entry tid of C^T A C
        double dg = 0.0;
        for (int i = 0; i < n; ++i) dg += P[i][tid];
        P[tid][JAC_MAXN] = dg * dg;                         // (padding column)
      }
      __syncthreads();
      double sc = 0.0;
      for (int i = 0; i < n; ++i) sc += P[i][JAC_MAXN];
      score[cand] = sc;
      __syncthreads();
    }
    if (!(score[1] > score[0])) {                           // keep DCT-II (V holds DCT-IV now)
      for (int e = tid; e < n * n; e += blockDim.x) {
        const int i = e / n, c = e - i * n;
        V[i][c] = sqrt((c ? 2.0 : 1.0) / n) * cospi((double)((2 * i + 1) * c) / (double)(2 * n));
      }
    }
    __syncthreads();
    double tacc[4];                                   // T[i][c] = sum_l A[i][l] C[l][c]; each thread owns <= 4 entries
    int cnt = 0;
    for (int e = tid; e < n * n; e += blockDim.x, ++cnt) {
      const int i = e / n, c = e - i * n;
      double a0 = 0.0;
      for (int l = 0; l < n; ++l) a0 += A[i][l] * V[l][c];
      tacc[cnt & 3] = a0;
    }
    __syncthreads();
    cnt = 0;
    for (int e = tid; e < n * n; e += blockDim.x, ++cnt) A[e / n][e % n] = tacc[cnt & 3];     // A now holds T
    __syncthreads();
    cnt = 0;
    for (int e = tid; e < n * n; e += blockDim.x, ++cnt) {     // A' [i][c] = sum_l C[l][i] T[l][c]
      const int i = e / n, c = e - i * n;
      double a0 = 0.0;
      for (int l = 0; l < n; ++l) a0 += V[l][i] * A[l][c];
      tacc[cnt & 3] = a0;
    }
    __syncthreads();
    cnt = 0;
    for (int e = tid; e < n * n; e += blockDim.x, ++cnt) A[e / n][e % n] = tacc[cnt & 3];
    __syncthreads();
  }
  if (tid == 0) {                                    // scale of the matrix (fixed order: results are bit-reproducible)
    double mx = 0.0;
    for (int i = 0; i < n; ++i)
      for (int c = 0; c <= i; ++c) mx = fmax(mx, fabs(A[i][c]));
    fro = mx;
  }
  __syncthreads();
  if (bad) {
    for (int e = tid; e < n * n; e += blockDim.x) QT[(long)(e / n) * ldq + e % n] = nan("");
    if (tid < n) W[tid] = nan("");
    if (tid == 0 && info) info[mat] = 1;
    return;
  }
  const double tiny = 2.7755575615628914e-17 * fro;     // eps/8 max|a_ij|: LAPACK-grade absolute accuracy, see the header comment
  // (V starts as the pre-rotation C, or the identity for n < 4 / after the load loop above)
  // Round-robin pairing in closed form (circle method on np players, np even): in round r player np-1 meets player r, and for
  // k = 1..m-1 player (r + k) mod (np-1) meets player (r - k) mod (np-1) -- m disjoint pairs, every pair once per sweep.
  // Three barriers per round: rotation parameters | row update | column update.
  const int nm1 = np - 1;
  for (int sweep = 0; sweep < 60; ++sweep) {
    for (int round = 0; round < nm1; ++round) {
      if (tid < m) {
        int a, b;
        if (tid == 0) {
          a = round;
          b = nm1;
        } else {
          a = round + tid;
          if (a >= nm1) a -= nm1;
          b = round - tid;
          if (b < 0) b += nm1;
        }
        const int p = a < b ? a : b, q = a < b ? b : a;
        pp[tid] = p;
        qq[tid] = q;
        double c = 1.0, sv = 0.0;
        if (q < n) {
          const double apq = 0.5 * (A[p][q] + A[q][p]), app = A[p][p], aqq = A[q][q];
          // rotate when |a_pq| > max(8 eps sqrt(|a_pp a_qq|), eps/8 max|a|), tested on squares (no sqrt on this chain)
          if (apq * apq > fmax(3.1554436208840472e-30 * fabs(app * aqq), tiny * tiny)) {
            // symmetric Schur rotation t = sign(D) 2 a_pq / (|D| + sqrt(D^2 + 4 a_pq^2)), D = a_qq - a_pp.  The angle is
            // formed in single precision (this chain sits on the critical path of every round: ~40 dependent FP64
            // operations otherwise); c = 1 / sqrt(1 + t^2), s = t c are then normalised in double, so the rotation is
            // orthogonal to double precision whatever the angle -- an inexact angle only leaves a_pq reduced by ~1e-7
            // instead of annihilated, which the next sweep finishes (the final sweeps see tiny angles anyway).
            const double dl = aqq - app, ap2 = 2.0 * apq;
            // common power-of-two scale (from the exponent bits of the larger magnitude) so the floats neither overflow nor flush
            const double big = fmax(fabs(dl), fabs(ap2));                 // > 0 here
            const int ex = ((__double2hiint(big) >> 20) & 0x7ff) - 1023;
            const double sc2 = __hiloint2double((1023 - max(-1000, min(1000, ex))) << 20, 0);        // 2^-ex
            const float df = (float)(dl * sc2), af = (float)(ap2 * sc2);
            const float tf = copysignf(1.0f, df) * af / (fabsf(df) + sqrtf(df * df + af * af));
            const double t = (double)tf;
            c = fast_rsqrt(1.0 + t * t);
            sv = t * c;
            rotated = 1;
          }
        }
        cs[tid] = c;
        sn[tid] = sv;
      }
      __syncthreads();
      const int p = pp[k], q = qq[k];
      const double c = cs[k], sv = sn[k];
      const bool act = (sv != 0.0);
      if (act) {               // rows p, q:  a_pj <- c a_pj - s a_qj,  a_qj <- s a_pj + c a_qj
        const double x = A[p][j], y = A[q][j];
        A[p][j] = c * x - sv * y;
        A[q][j] = sv * x + c * y;
      }
      __syncthreads();
      if (act) {               // columns p, q of A and of V
        const double x = A[j][p], y = A[j][q];
        A[j][p] = c * x - sv * y;
        A[j][q] = sv * x + c * y;
        const double u = V[j][p], w = V[j][q];
        V[j][p] = c * u - sv * w;
        V[j][q] = sv * u + c * w;
      }
      __syncthreads();
    }
    const int any = rotated;
    __syncthreads();
    if (tid == 0) rotated = 0;
    __syncthreads();
    if (!any) break;
  }
  // ascending order by counting ranks (ties by index), eigenvectors (columns of V) as rows of QT
  if (tid < n) {
    const double l = A[tid][tid];
    int rank = 0;
    for (int i = 0; i < n; ++i) {
      const double li = A[i][i];
      rank += (li < l || (li == l && i < tid)) ? 1 : 0;
    }
    W[rank] = l;
    V[tid][JAC_MAXN] = (double)rank;                  // the padding column of V carries the destination row
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, c = e - i * n;               // eigenvector i, component c
    QT[(long)((int)V[i][JAC_MAXN]) * ldq + c] = V[c][i];
  }
  if (tid == 0 && info) info[mat] = 0;
}

}  // namespace gpcsd

using namespace gpcsd;

extern "C" {
#ifdef GPCSD_EIG_PROF
int gpcsd_dbg_prof(long long* out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_eig_prof, sizeof(long long) * 64);
}
int gpcsd_dbg_lp(long long* out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_eig_lp, sizeof(long long) * 256);
}
int gpcsd_dbg_trace(long long* out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_eig_trace, sizeof(long long) * 8 * 16 * 12);
}
#endif


// Householder tridiagonalisation of `nmat` symmetric matrices of order n (3 <= n <= 256) on 8-CTA clusters.
int gpcsd_tridiag(int n, int nmat, const double* M, long ldm, double* d, double* e, double* V, long ldv, double* tau,
                  void* stream) {
  if (n < 3 || n > TRD_MAXN) return gp_fail("gpcsd_tridiag: order must be in 3..256");
  const int slots = (n + TRD_CLUSTER - 1) / TRD_CLUSTER;               // rows per CTA
  const int threads = 32 * (slots < TRD_WARPS ? slots : TRD_WARPS);
  const int grid = TRD_CLUSTER * nmat;
  cudaStream_t st = (cudaStream_t)stream;
  const int nr = (n + 31) / 32;                                        // register blocks per row (compile-time specialisation)
  if (nr <= 1) tridiag_cluster_kernel<1><<<grid, threads, 0, st>>>(n, M, ldm, d, e, V, ldv, tau);
  else if (nr <= 2) tridiag_cluster_kernel<2><<<grid, threads, 0, st>>>(n, M, ldm, d, e, V, ldv, tau);
  else if (nr <= 4) tridiag_cluster_kernel<4><<<grid, threads, 0, st>>>(n, M, ldm, d, e, V, ldv, tau);
  else if (nr <= 6) tridiag_cluster_kernel<6><<<grid, threads, 0, st>>>(n, M, ldm, d, e, V, ldv, tau);
  else tridiag_cluster_kernel<8><<<grid, threads, 0, st>>>(n, M, ldm, d, e, V, ldv, tau);
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
// info[nmat] (device, may be null): 0, or 1 when the matrix holds a non-finite entry (outputs are then NaN).
long gpcsd_tridiag_eig_ws_doubles(int n, long ldx, int nmat) { return 2L * nmat * n * ldx; }

static int dc_launch(int n, int nmat, const double* d, const double* e, double* Qa, double* Qb, long ldq, double* W, double* XT,
                     long ldx, double* Cmat, long ldc, int* info, cudaStream_t stream) {
  static int attr[GP_MAX_DEVICES];
  if (gp_first_use_on_device(attr)) {
    GP_CUDA(cudaFuncSetAttribute(dc_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DcSmem)));
  }
  dc_cluster_kernel<<<DC_CLUSTER * nmat, DC_THREADS, sizeof(DcSmem), stream>>>(n, d, e, Qa, Qb, ldq, W, XT, ldx, Cmat, ldc, info);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_tridiag_eig(int n, int nmat, const double* d, const double* e, double* W, double* XT, long ldx, double* ws,
                      long ws_doubles, int* info, void* stream) {
  if (n < 2 || n > DC_MAXN) return gp_fail("gpcsd_tridiag_eig: order must be in 2..256");
  if (ws_doubles < gpcsd_tridiag_eig_ws_doubles(n, ldx, nmat)) return gp_fail("gpcsd_tridiag_eig: workspace too small");
  return dc_launch(n, nmat, d, e, ws, ws + (long)nmat * n * ldx, ldx, W, XT, ldx, nullptr, 0, info, (cudaStream_t)stream);
}

// Full symmetric eigen-decomposition of `nmat` stacked matrices M[nmat][n][ldm] (3 <= n <= 256): QT[nmat][n][ldq] rows =
// eigenvectors, W[nmat][n] ascending.  Replaces np.linalg.eigh of utility_functions.py:58-59 for the factor orders that
// dominate GPCSD evaluations.  M is not modified.  ws: gpcsd_eigh_dc_ws_doubles(n, ldq, nmat) doubles.
//
// Orders above DC_EXTERNAL_MIN take the long tail off the critical path: while the divide-and-conquer kernel runs, a side
// stream forms H^T explicitly (the reflectors applied to the identity); the top-level merge and the back-transformation
// then are two full-GPU DMMA GEMMs,  QT = (C * Q_below) * H^T.  The side stream and its events are per caller stream.
constexpr int DC_EXTERNAL_MIN = 97;

long gpcsd_eigh_dc_ws_doubles(int n, long ldq, int nmat) { return (long)nmat * (6L * n * ldq + 3L * n); }

struct EighSide {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

int gpcsd_eigh_dc(int n, int nmat, const double* M, long ldm, double* QT, long ldq, double* W, double* ws, long ws_doubles,
                  int* info, void* stream) {
  if (n < 1 || n > DC_MAXN) return gp_fail("gpcsd_eigh_dc: order must be in 1..256");
  cudaStream_t st = (cudaStream_t)stream;
  // Orders <= 32: the Jacobi kernel -- one CTA per matrix, ~80 us (n = 24) to ~125 us (n = 32) WHATEVER the batch, against
  // ~140-155 us per wave of 18 matrices for the cluster solver below (8 SMs per matrix; profiles/r02i_small_eigh.md).
  if (n <= JAC_MAXN) {
    if (nmat <= 0) return 0;
    const int m = (n + 1) / 2;
    jacobi_small_kernel<<<nmat, m * 2 * m, 0, st>>>(n, M, ldm, QT, ldq, W, info);
    GP_CUDA(cudaGetLastError());
    return 0;
  }
  if (ws_doubles < gpcsd_eigh_dc_ws_doubles(n, ldq, nmat)) return gp_fail("gpcsd_eigh_dc: workspace too small");
  const long slab = (long)nmat * n * ldq, mstride = (long)n * ldq;
  double* V = ws;                     // reflectors
  double* Qa = V + slab;              // D&C ping-pong
  double* Qb = Qa + slab;
  double* Cm = Qb + slab;             // top-level coefficient matrix
  double* T1 = Cm + slab;             // eigenvectors of T
  double* R = T1 + slab;              // H^T
  double* d = R + slab;
  double* e = d + (long)nmat * n;
  double* tau = e + (long)nmat * n;
  if (gpcsd_tridiag(n, nmat, M, ldm, d, e, V, ldq, tau, stream)) return 1;
  if (n < DC_EXTERNAL_MIN) {
    if (dc_launch(n, nmat, d, e, Qa, Qb, ldq, W, QT, ldq, nullptr, 0, info, st)) return 1;
    return gpcsd_backtransform(n, nmat, V, ldq, tau, QT, ldq, stream);
  }
  // one side stream + event pair per CALLER stream (created on first use, kept for the life of the process)
  static std::mutex side_mutex;
  static std::unordered_map<unsigned long long, EighSide> sides;      // key: (device, caller stream)
  EighSide side;
  {
    std::lock_guard<std::mutex> lock(side_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    EighSide& slot = sides[((unsigned long long)(uintptr_t)st) ^ ((unsigned long long)(dev + 1) << 56)];
    if (!slot.stream) {
      GP_CUDA(cudaStreamCreateWithFlags(&slot.stream, cudaStreamNonBlocking));
      GP_CUDA(cudaEventCreateWithFlags(&slot.fork, cudaEventDisableTiming));
      GP_CUDA(cudaEventCreateWithFlags(&slot.join, cudaEventDisableTiming));
    }
    side = slot;
  }
  // side stream: R = H^T (rows H e_i)
  GP_CUDA(cudaEventRecord(side.fork, st));
  GP_CUDA(cudaStreamWaitEvent(side.stream, side.fork, 0));
  {
    const long total = (long)nmat * n * ldq;
    identity_rows_kernel<<<(int)((total + 255) / 256), 256, 0, side.stream>>>(n, ldq, total, R);
    GP_CUDA(cudaGetLastError());
  }
  if (gpcsd_backtransform(n, nmat, V, ldq, tau, R, ldq, side.stream)) return 1;
  GP_CUDA(cudaEventRecord(side.join, side.stream));
  // main stream: divide and conquer up to the top-level coefficients, then the two GEMMs
  if (dc_launch(n, nmat, d, e, Qa, Qb, ldq, W, nullptr, ldq, Cm, ldq, info, st)) return 1;
  const double* Qbelow = (dc::num_levels(n) & 1) ? Qa : Qb;
  if (gpcsd_dgemm(0, n, n, n, Cm, ldq, mstride, Qbelow, ldq, mstride, T1, ldq, mstride, nmat, stream)) return 1;
  GP_CUDA(cudaStreamWaitEvent(st, side.join, 0));
  return gpcsd_dgemm(0, n, n, n, T1, ldq, mstride, R, ldq, mstride, QT, ldq, mstride, nmat, stream);
}

}  // extern "C"
