// gpcsd_plan: ONE C-ABI call per loglik+gradient evaluation, batched over R hyperparameter vectors (multi-start restarts).
//
// Replaces the per-evaluation Python orchestration (~49 ctypes calls) of the reference's hot path -- obj_fun + autograd.grad
// inside the restart loop gpcsd1d.py:193-211 / gpcsd2d.py:223-260 -- by a native driver:
//   * the hyperparameters live in DEVICE memory (theta[R][P], natural units, the order of the gradient), so nothing in the
//     launch sequence depends on their values: the whole evaluation is captured once per (R, mode) into a CUDA graph and
//     replayed with one cudaGraphLaunch;
//   * every small kernel (forward-model weights, SE factors, Kt, symmetry splits/assemblies, 1/D and its reductions, gradient
//     cores, transposes, dot products, the final assembly) carries a restart dimension in its grid; dense products are
//     strided-batch DMMA GEMMs over the restarts; the eigensolver sees ONE stacked batch (gpcsd_eigh_dc, nmat = restarts x
//     blocks); only the trial-sized contractions (projection, SYRKs) are issued per restart;
//   * the result (loglik, gradient, solver flags) is assembled on the device and read back with one copy.
// The arithmetic is the one DESIGN.md section 3 derives and engine.py orchestrates call by call; the kernels that do the
// flops (gpcsd_dgemm, gpcsd_project_quad, gpcsd_wsyrk, gpcsd_eigh_dc) are shared with it.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>

#include <unordered_map>
#include <vector>

#include "common.h"
#include "dmma_gemm.cuh"

namespace gpcsd {
namespace plan {

constexpr int RESW = 32;   // per-restart scalar results: [0:2] quad,bsq (block s) [2:4] sum log D, sum 1/D [4] <dA,GU> [5:7] ell dots
                           // [8:8+2ntc] temporal dots [24:26] quad,bsq (block a) [26] info

struct Dims {
  int dim, nx, nt, G, G1, G2, ntc, nsig, nsp, P;
  int kinds[8];
};

__device__ __forceinline__ double kt_term(int kind, double ell, double d) {
  return kind == GPCSD_KIND_SE ? exp(-0.5 * d * d / (ell * ell)) : exp(-fabs(d) / ell);
}

// ---- covariance construction, restart dimension in blockIdx.y ---------------------------------------------------------
__global__ void fwd_weights_kernel(Dims d, const double* __restrict__ theta, const double* __restrict__ pts, int npts,
                                   const double* __restrict__ g1, const double* __restrict__ w1, const double* __restrict__ g2,
                                   const double* __restrict__ w2, double eps, double* __restrict__ A, double* __restrict__ dA,
                                   long sA) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)npts * d.G) return;
  const int r = blockIdx.y;
  const double R = theta[(long)r * d.P];
  const int i = (int)(idx / d.G), gq = (int)(idx % d.G);
  double a, da;
  if (d.dim == 1) {                                     // b_fwd_1d, forward_models.py:9-17 inside covariances.py:86-88
    const double u = (g1[gq] - pts[i]) / R, qd = u * u;
    const double s1 = sqrt(qd + 1.0), s0 = sqrt(qd), w = w1[gq];
    a = w * (s1 - s0);
    da = w * (qd / s1 - s0) * (-1.0 / R);
  } else {                                              // b_fwd_2d, forward_models.py:42-54 inside covariances.py:220-221
    const int a1 = gq / d.G2, a2 = gq % d.G2;
    const double d1 = g1[a1] - pts[2 * i], d2 = g2[a2] - pts[2 * i + 1];
    const double ww = d1 * d1 + d2 * d2, Re = R + eps, sR = sqrt(Re * Re + ww), wp = w1[a1] * w2[a2];
    a = wp * (log(Re + sR) - log(eps + sqrt(eps * eps + ww)));
    da = wp / sR;
  }
  A[(long)r * sA + idx] = a;
  if (dA) dA[(long)r * sA + idx] = da;
}

// K[r][i][j] = exp(-0.5 ((p_i - p_j)/ell_r)^2), dK = K * (p_i - p_j)^2 / ell^3   (covariances.py:89, 216 and d/d ell)
__global__ void se_pair_kernel(int n, const double* __restrict__ p, const double* __restrict__ theta, int P, int ell_idx,
                               double* __restrict__ K, double* __restrict__ dK, long ld, long sK) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int r = blockIdx.y;
  const double ell = theta[(long)r * P + ell_idx];
  const int i = (int)(idx / n), j = (int)(idx % n);
  const double dd = p[i] - p[j], u = dd / ell;
  const double v = exp(-0.5 * u * u);
  K[(long)r * sK + (long)i * ld + j] = v;
  if (dK) dK[(long)r * sK + (long)i * ld + j] = v * dd * dd / (ell * ell * ell);
}

__global__ void add_diag_kernel(int n, double* __restrict__ K, long ld, long sK, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) K[(long)blockIdx.y * sK + (long)i * ld + i] += v;
}

__global__ void kt_build_kernel(Dims d, const double* __restrict__ theta, const double* __restrict__ t, double* __restrict__ Kt,
                                long ld, long sK) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)d.nt * d.nt) return;
  const int r = blockIdx.y;
  const double* th = theta + (long)r * d.P + 1 + d.nsp;
  const int i = (int)(idx / d.nt), j = (int)(idx % d.nt);
  const double dt = t[i] - t[j];
  double v = 0.0;
  for (int k = 0; k < d.ntc; ++k) v += th[2 * k + 1] * kt_term(d.kinds[k], th[2 * k], dt);     // compute_Kt cov:257-305, summed 1d:118-120
  Kt[(long)r * sK + (long)i * ld + j] = v;
}

// per-block partials ws[r][block][16]; out via sum_cols_kernel
// marks the end of an evaluation's GEMM phase for the host (gpcsd_plan_loglik_grad's token): counter on the device, copy in
// host-mapped memory
__global__ void gemm_done_kernel(unsigned long long* cnt, unsigned long long* host_flag) {
  const unsigned long long v = *cnt + 1ull;
  *cnt = v;
  *reinterpret_cast<volatile unsigned long long*>(host_flag) = v;
  __threadfence_system();
}

__global__ void kt_grad_kernel(Dims d, const double* __restrict__ theta, const double* __restrict__ t, const double* __restrict__ Gm,
                               long ldg, long sG, double* __restrict__ ws, long sW) {
  __shared__ double red[8];
  const int r = blockIdx.y;
  const double* th = theta + (long)r * d.P + 1 + d.nsp;
  const double* G = Gm + (long)r * sG;
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  const long total = (long)d.nt * d.nt;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / d.nt), j = (int)(idx % d.nt);
    const double dt = t[i] - t[j];
    const double gv = G[(long)i * ldg + j];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < d.ntc) {
        const double ell = th[2 * k], s2 = th[2 * k + 1];
        const double e = kt_term(d.kinds[k], ell, dt);
        const double dl = d.kinds[k] == GPCSD_KIND_SE ? dt * dt / (ell * ell * ell) : fabs(dt) / (ell * ell);
        acc[2 * k] += gv * s2 * e * dl;
        acc[2 * k + 1] += gv * e;
      }
    }
  }
  for (int k = 0; k < 2 * d.ntc; ++k) {
    const double s = block_sum(acc[k], red);
    if (threadIdx.x == 0) ws[(long)r * sW + (long)blockIdx.x * 16 + k] = s;
  }
}

// out[r*so + k] = sum_b ws[r*sW + b*stride + k], k < ncols   (grid: (ncols, R); fixed order)
__global__ void sum_cols_kernel(const double* __restrict__ ws, long sW, int nblocks, int stride, double* __restrict__ out, long so) {
  __shared__ double red[8];
  const int k = blockIdx.x, r = blockIdx.y;
  double a = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) a += ws[(long)r * sW + (long)b * stride + k];
  const double s = block_sum(a, red);
  if (threadIdx.x == 0) out[(long)r * so + k] = s;
}

__global__ void dot_kernel(int rows, int cols, const double* __restrict__ X, long ldx, long sX, const double* __restrict__ Y,
                           long ldy, long sY, double* __restrict__ ws, long sW) {
  __shared__ double red[8];
  const int r = blockIdx.y;
  const double* Xr = X + (long)r * sX;
  const double* Yr = Y + (long)r * sY;
  double a = 0.0;
  const long total = (long)rows * cols;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long i = idx / cols, c = idx % cols;
    a += Xr[i * ldx + c] * Yr[i * ldy + c];
  }
  const double s = block_sum(a, red);
  if (threadIdx.x == 0) ws[(long)r * sW + blockIdx.x] = s;
}

// ---- symmetry splits (see gpcsd_kernels.cu for the single-matrix versions and the references) -------------------------
__global__ void centro_split_kernel(int n, const double* __restrict__ K, long ldk, long sK, double* __restrict__ S, long lds,
                                    long sS, double* __restrict__ A, long lda, long sA) {
  const int m = n / 2, odd = n & 1, ms = m + odd;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)ms * ms) return;
  const int r = blockIdx.y;
  const double* Kr = K + (long)r * sK;
  double* Sr = S + (long)r * sS;
  double* Ar = A + (long)r * sA;
  const int i = (int)(idx / ms), j = (int)(idx % ms);
  if (i < m && j < m) {
    const double a = Kr[(long)i * ldk + j], b = Kr[(long)i * ldk + (n - 1 - j)];
    Sr[(long)i * lds + j] = a + b;
    Ar[(long)i * lda + j] = a - b;
  } else if (i == m && j == m) {
    Sr[(long)i * lds + j] = Kr[(long)m * ldk + m];
  } else {
    const int q = (i == m) ? j : i;
    Sr[(long)i * lds + j] = 1.4142135623730951 * Kr[(long)q * ldk + m];
  }
}

__global__ void centro_assemble_kernel(int n, const double* __restrict__ UsT, long lds, long sUs, const double* __restrict__ Ws,
                                       long sWs, const double* __restrict__ UaT, long lda, long sUa, const double* __restrict__ Wa,
                                       long sWa, double* __restrict__ QT, long ldq, long sQ, double* __restrict__ W, long sW) {
  const int m = n / 2, odd = n & 1, ms = m + odd;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int r = blockIdx.y;
  const double* Us = UsT + (long)r * sUs;
  const double* Ua = UaT + (long)r * sUa;
  const int a = (int)(idx / n), c = (int)(idx % n);
  const double h = 0.70710678118654752;
  double v;
  if (a < ms) {
    if (c < m) v = h * Us[(long)a * lds + c];
    else if (odd && c == m) v = Us[(long)a * lds + m];
    else v = h * Us[(long)a * lds + (n - 1 - c)];
  } else {
    const int b = a - ms;
    if (c < m) v = h * Ua[(long)b * lda + c];
    else if (odd && c == m) v = 0.0;
    else v = -h * Ua[(long)b * lda + (n - 1 - c)];
  }
  QT[(long)r * sQ + (long)a * ldq + c] = v;
  if (c == 0) W[(long)r * sW + a] = (a < ms) ? Ws[(long)r * sWs + a] : Wa[(long)r * sWa + a - ms];
}

__global__ void pairsym_split_kernel(int m, const double* __restrict__ K, long ldk, long sK, const int* __restrict__ ra,
                                     const int* __restrict__ rb, double* __restrict__ S, long lds, long sS, double* __restrict__ A,
                                     long sA) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)m * m) return;
  const int r = blockIdx.y;
  const double* Kr = K + (long)r * sK;
  const int i = (int)(idx / m), j = (int)(idx % m);
  const double a = Kr[(long)ra[i] * ldk + ra[j]], b = Kr[(long)ra[i] * ldk + rb[j]];
  S[(long)r * sS + (long)i * lds + j] = a + b;
  A[(long)r * sA + (long)i * lds + j] = a - b;
}

__global__ void pairsym_assemble_kernel(int m, const int* __restrict__ ra, const int* __restrict__ rb, const double* __restrict__ UsT,
                                        long lds, long sUs, const double* __restrict__ Ws, long sWs, const double* __restrict__ UaT,
                                        long sUa, const double* __restrict__ Wa, long sWa, double* __restrict__ QT, long ldq, long sQ,
                                        double* __restrict__ W, long sW) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2L * m * m) return;
  const int r = blockIdx.y;
  const int a = (int)(idx / m), c = (int)(idx % m);
  const double h = 0.70710678118654752;
  double* Q = QT + (long)r * sQ;
  if (a < m) {
    const double v = h * UsT[(long)r * sUs + (long)a * lds + c];
    Q[(long)a * ldq + ra[c]] = v;
    Q[(long)a * ldq + rb[c]] = v;
  } else {
    const double v = h * UaT[(long)r * sUa + (long)(a - m) * lds + c];
    Q[(long)a * ldq + ra[c]] = v;
    Q[(long)a * ldq + rb[c]] = -v;
  }
  if (c == 0) W[(long)r * sW + a] = (a < m) ? Ws[(long)r * sWs + a] : Wa[(long)r * sWa + a - m];
}

// ---- D, 1/D and its reductions (utility_functions.py:54-63; gpcsd1d.py:122), grid (nx, R) / (ceil(nt/256), R) -------------
__global__ void eig_D_rows_kernel(Dims d, const double* __restrict__ theta, const double* __restrict__ ls, const double* __restrict__ lt,
                                  double* __restrict__ rD, long ldrd, double* __restrict__ rowA, double* __restrict__ rowC,
                                  double* __restrict__ rowL) {
  __shared__ double red[8];
  const int i = blockIdx.x, r = blockIdx.y;
  const double* sig = theta + (long)r * d.P + 1 + d.nsp + 2 * d.ntc;
  const double l = ls[(long)r * d.nx + i], s = (d.nsig == 1) ? sig[0] : sig[i];
  const double* ltr = lt + (long)r * d.nt;
  double a = 0.0, c = 0.0, lg = 0.0;
  for (int j = threadIdx.x; j < d.nt; j += blockDim.x) {
    const double D = l * ltr[j] + s;
    const double q = 1.0 / D;
    rD[((long)r * d.nx + i) * ldrd + j] = q;
    a += ltr[j] * q;
    c += q;
    lg += log(D);
  }
  const double sa = block_sum(a, red);
  const double sc = block_sum(c, red);
  const double sl = block_sum(lg, red);
  if (threadIdx.x == 0) {
    rowA[(long)r * d.nx + i] = sa;
    rowC[(long)r * d.nx + i] = sc;
    rowL[(long)r * d.nx + i] = sl;
  }
}

// colB[j] = sum_i ls_i / D_ij.  Block = 32 columns x 8 row groups (row group g sums i = g, g + 8, ... with four loads in
// flight; the eight group sums are added in order through shared memory): one thread per column walking all nx rows was a
// chain of nx dependent-latency loads on a single CTA (142 us at nx = 384).
__global__ void __launch_bounds__(256) eig_D_cols_kernel(Dims d, const double* __restrict__ ls, const double* __restrict__ rD,
                                                         long ldrd, const double* __restrict__ rowC,
                                                         const double* __restrict__ rowL, double* __restrict__ colB,
                                                         double* __restrict__ res) {
  __shared__ double part[8][32];
  const int tx = threadIdx.x & 31, gq = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx, r = blockIdx.y;
  const double* lsr = ls + (long)r * d.nx;
  const double* rDr = rD + (long)r * d.nx * ldrd;
  double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
  if (j < d.nt) {
    int i = gq;
    for (; i + 24 < d.nx; i += 32) {
      b0 += lsr[i] * rDr[(long)i * ldrd + j];
      b1 += lsr[i + 8] * rDr[(long)(i + 8) * ldrd + j];
      b2 += lsr[i + 16] * rDr[(long)(i + 16) * ldrd + j];
      b3 += lsr[i + 24] * rDr[(long)(i + 24) * ldrd + j];
    }
    for (; i < d.nx; i += 8) b0 += lsr[i] * rDr[(long)i * ldrd + j];
  }
  part[gq][tx] = (b0 + b1) + (b2 + b3);
  __syncthreads();
  if (gq == 0 && j < d.nt) {
    double b = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) b += part[q][tx];
    colB[(long)r * d.nt + j] = b;
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {          // warp 0: the two scalars (fixed lane order -> deterministic)
    double sl = 0.0, sc = 0.0;
    for (int i = threadIdx.x; i < d.nx; i += 32) {
      sl += rowL[(long)r * d.nx + i];
      sc += rowC[(long)r * d.nx + i];
    }
    sl = warp_sum(sl);
    sc = warp_sum(sc);
    if (threadIdx.x == 0) {
      res[(long)r * RESW + 2] = sl;
      res[(long)r * RESW + 3] = sc;
    }
  }
}

// ---- eigen-basis gradient cores --------------------------------------------------------------------
__global__ void grad_core_kernel(int n, const double* __restrict__ Mm, long ldm, long sM, const double* __restrict__ Nm,
                                 const double* __restrict__ lam, const double* __restrict__ theta, int P, int sig_idx,
                                 const double* __restrict__ rowsum, double ntot, double scale_det, double* __restrict__ X, long ldx,
                                 long sX) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int r = blockIdx.y;
  const int i = (int)(idx / n), j = (int)(idx % n);
  double v = 0.5 * Mm[(long)r * sM + (long)i * ldm + j];
  if (i == j) {
    v += -0.5 * ntot * scale_det * rowsum[(long)r * n + i];
  } else if (Nm) {
    const double dl = lam[(long)r * n + i] - lam[(long)r * n + j];
    const double* s = theta + (long)r * P + sig_idx;
    if (dl != 0.0) v += 0.5 * ((s[i] - s[j]) / dl) * Nm[(long)r * sM + (long)i * ldm + j];
  }
  X[(long)r * sX + (long)i * ldx + j] = v;
}

__global__ void transpose_kernel(int rows, int cols, const double* __restrict__ in, long ldi, long sI, double* __restrict__ out,
                                 long ldo, long sO) {
  __shared__ double tile[32][33];
  const double* inr = in + (long)blockIdx.z * sI;
  double* outr = out + (long)blockIdx.z * sO;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int rr = r0 + q, cc = c0 + threadIdx.x;
    if (rr < rows && cc < cols) tile[q][threadIdx.x] = inr[(long)rr * ldi + cc];
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int oc = r0 + threadIdx.x, orow = c0 + q;
    if (orow < cols && oc < rows) outr[(long)orow * ldo + oc] = tile[threadIdx.x][q];
  }
}

__global__ void zero_kernel(double* __restrict__ p, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = 0.0;
}

// info flags of the batched eigensolver calls -> res[r][26] (restart of matrix k in a call with nmat matrices: k % R)
__global__ void info_kernel(const int* __restrict__ info, int ninfo, int R, double* __restrict__ res) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double s = 0.0;
  for (int k = r; k < ninfo; k += R) s += (double)abs(info[k]);
  res[(long)r * RESW + 26] = s;
}

// ---- final assembly: out[r] = [loglik, d loglik / d theta (P entries), info, c * f, c^2 * f]; one thread per output entry.
// c is a fixed weighted checksum of theta[r], f = det_frac (1 / world for trial-sharded models): after the sum over ranks the
// last two entries are mean(c) and mean(c^2), whose variance exposes ranks that evaluated different hyperparameters.
__global__ void assemble_kernel(Dims d, const double* __restrict__ theta, const double* __restrict__ res,
                                const double* __restrict__ rowC, const double* __restrict__ Ns,
                                long ldns, long sNs, double ntot, double det_frac, int want_grad, double* __restrict__ out, int outw) {
  const int r = blockIdx.x;
  const double* q = res + (long)r * RESW;
  double* o = out + (long)r * outw;
  const double quad = q[0] + q[24], bsq = q[1] + q[25];
  for (int e = threadIdx.x; e < outw; e += blockDim.x) {
    double v = 0.0;
    if (e == 0) {
      v = -0.5 * ntot * det_frac * q[2] - 0.5 * quad;                       // gpcsd1d.py:122-128
    } else if (e == d.P + 1) {
      v = q[26];
    } else if (e >= d.P + 2) {
      v = theta[(long)gridDim.x * d.P + 2 * r + (e - d.P - 2)] * det_frac;     // checksum pair, computed by the host with theta
    } else if (want_grad) {
      const int k = e - 1;
      if (k == 0) v = 2.0 * q[4];
      else if (k < 1 + d.nsp) v = q[5 + (k - 1)];
      else if (k < 1 + d.nsp + 2 * d.ntc) v = q[8 + (k - 1 - d.nsp)];
      else {
        const int i = k - (1 + d.nsp + 2 * d.ntc);
        if (d.nsig == 1) v = -0.5 * ntot * det_frac * q[3] + 0.5 * bsq;
        else v = -0.5 * ntot * det_frac * rowC[(long)r * d.nx + i] + 0.5 * Ns[(long)r * sNs + (long)i * ldns + i];
      }
    }
    o[e] = v;
  }
}

// ---- result mailbox: this rank's [n] result vector -> its slot of a host-mapped shared segment, then the sequence flag.
// The ranks of a trial-sharded model read each other's slots on the host (gpcsd_plan_finish): the all-reduce of ~70 doubles
// costs one posted PCIe write per rank, no device-side waiting, no collective kernel and no device->host copy.
__global__ void publish_kernel(const double* __restrict__ out, int n, unsigned char* __restrict__ slot, long parity_bytes,
                               unsigned long long* __restrict__ seq_dev) {
  __shared__ unsigned long long seq;
  if (threadIdx.x == 0) seq = seq_dev[0] + 1ull;
  __syncthreads();
  unsigned char* base = slot + (seq & 1ull) * parity_bytes;
  volatile double* dst = reinterpret_cast<volatile double*>(base + 64);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = out[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned long long*>(base) = seq;     // data first, flag second (fence above)
    seq_dev[0] = seq;
  }
}

}  // namespace plan
}  // namespace gpcsd

using namespace gpcsd;
namespace pl = gpcsd::plan;

// ========================================================================================================================
// host side
// ========================================================================================================================
namespace {

inline long even(long n) { return (n + 1) / 2 * 2; }
inline unsigned blocks256(long n) { return (unsigned)((n + 255) / 256); }

struct Bump {
  char* base;
  size_t off;
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

constexpr int DC_MAX = 256;      // order limit of gpcsd_eigh_dc

struct Plan {
  pl::Dims d;
  double jitter, eps;
  int t_uniform, t_split, t_fold, has_pairs, s_split;
  int ldx, ldt, G1, G2p;         // G2p: 2-D second axis (padded even)
  int tm, tms, tlds, tlda;       // temporal split: orders m (skew) / ms (symmetric), leading dims
  int sm, sldm;                  // spatial split order nx/2
  int Rmax;
  // device geometry (owned)
  double *x, *t, *g1, *w1, *g2, *w2;
  int *ra, *rb;
  // data
  const double* Y;
  long ldn;
  int N;
  double ntot, det_frac;
  int yf_valid;
  // workspace
  char* ws;
  size_t ws_bytes;
  // per-restart arrays (R-major)
  double *theta, *out, *res;
  double *A, *dA, *U, *GU, *GA, *Wq, *Tk;             // [R][nx][G]
  double *Kg, *dKg, *K1, *dK1, *K2, *dK2;              // quadrature-grid kernels
  long ldk1, ldk2;
  double *Ks, *QsT, *Qs, *ls, *Xs, *T1s, *Gs, *Ms, *Ns;   // [R][nx][ldx]
  double *Kt, *QtT, *Qt, *lt, *Xt, *T1t, *Gt, *Mt;        // [R][nt][ldt]
  double *stS, *stT;                                   // split stacks: spatial [2][R][sm][sldm], temporal S [R][tms][tlds] + A [R][tm][tlda]
  double *uS, *wS, *uT, *wT;                           // eigensolver outputs on the stacks
  double *eigws;  long eigws_doubles;
  int* info;      int ninfo;
  double *rD, *rowA, *rowC, *rowL, *colB;
  double *Z, *Bm, *Yf;                                 // Z, Bm: [R][nx][nt][ldn]; Yf: the LFP in the evaluation basis [nx][nt][ldn]
  double *pq_ws;  long pq_ws_doubles;                  // per restart
  double *syrk_ws_t, *syrk_ws_s; long syrk_t_doubles, syrk_s_doubles;
  double *dot_ws; long dot_ws_doubles;                 // per restart
  double *ktg_ws; long ktg_ws_doubles;
  // streams / events
  cudaStream_t side[2];
  cudaStream_t own;                                   // every evaluation runs here (capturable whatever the caller's stream is)
  cudaEvent_t ev[8];
  // GEMM-phase token: a one-thread kernel after the last full-GPU kernel bumps gemm_cnt (device) and writes it to a
  // host-mapped word the host spins on (works inside a replayed graph, where an event-record node would not be seen by a
  // host-side wait issued before the node has executed)
  unsigned long long *gemm_flag_host, *gemm_flag_dev, *gemm_cnt;
  unsigned long long gemm_seq_host;
  cudaEvent_t ev_p1;                                  // end of phase 1 (recorded on `own` between the two launches)
  // token ordering (two-phase path): creation index of the plan, evaluations issued, "prologue in flight, token not yet taken"
  // (plain fields, read and written under g_registry_mutex)
  long model_index;
  long eval_count;
  int in_prologue;
  // host staging (pinned)
  double *h_theta, *h_out;
  // graphs
  std::unordered_map<long, cudaGraphExec_t> graphs;
  std::unordered_map<long, int> warmed;
  int use_graph;
  long launches;                                      // kernels launched by the last enqueue
  // result mailbox shared by the ranks of a trial-sharded model (host-mapped POSIX shared memory)
  unsigned char *mb_host, *mb_dev;
  long mb_slot_bytes, mb_parity_bytes;
  int world, rank;
  unsigned long long* seq_dev;
  unsigned long long seq_host;
};

// Registry of live plans for the token's grant order.  Ranks of a trial-sharded job must take the token in the SAME order
// (an evaluation only completes when every rank has finished it: if rank X runs model A's GEMM phase first and rank Y model
// B's, both models complete only after both phases everywhere and their next prologues start together -- the lock step the
// token exists to break).  Arrival order is timing; (evaluations issued, creation index) is the same on every rank of an
// SPMD job.  So a requester defers to any other plan whose prologue is in flight and whose (count, index) is smaller: that
// plan is certain to ask for the token within its prologue's duration, and it will not defer back.
static std::mutex g_registry_mutex;
static Plan* g_registry[64];
static long g_plans_created = 0;
static void registry_add(Plan* p) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  p->model_index = g_plans_created++;
  p->eval_count = 0;
  p->in_prologue = 0;
  for (auto& slot : g_registry)
    if (!slot) {
      slot = p;
      return;
    }
}
static void registry_remove(Plan* p) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  for (auto& slot : g_registry)
    if (slot == p) slot = nullptr;
}
static void set_token_state(Plan* p, int v) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  p->in_prologue = v;
}
static long next_eval_count(Plan* p) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  return ++p->eval_count;
}
// true while some other plan with a smaller (evaluations issued, creation index) has its prologue in flight
static bool must_defer(const Plan* p, long my_count) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  for (Plan* q : g_registry) {
    if (!q || q == p || !q->in_prologue) continue;
    const long qc = q->eval_count;
    if (qc < my_count || (qc == my_count && q->model_index < p->model_index)) return true;
  }
  return false;
}


int fail_plan(const char* m) { return gp_fail(m); }

}  // namespace

extern "C" {

int gpcsd_plan_create(void** out_plan, int dim, int nx, int nt, const double* h_x, const double* h_t, int G1, const double* h_g1,
                      const double* h_w1, int G2, const double* h_g2, const double* h_w2, int ntc, const int* h_kinds,
                      int n_sig2n, double jitter, double eps, int t_uniform, int fold_min_nt, const int* h_ra, const int* h_rb,
                      int max_restarts) {
  if (!out_plan) return fail_plan("plan_create: null output");
  if (dim != 1 && dim != 2) return fail_plan("plan_create: dim must be 1 or 2");
  if (nx < 1 || nt < 2 || G1 < 1 || (dim == 2 && G2 < 1)) return fail_plan("plan_create: bad sizes");
  if (ntc < 1 || ntc > 8) return fail_plan("plan_create: temporal covariance list must have 1..8 entries");
  if (n_sig2n != 1 && n_sig2n != nx) return fail_plan("plan_create: sig2n must be a scalar or have one entry per electrode");
  if (max_restarts < 1) return fail_plan("plan_create: max_restarts must be >= 1");
  const int Gtot = dim == 1 ? G1 : G1 * G2;
  if (Gtot & 1) return fail_plan("plan_create: the (last) quadrature axis must be padded to an even length");
  Plan* p = new Plan();
  memset(&p->d, 0, sizeof(p->d));
  p->d.dim = dim; p->d.nx = nx; p->d.nt = nt; p->d.G = Gtot; p->d.G1 = G1; p->d.G2 = dim == 2 ? G2 : 0;
  p->d.ntc = ntc; p->d.nsig = n_sig2n; p->d.nsp = dim == 1 ? 1 : 2;
  p->d.P = 1 + p->d.nsp + 2 * ntc + n_sig2n;
  for (int k = 0; k < ntc; ++k) {
    if (h_kinds[k] != GPCSD_KIND_SE && h_kinds[k] != GPCSD_KIND_MATERN) {
      delete p;
      return fail_plan("plan_create: unknown temporal kernel kind");
    }
    p->d.kinds[k] = h_kinds[k];
  }
  p->jitter = jitter; p->eps = eps;
  p->t_uniform = t_uniform ? 1 : 0;
  p->t_split = (p->t_uniform && nt >= 32 && (nt % 2 == 0)) ? 1 : 0;   // odd nt: unsplit solve (the engine path splits it)
  p->t_fold = (p->t_split && nt >= fold_min_nt) ? 1 : 0;
  p->has_pairs = (h_ra && h_rb) ? 1 : 0;
  p->s_split = (p->has_pairs && n_sig2n == 1 && nx >= 64 && (nx % 2 == 0)) ? 1 : 0;
  p->ldx = (int)even(nx); p->ldt = (int)even(nt);
  p->G1 = G1; p->G2p = dim == 2 ? G2 : 0;
  p->tm = nt / 2; p->tms = nt - nt / 2; p->tlds = (int)even(p->tms); p->tlda = (int)even(p->tm);
  p->sm = nx / 2; p->sldm = (int)even(p->sm);
  p->Rmax = max_restarts;
  p->Y = nullptr; p->ldn = 0; p->N = 0; p->ntot = 0.0; p->det_frac = 1.0; p->yf_valid = 0;
  p->ws = nullptr; p->ws_bytes = 0;
  p->use_graph = 1;
  p->launches = 0;
  p->mb_host = p->mb_dev = nullptr; p->mb_slot_bytes = p->mb_parity_bytes = 0; p->world = 1; p->rank = 0;
  p->seq_dev = nullptr; p->seq_host = 0;
  // geometry
  auto upload = [&](double** dst, const double* src, size_t n) -> int {
    GP_CUDA(cudaMalloc((void**)dst, (n ? n : 1) * sizeof(double)));
    if (n) GP_CUDA(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
  };
  p->x = p->t = p->g1 = p->w1 = p->g2 = p->w2 = nullptr;
  p->ra = p->rb = nullptr;
  int e = 0;
  e |= upload(&p->x, h_x, (size_t)nx * dim);
  e |= upload(&p->t, h_t, nt);
  e |= upload(&p->g1, h_g1, G1);
  e |= upload(&p->w1, h_w1, G1);
  e |= upload(&p->g2, h_g2, dim == 2 ? G2 : 0);
  e |= upload(&p->w2, h_w2, dim == 2 ? G2 : 0);
  if (p->has_pairs) {
    if (cudaMalloc((void**)&p->ra, (nx / 2) * sizeof(int)) != cudaSuccess || cudaMalloc((void**)&p->rb, (nx / 2) * sizeof(int)) != cudaSuccess) e = 1;
    if (!e && (cudaMemcpy(p->ra, h_ra, (nx / 2) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
               cudaMemcpy(p->rb, h_rb, (nx / 2) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)) e = 1;
  }
  const size_t outw = (size_t)p->d.P + 4;
  if (cudaMallocHost((void**)&p->h_theta, (size_t)max_restarts * (p->d.P + 2) * sizeof(double)) != cudaSuccess) e = 1;
  if (cudaMallocHost((void**)&p->h_out, (size_t)max_restarts * outw * sizeof(double)) != cudaSuccess) e = 1;
  for (int k = 0; k < 2; ++k)
    if (cudaStreamCreateWithFlags(&p->side[k], cudaStreamNonBlocking) != cudaSuccess) e = 1;
  if (cudaStreamCreateWithFlags(&p->own, cudaStreamNonBlocking) != cudaSuccess) e = 1;
  for (int k = 0; k < 8; ++k)
    if (cudaEventCreateWithFlags(&p->ev[k], cudaEventDisableTiming) != cudaSuccess) e = 1;
  p->gemm_flag_host = p->gemm_flag_dev = p->gemm_cnt = nullptr;
  p->gemm_seq_host = 0;
  if (cudaEventCreateWithFlags(&p->ev_p1, cudaEventDisableTiming) != cudaSuccess) e = 1;
  if (cudaHostAlloc((void**)&p->gemm_flag_host, 64, cudaHostAllocMapped) != cudaSuccess) e = 1;
  if (!e) {
    *p->gemm_flag_host = 0ull;
    if (cudaHostGetDevicePointer((void**)&p->gemm_flag_dev, p->gemm_flag_host, 0) != cudaSuccess) e = 1;
    if (cudaMalloc((void**)&p->gemm_cnt, sizeof(unsigned long long)) != cudaSuccess) e = 1;
    else if (cudaMemset(p->gemm_cnt, 0, sizeof(unsigned long long)) != cudaSuccess) e = 1;
  }
  if (e) return fail_plan("plan_create: device / pinned allocation failed");
  registry_add(p);                    // (only plans handed to the caller take part in the token's grant order)
  *out_plan = p;
  return 0;
}

int gpcsd_plan_num_params(void* plan) { return plan ? ((Plan*)plan)->d.P : -1; }
long gpcsd_plan_last_launches(void* plan) { return plan ? ((Plan*)plan)->launches : -1; }
int gpcsd_plan_set_graph(void* plan, int enable) {
  if (!plan) return fail_plan("null plan");
  ((Plan*)plan)->use_graph = enable ? 1 : 0;
  return 0;
}

static void drop_graphs(Plan* p) {
  for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
  p->graphs.clear();
  p->warmed.clear();
}

int gpcsd_plan_destroy(void* plan) {
  if (!plan) return 0;
  Plan* p = (Plan*)plan;
  registry_remove(p);
  drop_graphs(p);
  cudaFree(p->x); cudaFree(p->t); cudaFree(p->g1); cudaFree(p->w1); cudaFree(p->g2); cudaFree(p->w2);
  cudaFree(p->ra); cudaFree(p->rb);
  cudaFreeHost(p->h_theta); cudaFreeHost(p->h_out);
  if (p->mb_host) cudaHostUnregister(p->mb_host);
  cudaFree(p->seq_dev);
  for (int k = 0; k < 2; ++k) cudaStreamDestroy(p->side[k]);
  cudaStreamDestroy(p->own);
  for (int k = 0; k < 8; ++k) cudaEventDestroy(p->ev[k]);
  cudaEventDestroy(p->ev_p1);
  if (p->gemm_flag_host) cudaFreeHost(p->gemm_flag_host);
  if (p->gemm_cnt) cudaFree(p->gemm_cnt);
  delete p;
  return 0;
}

// carve the workspace; base == nullptr only measures
static size_t layout(Plan* p, char* base, long ldn, int N) {
  Bump b{base, 0};
  const pl::Dims& d = p->d;
  const long R = p->Rmax, nx = d.nx, nt = d.nt, G = d.G, ldx = p->ldx, ldt = p->ldt;
  const int outw = d.P + 4;
  p->theta = b.take<double>(R * (d.P + 2));
  p->out = b.take<double>(R * outw);
  p->res = b.take<double>(R * pl::RESW);
  p->A = b.take<double>(R * nx * G); p->dA = b.take<double>(R * nx * G); p->U = b.take<double>(R * nx * G);
  p->GU = b.take<double>(R * nx * G); p->GA = b.take<double>(R * nx * G); p->Wq = b.take<double>(R * nx * G);
  p->Tk = b.take<double>(R * nx * G);
  if (d.dim == 1) {
    p->ldk1 = G; p->ldk2 = 0;
    p->Kg = b.take<double>(R * G * G); p->dKg = b.take<double>(R * G * G);
    p->K1 = p->dK1 = p->K2 = p->dK2 = nullptr;
  } else {
    p->ldk1 = even(d.G1); p->ldk2 = even(d.G2);
    p->K1 = b.take<double>(R * d.G1 * p->ldk1); p->dK1 = b.take<double>(R * d.G1 * p->ldk1);
    p->K2 = b.take<double>(R * d.G2 * p->ldk2); p->dK2 = b.take<double>(R * d.G2 * p->ldk2);
    p->Kg = p->dKg = nullptr;
  }
  double** sx[] = {&p->Ks, &p->QsT, &p->Qs, &p->Xs, &p->T1s, &p->Gs, &p->Ms, &p->Ns};
  for (auto q : sx) *q = b.take<double>(R * nx * ldx);
  p->ls = b.take<double>(R * nx);
  double** tx[] = {&p->Kt, &p->QtT, &p->Qt, &p->Xt, &p->T1t, &p->Gt, &p->Mt};
  for (auto q : tx) *q = b.take<double>(R * nt * ldt);
  p->lt = b.take<double>(R * nt);
  // split stacks + eigensolver outputs
  p->stS = b.take<double>(2 * R * (long)p->sm * p->sldm); p->uS = b.take<double>(2 * R * (long)p->sm * p->sldm);
  p->wS = b.take<double>(2 * R * p->sm + 2);
  p->stT = b.take<double>(R * ((long)p->tms * p->tlds + (long)p->tm * p->tlda));
  p->uT = b.take<double>(R * ((long)p->tms * p->tlds + (long)p->tm * p->tlda));
  p->wT = b.take<double>(R * nt + 2);
  long ew = 0;
  auto need = [&](int n, long ld, long nmat) {
    long w = (n >= 3 && n <= DC_MAX) ? gpcsd_eigh_dc_ws_doubles(n, ld, (int)nmat) : (n > DC_MAX ? gpcsd_eigh_ws_doubles(n, ld) : 0);
    if (w > ew) ew = w;
  };
  if (p->s_split) need(p->sm, p->sldm, 2 * R); else need(d.nx, ldx, R);
  if (p->t_split) need(p->tm, p->tlds, 2 * R); else need(d.nt, ldt, R);
  ew = (ew + 31) / 32 * 32;                         // the second solver's scratch starts at eigws + ew: keep it 16-byte aligned
  p->eigws_doubles = ew;
  p->eigws = b.take<double>(2 * ew + 2);            // two independent solver calls may be in flight (spatial || temporal)
  p->ninfo = (int)(6 * R);
  p->info = b.take<int>(p->ninfo);
  p->rD = b.take<double>(R * nx * ldt);
  p->rowA = b.take<double>(R * nx); p->rowC = b.take<double>(R * nx); p->rowL = b.take<double>(R * nx);
  p->colB = b.take<double>(R * nt);
  const long slab = nx * nt * ldn;
  p->Z = b.take<double>(R * slab); p->Bm = b.take<double>(R * slab);
  p->Yf = (p->s_split || p->t_fold) ? b.take<double>(slab) : nullptr;
  const int Nn = N > 0 ? N : 1;
  long pq = 0;
  for (int o : {d.nt, p->tm, p->tms}) { long w = gpcsd_project_quad_batched_ws_doubles((int)R, d.nx, o, Nn); if (w > pq) pq = w; }
  p->pq_ws_doubles = pq; p->pq_ws = b.take<double>(pq);
  long st = 2, ss = 2;
  for (int o : {d.nt, p->tm, p->tms}) { long w = gpcsd_wsyrk_batched_ws_doubles((int)R, o, d.nx, Nn, 0); if (w > st) st = w; }
  for (int o : {d.nx, p->sm > 0 ? p->sm : 1}) { long w = gpcsd_wsyrk_batched_ws_doubles((int)R, o, d.nt, Nn, 1); if (w > ss) ss = w; }
  p->syrk_t_doubles = st; p->syrk_s_doubles = ss;
  p->syrk_ws_t = b.take<double>(st); p->syrk_ws_s = b.take<double>(ss);
  p->dot_ws_doubles = 4L * gp_num_sms() + 8; p->dot_ws = b.take<double>(R * p->dot_ws_doubles);
  p->ktg_ws_doubles = 16L * (4L * gp_num_sms() + 8); p->ktg_ws = b.take<double>(R * p->ktg_ws_doubles);
  return (b.off + 255) & ~size_t(255);
}

long gpcsd_plan_ws_bytes(void* plan, long ldn, int ntrials_local) {
  if (!plan) return -1;
  Plan* p = (Plan*)plan;
  Plan tmp = *p;                      // measure on a copy: pointers of a bound plan stay valid
  tmp.graphs.clear(); tmp.warmed.clear();
  return (long)layout(&tmp, nullptr, ldn, ntrials_local);
}

// Bind this rank's trial slab Y[nx][nt][ldn] (device, caller-owned, zero padded to ldn) and the workspace.
// det_fraction: share of the trial-independent log-det terms this rank contributes (1/world for trial-sharded models).
int gpcsd_plan_set_lfp(void* plan, const double* Y, long ldn, int ntrials_local, double ntrials_total, double det_fraction,
                       void* ws, long ws_bytes) {
  if (!plan) return fail_plan("null plan");
  Plan* p = (Plan*)plan;
  if (ldn < ntrials_local || (ldn & 1)) return fail_plan("plan_set_lfp: ldn must be even and >= ntrials");
  if (((uintptr_t)Y | (uintptr_t)ws) & 15) return fail_plan("plan_set_lfp: pointers must be 16-byte aligned");
  const size_t need = layout(p, nullptr, ldn, ntrials_local);
  if ((size_t)ws_bytes < need) return fail_plan("plan_set_lfp: workspace too small (gpcsd_plan_ws_bytes)");
  const bool same = (p->Y == Y && p->ldn == ldn && p->N == ntrials_local && p->ws == (char*)ws && p->ntot == ntrials_total &&
                     p->det_frac == det_fraction);
  p->Y = Y; p->ldn = ldn; p->N = ntrials_local; p->ntot = ntrials_total; p->det_frac = det_fraction;
  p->ws = (char*)ws; p->ws_bytes = (size_t)ws_bytes;
  layout(p, p->ws, ldn, ntrials_local);
  p->yf_valid = 0;                    // new data: the channel-folded copy is rebuilt by the next evaluation
  if (!same) drop_graphs(p);
  return 0;
}

// tell the plan the data in the bound Y buffer changed in place (same pointer / shape): refolds Yf at the next evaluation
int gpcsd_plan_touch_lfp(void* plan) {
  if (!plan) return fail_plan("null plan");
  ((Plan*)plan)->yf_valid = 0;
  return 0;
}

double* gpcsd_plan_device_result(void* plan) { return plan ? ((Plan*)plan)->out : nullptr; }
double* gpcsd_plan_device_theta(void* plan) { return plan ? ((Plan*)plan)->theta : nullptr; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------------
// the evaluation
// ------------------------------------------------------------------------------------------------------------------------
namespace {

#define PL_CHECK(call)            \
  do {                            \
    if (int _e = (call)) return _e; \
  } while (0)
#define PL_LAUNCH(n)              \
  do {                            \
    GP_CUDA(cudaGetLastError());  \
    p->launches += (n);           \
  } while (0)

// eigen-factors of `nmat` stacked matrices of order n; orders outside the cluster solver's range go to cuSOLVER one by one
int eigh_stack(Plan* p, int n, int nmat, const double* M, long ld, double* QT, double* W, double* ws, int* info, cudaStream_t st) {
  if (n >= 3 && n <= DC_MAX) {
    PL_CHECK(gpcsd_eigh_dc(n, nmat, M, ld, QT, ld, W, ws, p->eigws_doubles, info, st));
    p->launches += (n >= 97) ? 6 : 3;
    return 0;
  }
  const long lw = gpcsd_eigh_ws_doubles(n, ld);
  if (lw < 0) return 1;
  // cuSOLVER workspace: carved from the solver scratch if it fits, else this order is not supported by the plan
  if (lw > p->eigws_doubles) return fail_plan("plan: eigensolver workspace too small for the cuSOLVER branch");
  for (int k = 0; k < nmat; ++k) {
    PL_CHECK(gpcsd_eigh(n, M + (long)k * n * ld, ld, QT + (long)k * n * ld, ld, W + (long)k * n, ws, lw, info + k, st));
    p->launches += 1;
  }
  return 0;
}

bool needs_cusolver(const Plan* p) {
  auto big = [](int n) { return n > DC_MAX || n < 3; };
  const bool sp = p->s_split ? big(p->sm) : big(p->d.nx);
  const bool tp = p->t_split ? (big(p->tm) || big(p->tms)) : big(p->d.nt);
  return sp || tp;
}

// G = Q X Q^T for R stacked problems of order n (QT rows = eigenvectors): Q = QT^T, T1 = X QT, G = Q T1
int rotate(Plan* p, int R, int n, long ld, const double* X, const double* QT, double* Q, double* T1, double* G, cudaStream_t st) {
  dim3 grid((n + 31) / 32, (n + 31) / 32, R), block(32, 8);
  pl::transpose_kernel<<<grid, block, 0, st>>>(n, n, QT, ld, (long)n * ld, Q, ld, (long)n * ld);
  PL_LAUNCH(1);
  PL_CHECK(gpcsd_dgemm(0, n, n, n, X, ld, (long)n * ld, QT, ld, (long)n * ld, T1, ld, (long)n * ld, R, st));
  PL_CHECK(gpcsd_dgemm(0, n, n, n, Q, ld, (long)n * ld, T1, ld, (long)n * ld, G, ld, (long)n * ld, R, st));
  p->launches += 2;
  return 0;
}

// out[r] = X[r] (nx x G) * (quadrature-grid kernel of restart r); 1-D dense, 2-D Kronecker K1 (x) K2 (covariances.py:216)
int apply_quad_kernel(Plan* p, int R, const double* X, const double* Kd, const double* K1, const double* K2, double* out,
                      cudaStream_t st) {
  const pl::Dims& d = p->d;
  const long nx = d.nx, G = d.G;
  if (d.dim == 1) {
    PL_CHECK(gpcsd_dgemm(0, (int)nx, (int)G, (int)G, X, G, nx * G, Kd, G, G * G, out, G, nx * G, R, st));
    p->launches += 1;
    return 0;
  }
  const int G1 = d.G1, G2 = d.G2;
  // T[(i,a), b'] = sum_b X[(i,a), b] K2[b, b']
  PL_CHECK(gpcsd_dgemm(0, (int)(nx * G1), G2, G2, X, G2, nx * G, K2, p->ldk2, (long)G2 * p->ldk2, p->Tk, G2, nx * G, R, st));
  p->launches += 1;
  // out_i[a', b'] = sum_a K1[a', a] T_i[a, b'], batched over the nx rows; K1 differs per restart
  for (int r = 0; r < R; ++r) {
    PL_CHECK(gpcsd_dgemm(0, G1, G2, G1, K1 + (long)r * G1 * p->ldk1, p->ldk1, 0, p->Tk + r * nx * G, G2, G, out + r * nx * G, G2, G,
                         (int)nx, st));
    p->launches += 1;
  }
  return 0;
}

// C[r] = A[r] * B for r < R with ONE shared B operand (Z = Qs^T Y): a single strided launch on the small-M kernel; the TMA
// kernel's tensor map cannot express a zero batch stride, so there the restarts are issued one by one
int gemm_shared_b(Plan* p, int R, int M, int N, int K, const double* A, long lda, long sA, const double* B, long ldb, double* C,
                  long ldc, long sC, cudaStream_t st) {
  if (M <= 32 || R == 1) {
    PL_CHECK(gpcsd_dgemm(0, M, N, K, A, lda, sA, B, ldb, 0, C, ldc, sC, R, st));
    p->launches += 1;
    return 0;
  }
  for (int r = 0; r < R; ++r) {
    PL_CHECK(gpcsd_dgemm(0, M, N, K, A + r * sA, lda, 0, B, ldb, 0, C + r * sC, ldc, 0, 1, st));
    p->launches += 1;
  }
  return 0;
}

struct Factors {          // caller-supplied eigen-factors (device, R == 1): QsT [nx][ldx], ls, QtT [nt][ldt], lt
  const double *QsT, *ls, *QtT, *lt;
};

// everything between "theta is on the device" and "out[r] is assembled", on `st` (+ the plan's side streams).
// phase 0: the whole evaluation.  phase 1: up to and including 1/D -- covariance build, eigendecompositions, Z = Qs^T Yf:
// latency-bound work on a few SMs (plus one HBM-bound pass); phase 2: the rest -- projection, SYRKs, gradient assembly: the
// kernels that fill the GPU.  The two phases are separate launches (separate CUDA graphs) only when several models are
// evaluated concurrently, so that a host-side token can keep their GEMM phases from interleaving (gpcsd_plan_loglik_grad).
int enqueue_body(Plan* p, int R, int want_grad, const Factors* fac, cudaStream_t st, int phase = 0) {
  const pl::Dims& d = p->d;
  const int nx = d.nx, nt = d.nt, G = d.G, N = p->N;
  const long ldx = p->ldx, ldt = p->ldt, ldn = p->ldn, row = (long)nt * ldn, slab = (long)nx * row;
  const bool vec = d.nsig > 1;
  const bool use_ssplit = p->s_split && !fac, use_tsplit = p->t_split && !fac, use_tfold = p->t_fold && !fac;
  cudaStream_t sS = p->side[0];
  const int sig_idx = 1 + d.nsp + 2 * d.ntc;
  const double *QsT = fac ? fac->QsT : p->QsT, *ls = fac ? fac->ls : p->ls, *QtT = fac ? fac->QtT : p->QtT, *lt = fac ? fac->lt : p->lt;
  if (phase != 2) p->launches = 0;

  // Z = Qs^T Yf for all trials (the one HBM-bound full-GPU pass that does not need Qt).  One-launch evaluations issue it on
  // the spatial side stream underneath the temporal eigensolve; two-phase evaluations (concurrent models) issue it at the
  // head of the GEMM phase instead: squeezed onto the SMs a concurrent model's persistent GEMM leaves free it took ~0.4 ms.
  auto enqueue_Z = [&](cudaStream_t zs) -> int {
    if (N > 0) {
      // The symmetry folds act on the DATA axes and commute with the projections: the LFP is moved to the evaluation basis
      // (channel fold for reflection-symmetric geometries, centrosymmetric time fold on uniform grids) ONCE per upload, so
      // Z = Qs^T Y comes out directly in the folded time basis -- no per-evaluation fold pass over the trial data.
      if ((p->s_split || p->t_fold) && !p->yf_valid) {
        if (p->s_split && p->t_fold) {
          PL_CHECK(gpcsd_pairsym_fold(nx, p->ra, p->rb, row, p->Y, p->Z, zs));       // Z (restart 0) is free here: scratch
          PL_CHECK(gpcsd_centro_fold(nx, nt, ldn, p->Z, p->Yf, zs));
          p->launches += 2;
        } else if (p->s_split) {
          PL_CHECK(gpcsd_pairsym_fold(nx, p->ra, p->rb, row, p->Y, p->Yf, zs));
          p->launches += 1;
        } else {
          PL_CHECK(gpcsd_centro_fold(nx, nt, ldn, p->Y, p->Yf, zs));
          p->launches += 1;
        }
      }
      const double* Ysrc = (p->s_split || p->t_fold) ? p->Yf : p->Y;
      if (use_ssplit) {
        const int m = p->sm;
        const long ldm = p->sldm, sM = (long)m * ldm;
        PL_CHECK(gemm_shared_b(p, R, m, (int)row, m, p->uS, ldm, sM, Ysrc, row, p->Z, row, slab, zs));
        PL_CHECK(gemm_shared_b(p, R, m, (int)row, m, p->uS + R * sM, ldm, sM, Ysrc + (long)m * row, row, p->Z + (long)m * row, row, slab, zs));
      } else {
        PL_CHECK(gemm_shared_b(p, R, nx, (int)row, nx, p->QsT, ldx, nx * ldx, Ysrc, row, p->Z, row, slab, zs));
      }
    }
    return 0;
  };

  auto prologue = [&]() -> int {
  pl::zero_kernel<<<1, 256, 0, st>>>(p->res, (long)R * pl::RESW);
  PL_LAUNCH(1);

  // ---------------- spatial covariance  Ks = (A Kg) A^T + jitter I   (covariances.py:74-96 / 204-232; gpcsd1d.py:117)
  pl::fwd_weights_kernel<<<dim3(blocks256((long)nx * G), R), 256, 0, st>>>(d, p->theta, p->x, nx, p->g1, p->w1, p->g2, p->w2, p->eps,
                                                                         p->A, want_grad ? p->dA : nullptr, (long)nx * G);
  PL_LAUNCH(1);
  if (d.dim == 1) {
    pl::se_pair_kernel<<<dim3(blocks256((long)G * G), R), 256, 0, st>>>(G, p->g1, p->theta, d.P, 1, p->Kg, want_grad ? p->dKg : nullptr,
                                                                        G, (long)G * G);
    PL_LAUNCH(1);
  } else {
    pl::se_pair_kernel<<<dim3(blocks256((long)d.G1 * d.G1), R), 256, 0, st>>>(d.G1, p->g1, p->theta, d.P, 1, p->K1,
                                                                              want_grad ? p->dK1 : nullptr, p->ldk1, (long)d.G1 * p->ldk1);
    PL_LAUNCH(1);
    pl::se_pair_kernel<<<dim3(blocks256((long)d.G2 * d.G2), R), 256, 0, st>>>(d.G2, p->g2, p->theta, d.P, 2, p->K2,
                                                                              want_grad ? p->dK2 : nullptr, p->ldk2, (long)d.G2 * p->ldk2);
    PL_LAUNCH(1);
  }
  PL_CHECK(apply_quad_kernel(p, R, p->A, p->Kg, p->K1, p->K2, p->U, st));
  PL_CHECK(gpcsd_dgemm(1, nx, nx, G, p->U, G, (long)nx * G, p->A, G, (long)nx * G, p->Ks, ldx, nx * ldx, R, st));
  p->launches += 1;
  if (p->jitter != 0.0) {
    pl::add_diag_kernel<<<dim3(blocks256(nx), R), 256, 0, st>>>(nx, p->Ks, ldx, nx * ldx, p->jitter);
    PL_LAUNCH(1);
  }

  int ninfo = 0;
  if (fac) {
    if (N > 0) {
      PL_CHECK(gpcsd_dgemm(0, nx, (int)row, nx, QsT, ldx, 0, p->Y, row, 0, p->Z, row, 0, 1, st));
      p->launches += 1;
    }
  } else {
    // ---------------- spatial eigen-factors + Z = Qs^T Y (+ fold) on a side stream, underneath the temporal eigensolve
    GP_CUDA(cudaEventRecord(p->ev[0], st));
    GP_CUDA(cudaStreamWaitEvent(sS, p->ev[0], 0));
    if (use_ssplit) {
      const int m = p->sm;
      const long ldm = p->sldm, sM = (long)m * ldm;
      pl::pairsym_split_kernel<<<dim3(blocks256((long)m * m), R), 256, 0, sS>>>(m, p->Ks, ldx, nx * ldx, p->ra, p->rb, p->stS, ldm, sM,
                                                                               p->stS + R * sM, sM);
      PL_LAUNCH(1);
      PL_CHECK(eigh_stack(p, m, 2 * R, p->stS, ldm, p->uS, p->wS, p->eigws, p->info + ninfo, sS));
      ninfo += 2 * R;
      pl::pairsym_assemble_kernel<<<dim3(blocks256(2L * m * m), R), 256, 0, sS>>>(m, p->ra, p->rb, p->uS, ldm, sM, p->wS, m,
                                                                                 p->uS + R * sM, sM, p->wS + (long)R * m, m, p->QsT, ldx,
                                                                                 nx * ldx, p->ls, nx);
      PL_LAUNCH(1);
    } else {
      PL_CHECK(eigh_stack(p, nx, R, p->Ks, ldx, p->QsT, p->ls, p->eigws, p->info + ninfo, sS));
      ninfo += R;
    }
    if (phase == 0) PL_CHECK(enqueue_Z(sS));
    GP_CUDA(cudaEventRecord(p->ev[1], sS));

    // ---------------- temporal covariance + eigen-factors (main stream)
    pl::kt_build_kernel<<<dim3(blocks256((long)nt * nt), R), 256, 0, st>>>(d, p->theta, p->t, p->Kt, ldt, nt * ldt);
    PL_LAUNCH(1);
    if (use_tsplit) {
      const int m = p->tm, ms = p->tms;
      const long lds = p->tlds, lda = p->tlda, sS_ = (long)ms * lds, sA_ = (long)m * lda;
      double* stA = p->stT + R * sS_;
      double* uA = p->uT + R * sS_;
      double* wA = p->wT + (long)R * ms;
      pl::centro_split_kernel<<<dim3(blocks256((long)ms * ms), R), 256, 0, st>>>(nt, p->Kt, ldt, nt * ldt, p->stT, lds, sS_, stA, lda, sA_);
      PL_LAUNCH(1);
      double* ews = p->eigws + p->eigws_doubles;
      // even nt (the plan only splits even orders): ONE batched solver call for the 2R half-order blocks, stacked
      // [S x R][A x R] with the same leading dimension
      PL_CHECK(eigh_stack(p, m, 2 * R, p->stT, lds, p->uT, p->wT, ews, p->info + ninfo, st));
      ninfo += 2 * R;
      pl::centro_assemble_kernel<<<dim3(blocks256((long)nt * nt), R), 256, 0, st>>>(nt, p->uT, lds, sS_, p->wT, ms, uA, lda, sA_, wA, m,
                                                                                  p->QtT, ldt, nt * ldt, p->lt, nt);
      PL_LAUNCH(1);
    } else {
      PL_CHECK(eigh_stack(p, nt, R, p->Kt, ldt, p->QtT, p->lt, p->eigws + p->eigws_doubles, p->info + ninfo, st));
      ninfo += R;
    }
    GP_CUDA(cudaStreamWaitEvent(st, p->ev[1], 0));
    pl::info_kernel<<<blocks256(R), 256, 0, st>>>(p->info, ninfo, R, p->res);
    PL_LAUNCH(1);
  }

  // ---------------- D, 1/D and reductions  (comp_eig_D, utility_functions.py:54-63)
  if (fac) {
    // factors were uploaded by the caller: single restart, arrays are the caller's
  }
  pl::eig_D_rows_kernel<<<dim3(nx, R), 256, 0, st>>>(d, p->theta, ls, lt, p->rD, ldt, p->rowA, p->rowC, p->rowL);
  PL_LAUNCH(1);
  pl::eig_D_cols_kernel<<<dim3((nt + 31) / 32, R), 256, 0, st>>>(d, ls, p->rD, ldt, p->rowC, p->rowL, p->colB, p->res);
  PL_LAUNCH(1);
  return 0;
  };   // prologue
  if (phase != 2) PL_CHECK(prologue());
  if (phase == 1) return 0;
  if (phase == 2) PL_CHECK(enqueue_Z(st));
  // (phase 2 only) a host-visible mark right after the last full-GPU kernel: the host releases the GEMM token there
  auto mark_gemm_done = [&]() -> int {
    if (phase != 2) return 0;
    pl::gemm_done_kernel<<<1, 1, 0, st>>>(p->gemm_cnt, p->gemm_flag_dev);
    PL_LAUNCH(1);
    return 0;
  };

  // ---------------- projection A_i = Qt^T Z_i with the fused /D + quadratic form (hot loop gpcsd1d.py:124-126)
  if (N > 0) {
    // all restarts x all spatial eigen-indices in ONE launch per block (gpcsd_project_quad_batched)
    if (use_tfold) {
      const int m = p->tm, ms = p->tms;
      const long lds = p->tlds, sS_ = (long)ms * lds;
      PL_CHECK(gpcsd_project_quad_batched(R, nx, ms, N, p->uT, lds, sS_, p->Z, ldn, row, p->rD, ldt, p->Bm, p->pq_ws, p->res + 0,
                                          pl::RESW, st));
      PL_CHECK(gpcsd_project_quad_batched(R, nx, m, N, p->uT + R * sS_, p->tlda, (long)m * p->tlda, p->Z + (long)ms * ldn, ldn, row,
                                          p->rD + ms, ldt, p->Bm + (long)ms * ldn, p->pq_ws, p->res + 24, pl::RESW, st));
      p->launches += 4;
    } else {
      PL_CHECK(gpcsd_project_quad_batched(R, nx, nt, N, QtT, ldt, fac ? 0 : (long)nt * ldt, p->Z, ldn, row, p->rD, ldt, p->Bm, p->pq_ws,
                                          p->res + 0, pl::RESW, st));
      p->launches += 2;
    }
  }

  if (!want_grad) PL_CHECK(mark_gemm_done());
  if (want_grad) {
    // The temporal branch (Mt SYRK -> core -> rotation -> <Gt, dKt>) and the spatial branch (Ms/Ns SYRKs -> core -> rotation
    // -> <Gs, dKs>) only share the projected data Bm: they run on two streams (two parallel branches of the captured graph).
    cudaStream_t sG = p->side[1];
    GP_CUDA(cudaEventRecord(p->ev[2], st));
    GP_CUDA(cudaStreamWaitEvent(sG, p->ev[2], 0));
    const long nMt = (long)R * nt * ldt, nMs = (long)R * nx * ldx;
    const bool blockT = use_tfold, blockS = use_ssplit;
    auto zero = [&](double* ptr, long n, cudaStream_t s_) -> int {
      pl::zero_kernel<<<blocks256(n) < 1024 ? blocks256(n) : 1024, 256, 0, s_>>>(ptr, n);
      PL_LAUNCH(1);
      return 0;
    };
    // ================= temporal branch (main stream)
    if (blockT || N == 0) PL_CHECK(zero(p->Mt, nMt, st));
    if (N > 0) {   // all restarts in one launch (gpcsd_wsyrk_batched: blockIdx.z / 4th tensor-map dimension = restart)
      if (blockT) {     // only the two diagonal blocks of Mt enter <dL/dKt, dKt/dtheta>
        const int m = p->tm, ms = p->tms;
        PL_CHECK(gpcsd_wsyrk_batched(R, ms, nx, N, p->Bm, ldn, row, slab, ls, nx, p->Mt, nullptr, ldt, nt * ldt, p->syrk_ws_t, st));
        PL_CHECK(gpcsd_wsyrk_batched(R, m, nx, N, p->Bm + (long)ms * ldn, ldn, row, slab, ls, nx, p->Mt + (long)ms * ldt + ms, nullptr,
                                     ldt, nt * ldt, p->syrk_ws_t, st));
        p->launches += 4;
      } else {
        PL_CHECK(gpcsd_wsyrk_batched(R, nt, nx, N, p->Bm, ldn, row, slab, ls, nx, p->Mt, nullptr, ldt, nt * ldt, p->syrk_ws_t, st));
        p->launches += 2;
      }
    }
    PL_CHECK(mark_gemm_done());
    pl::grad_core_kernel<<<dim3(blocks256((long)nt * nt), R), 256, 0, st>>>(nt, p->Mt, ldt, nt * ldt, nullptr, lt, p->theta, d.P, sig_idx,
                                                                          p->colB, p->ntot, p->det_frac, p->Xt, ldt, nt * ldt);
    PL_LAUNCH(1);
    PL_CHECK(rotate(p, R, nt, ldt, p->Xt, QtT, p->Qt, p->T1t, p->Gt, st));
    {   // <Gt, dKt_k/d(ell_k, sigma2_k)>
      long nb = ((long)nt * nt + 255) / 256;
      const long cap = 4L * gp_num_sms();
      if (R > 1 && nb > 8) nb = nb < cap / R + 1 ? nb : cap / R + 1;
      if (nb > cap) nb = cap;
      if (nb < 1) nb = 1;
      pl::kt_grad_kernel<<<dim3((unsigned)nb, R), 256, 0, st>>>(d, p->theta, p->t, p->Gt, ldt, nt * ldt, p->ktg_ws, p->ktg_ws_doubles);
      PL_LAUNCH(1);
      pl::sum_cols_kernel<<<dim3(2 * d.ntc, R), 256, 0, st>>>(p->ktg_ws, p->ktg_ws_doubles, (int)nb, 16, p->res + 8, pl::RESW);
      PL_LAUNCH(1);
    }
    // ================= spatial branch (side stream)
    if (blockS || N == 0) PL_CHECK(zero(p->Ms, nMs, sG));
    if (vec && N == 0) PL_CHECK(zero(p->Ns, nMs, sG));
    if (N > 0) {
      if (blockS) {
        const int mh = p->sm;
        PL_CHECK(gpcsd_wsyrk_batched(R, mh, nt, N, p->Bm, row, ldn, slab, lt, nt, p->Ms, nullptr, ldx, nx * ldx, p->syrk_ws_s, sG));
        PL_CHECK(gpcsd_wsyrk_batched(R, mh, nt, N, p->Bm + (long)mh * row, row, ldn, slab, lt, nt, p->Ms + (long)mh * ldx + mh, nullptr,
                                     ldx, nx * ldx, p->syrk_ws_s, sG));
        p->launches += 4;
      } else {         // per-electrode noise: Ms and Ns from one pass over Bm (nx <= 32)
        PL_CHECK(gpcsd_wsyrk_batched(R, nx, nt, N, p->Bm, row, ldn, slab, lt, nt, p->Ms, vec ? p->Ns : nullptr, ldx, nx * ldx,
                                     p->syrk_ws_s, sG));
        p->launches += (vec && nx > 32) ? 4 : 2;
      }
    }
    pl::grad_core_kernel<<<dim3(blocks256((long)nx * nx), R), 256, 0, sG>>>(nx, p->Ms, ldx, nx * ldx, vec ? p->Ns : nullptr, ls, p->theta,
                                                                          d.P, sig_idx, p->rowA, p->ntot, p->det_frac, p->Xs, ldx, nx * ldx);
    PL_LAUNCH(1);
    PL_CHECK(rotate(p, R, nx, ldx, p->Xs, QsT, p->Qs, p->T1s, p->Gs, sG));
    // dL/dR = 2 <dA, Gs U>,  dL/dell_k = <A, (Gs A) dKg_k>
    auto dot = [&](const double* X, const double* Yv, int slot) -> int {
      long nb = ((long)nx * G + 255) / 256;
      const long cap = 4L * gp_num_sms();
      if (nb > cap) nb = cap;
      pl::dot_kernel<<<dim3((unsigned)nb, R), 256, 0, sG>>>(nx, G, X, G, (long)nx * G, Yv, G, (long)nx * G, p->dot_ws, p->dot_ws_doubles);
      PL_LAUNCH(1);
      pl::sum_cols_kernel<<<dim3(1, R), 256, 0, sG>>>(p->dot_ws, p->dot_ws_doubles, (int)nb, 1, p->res + slot, pl::RESW);
      PL_LAUNCH(1);
      return 0;
    };
    PL_CHECK(gpcsd_dgemm(0, nx, G, nx, p->Gs, ldx, nx * ldx, p->U, G, (long)nx * G, p->GU, G, (long)nx * G, R, sG));
    p->launches += 1;
    PL_CHECK(dot(p->dA, p->GU, 4));
    PL_CHECK(gpcsd_dgemm(0, nx, G, nx, p->Gs, ldx, nx * ldx, p->A, G, (long)nx * G, p->GA, G, (long)nx * G, R, sG));
    p->launches += 1;
    if (d.dim == 1) {
      PL_CHECK(apply_quad_kernel(p, R, p->GA, p->dKg, nullptr, nullptr, p->Wq, sG));
      PL_CHECK(dot(p->A, p->Wq, 5));
    } else {
      PL_CHECK(apply_quad_kernel(p, R, p->GA, nullptr, p->dK1, p->K2, p->Wq, sG));
      PL_CHECK(dot(p->A, p->Wq, 5));
      PL_CHECK(apply_quad_kernel(p, R, p->GA, nullptr, p->K1, p->dK2, p->Wq, sG));
      PL_CHECK(dot(p->A, p->Wq, 6));
    }
    GP_CUDA(cudaEventRecord(p->ev[3], sG));
    GP_CUDA(cudaStreamWaitEvent(st, p->ev[3], 0));
  }
  pl::assemble_kernel<<<R, 64, 0, st>>>(d, p->theta, p->res, p->rowC, p->Ns, ldx, nx * ldx, p->ntot, p->det_frac, want_grad, p->out,
                                        d.P + 4);
  PL_LAUNCH(1);
  if (p->mb_dev && !fac) {
    pl::publish_kernel<<<1, 256, 0, st>>>(p->out, R * (d.P + 4), p->mb_dev + (long)p->rank * p->mb_slot_bytes, p->mb_parity_bytes,
                                          p->seq_dev);
    PL_LAUNCH(1);
  }
  if ((p->s_split || p->t_fold) && !fac && N > 0) p->yf_valid = 1;
  return 0;
}

}  // namespace

extern "C" {

// Upload R hyperparameter vectors (host, natural units, [R][P] in the gradient's order R, ell(s), (ell_t, sigma2_t)..., sig2n[...])
// and enqueue the whole evaluation on `stream`; the partial result of this rank is left in gpcsd_plan_device_result()
// ([R][P+4]: loglik, gradient, solver flag, hyperparameter checksum pair) for an optional all-reduce by the caller before gpcsd_plan_finish.
// theta[R][P] followed by the checksum pairs (c, c^2)[R], c = sum_k sin(1 + k) log theta_k, into the pinned staging buffer
static void stage_theta(Plan* p, int R, const double* h_theta) {
  const int P = p->d.P;
  memcpy(p->h_theta, h_theta, (size_t)R * P * sizeof(double));
  for (int r = 0; r < R; ++r) {
    double c = 0.0;
    for (int k = 0; k < P; ++k) {
      const double v = h_theta[(size_t)r * P + k];
      c += sin(1.0 + k) * log(v > 1e-300 ? v : 1e-300);
    }
    p->h_theta[(size_t)R * P + 2 * r] = c;
    p->h_theta[(size_t)R * P + 2 * r + 1] = c * c;
  }
}

static int enqueue_on_own(Plan* p, int R, int want_grad, size_t nb, int phase = 0) {
  cudaStream_t st = p->own;
  const long key = ((long)R * 2 + (want_grad ? 1 : 0)) * 4 + phase;
  // (the channel-folded copy Yf is rebuilt eagerly after every new upload; graphs are recorded without that step)
  const bool graphable = p->use_graph && !needs_cusolver(p) && (!(p->s_split || p->t_fold) || p->yf_valid || p->N == 0);
  if (graphable) {
    auto it = p->graphs.find(key);
    if (it != p->graphs.end()) {
      GP_CUDA(cudaGraphLaunch(it->second, st));
      return 0;
    }
    if (p->warmed[key] >= 1) {
      // second use of this (R, mode): capture the launch sequence (theta upload included) and replay it from now on
      cudaGraph_t g = nullptr;
      GP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int e = 0;
      if (phase != 2 && cudaMemcpyAsync(p->theta, p->h_theta, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) e = 1;
      if (!e) e = enqueue_body(p, R, want_grad, nullptr, st, phase);
      cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (e || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        p->use_graph = 0;                       // fall back to eager launches for the life of this plan
      } else {
        cudaGraphExec_t ge = nullptr;
        if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
          cudaGraphDestroy(g);
          cudaGetLastError();
          p->use_graph = 0;
        } else {
          cudaGraphDestroy(g);
          p->graphs[key] = ge;
          GP_CUDA(cudaGraphLaunch(ge, st));
          return 0;
        }
      }
    }
  }
  if (phase != 2) GP_CUDA(cudaMemcpyAsync(p->theta, p->h_theta, nb, cudaMemcpyHostToDevice, st));
  static const bool trace = getenv("GPCSD_PLAN_TRACE") != nullptr;
  if (trace) {
    cudaStreamSynchronize(st);
    auto t0 = std::chrono::steady_clock::now();
    int e = enqueue_body(p, R, want_grad, nullptr, st, phase);
    auto t1 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(st);
    auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[plan trace] R=%d launches=%ld host enqueue %.1f us, drain %.1f us\n", R, p->launches,
            std::chrono::duration<double, std::micro>(t1 - t0).count(), std::chrono::duration<double, std::micro>(t2 - t1).count());
    if (e) return e;
  } else {
    PL_CHECK(enqueue_body(p, R, want_grad, nullptr, st, phase));
  }
  p->warmed[key] += 1;
  return 0;
}

int gpcsd_plan_enqueue(void* plan, int R, const double* h_theta, int want_grad, void* stream) {
  if (!plan) return fail_plan("null plan");
  Plan* p = (Plan*)plan;
  if (!p->ws) return fail_plan("plan: no LFP / workspace bound (gpcsd_plan_set_lfp)");
  if (R < 1 || R > p->Rmax) return fail_plan("plan: restart count out of range");
  cudaStream_t caller = (cudaStream_t)stream;
  const size_t nb = (size_t)R * (p->d.P + 2) * sizeof(double);
  stage_theta(p, R, h_theta);
  // the evaluation runs on the plan's own stream (the caller's may be the legacy default stream, which cannot be captured),
  // ordered after everything already enqueued on the caller's stream and before everything enqueued there afterwards
  GP_CUDA(cudaEventRecord(p->ev[6], caller));
  GP_CUDA(cudaStreamWaitEvent(p->own, p->ev[6], 0));
  PL_CHECK(enqueue_on_own(p, R, want_grad, nb));
  GP_CUDA(cudaEventRecord(p->ev[7], p->own));
  GP_CUDA(cudaStreamWaitEvent(caller, p->ev[7], 0));
  return 0;
}

// Read the (possibly all-reduced) device result back: h_out[R][P+4].  Blocks until the evaluation has finished.
int gpcsd_plan_finish(void* plan, int R, double* h_out, void* stream) {
  if (!plan) return fail_plan("null plan");
  Plan* p = (Plan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = (size_t)R * (p->d.P + 4) * sizeof(double);
  if (p->mb_host) {
    // all-reduce through the mailbox: wait for every rank's slot to carry this evaluation's sequence number, then sum the
    // vectors in RANK ORDER (deterministic, bit-identical on every rank).  Two parity buffers per slot: a rank can be at most
    // one evaluation ahead of the slowest one, because finishing evaluation k needs every rank's slot of evaluation k.
    const unsigned long long want = ++p->seq_host;
    const long n = (long)R * (p->d.P + 4);
    for (long i = 0; i < n; ++i) h_out[i] = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int q = 0; q < p->world; ++q) {
      const unsigned char* base = p->mb_host + (long)q * p->mb_slot_bytes + (long)(want & 1ull) * p->mb_parity_bytes;
      const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(base);
      unsigned long spins = 0;
      while (*flag != want) {
        if ((++spins & 0xFFFF) == 0) {
          const cudaError_t qe = cudaStreamQuery(st);
          if (qe != cudaSuccess && qe != cudaErrorNotReady) return gp_fail_cuda(qe, "waiting for the result mailbox", __LINE__);
          if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
            return fail_plan("plan: timed out waiting for a rank's result (ranks must issue the same evaluations)");
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      const volatile double* src = reinterpret_cast<const volatile double*>(base + 64);
      for (long i = 0; i < n; ++i) h_out[i] += src[i];
    }
    return 0;
  }
  GP_CUDA(cudaMemcpyAsync(p->h_out, p->out, nb, cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  memcpy(h_out, p->h_out, nb);
  return 0;
}

// Result mailbox of a trial-sharded model: ONE host shared-memory segment (POSIX shm, zero-initialised) mapped by every rank
// of the node.  gpcsd_plan_mailbox_bytes gives its size; gpcsd_plan_set_mailbox registers this rank's mapping with CUDA
// (host-mapped, so the evaluation's last kernel writes the result straight into it) and switches gpcsd_plan_finish to the
// mailbox all-reduce.  Every rank must then issue the same sequence of evaluations (as with any collective).
long gpcsd_plan_mailbox_bytes(void* plan, int world) {
  if (!plan || world < 1) return -1;
  Plan* p = (Plan*)plan;
  const long data = ((long)p->Rmax * (p->d.P + 4) * 8 + 63) / 64 * 64;
  const long slot = 2 * (64 + data);
  return ((long)world * slot + 4095) / 4096 * 4096;
}

int gpcsd_plan_set_mailbox(void* plan, void* h_shared, long bytes, int world, int rank) {
  if (!plan) return fail_plan("null plan");
  Plan* p = (Plan*)plan;
  if (world < 2 || rank < 0 || rank >= world) return fail_plan("plan_set_mailbox: bad world / rank");
  if (bytes < gpcsd_plan_mailbox_bytes(plan, world)) return fail_plan("plan_set_mailbox: segment too small");
  if (p->mb_host) return fail_plan("plan_set_mailbox: mailbox already set");
  GP_CUDA(cudaHostRegister(h_shared, (size_t)bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  void* dptr = nullptr;
  GP_CUDA(cudaHostGetDevicePointer(&dptr, h_shared, 0));
  GP_CUDA(cudaMalloc((void**)&p->seq_dev, sizeof(unsigned long long)));
  GP_CUDA(cudaMemset(p->seq_dev, 0, sizeof(unsigned long long)));
  const long data = ((long)p->Rmax * (p->d.P + 4) * 8 + 63) / 64 * 64;
  p->mb_parity_bytes = 64 + data;
  p->mb_slot_bytes = 2 * p->mb_parity_bytes;
  p->mb_host = (unsigned char*)h_shared;
  p->mb_dev = (unsigned char*)dptr;
  p->world = world; p->rank = rank; p->seq_host = 0;
  drop_graphs(p);
  return 0;
}

// One call per evaluation: h_out[r] = [loglik, d loglik / d theta (P), solver flag (0 = ok), checksum pair] for r < R.
// GEMM-phase token.  An evaluation is a latency-bound prologue on a few SMs (eigendecompositions, ~0.5 ms at configs[1])
// followed by kernels that fill the GPU (~1 ms).  Two models evaluated concurrently from two host threads should interleave
// as  A-GEMMs | B-GEMMs | A-GEMMs ...  with each prologue underneath the other model's GEMM phase; left to the hardware
// scheduler they drift into lock step instead (both GEMM phases share the SMs and finish together, then both prologues run
// together on an idle GPU: measured period = sum of everything).  So when calls overlap in time, the evaluation is issued
// as two launches and the second one -- the GEMM phase -- is taken under a process-wide mutex that is released as soon as
// the phase's last full-GPU kernel has finished (a host-mapped flag written from inside the graph).  Same kernels, same grids:
// results are bit-identical to the one-launch path.  Single-model processes never see an overlap and keep the one-launch
// path (the policy lapses four calls after the last overlap).  GPCSD_GEMM_TOKEN=0 / 1 forces the policy.
static std::mutex g_gemm_token;
static std::atomic<int> g_inflight{0};
static std::atomic<long> g_call_seq{0}, g_last_overlap{-(1L << 40)};

// GPCSD_GEMM_RESERVE=n (default 0): SMs the two-phase path leaves free when it sizes full-GPU grids.  With n = 16 a
// concurrent model's eigensolver clusters and small kernels find SMs at once instead of at the next kernel boundary of a
// persistent GEMM (+5 % throughput with two models in flight), but grid sizes -- hence the grouping of partial sums --
// then differ between overlapped and non-overlapped calls, so results are reproducible only to rounding (1e-15), not bit
// for bit, across concurrency patterns.  Off by default: n_workers = 2 and n_workers = 1 fits stay bit-identical.
static int gemm_reserve_sms() {
  static const int v = [] {
    const char* e = getenv("GPCSD_GEMM_RESERVE");
    const int n = e ? atoi(e) : 0;
    return n > 0 && n <= 64 ? n : 0;
  }();
  return v;
}

static int gemm_token_policy() {
  static const int v = [] {
    const char* e = getenv("GPCSD_GEMM_TOKEN");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  return v;
}

// GPCSD_TOKEN_STATS=1: mean host-side durations of the two-phase path, printed to stderr every 64 calls (developer aid)
struct TokenStats {
  std::atomic<long> n{0};
  std::atomic<long long> enq1{0}, wait_token{0}, enq2{0}, gemm{0}, finish{0};
};
static TokenStats g_tstats;
static bool token_stats_on() {
  static const bool v = getenv("GPCSD_TOKEN_STATS") != nullptr;
  return v;
}
static inline long long now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int loglik_grad_two_phase(Plan* p, int R, const double* h_theta, int want_grad, void* stream, long my_count) {
  if (!p->ws) return fail_plan("plan: no LFP / workspace bound (gpcsd_plan_set_lfp)");
  if (R < 1 || R > p->Rmax) return fail_plan("plan: restart count out of range");
  cudaStream_t caller = (cudaStream_t)stream;
  const size_t nb = (size_t)R * (p->d.P + 2) * sizeof(double);
  stage_theta(p, R, h_theta);
  GP_CUDA(cudaEventRecord(p->ev[6], caller));
  GP_CUDA(cudaStreamWaitEvent(p->own, p->ev[6], 0));
  const bool stats = token_stats_on();
  const long long t0s = stats ? now_ns() : 0;
  struct Prologue {                                   // "in flight, token not yet taken" (cleared on every exit path)
    explicit Prologue(Plan* p_) : p(p_) { set_token_state(p, 1); }
    ~Prologue() { set_token_state(p, 0); }
    Plan* p;
  } prologue_mark(p);
  gp_set_sm_reserve(gemm_reserve_sms());              // (see below; the prologue's one full-GPU pass, Z = Qs^T Yf, included)
  const int e1 = enqueue_on_own(p, R, want_grad, nb, 1);
  gp_set_sm_reserve(0);
  if (e1) return e1;
  // the token is only asked for once this model's own prologue is off the GPU: holding it while the GEMM phase still
  // waits (in stream order) for the eigendecompositions would keep the other model's GEMMs out for nothing
  GP_CUDA(cudaEventRecord(p->ev_p1, p->own));
  GP_CUDA(cudaEventSynchronize(p->ev_p1));
  const long long t1s = stats ? now_ns() : 0;
  {   // rank-consistent grant order (bounded: a prologue lasts ~1 ms; 20 ms means the other thread is not coming)
    const auto td = std::chrono::steady_clock::now();
    unsigned long spins = 0;
    while (must_defer(p, my_count)) {
      if ((++spins & 0x3FF) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - td).count() > 0.02) break;
    }
  }
  std::lock_guard<std::mutex> token(g_gemm_token);
  set_token_state(p, 0);
  const long long t2s = stats ? now_ns() : 0;
  // optionally the GEMM phase leaves SMs (16 = two 8-CTA clusters) to the other models' prologues: a persistent GEMM on all
  // SMs makes every kernel of a concurrent eigensolve wait for its next kernel boundary (measured: prologue 0.57 -> 0.84 ms)
  gp_set_sm_reserve(gemm_reserve_sms());
  const int e2 = enqueue_on_own(p, R, want_grad, nb, 2);
  gp_set_sm_reserve(0);
  if (e2) return e2;
  const long long t3s = stats ? now_ns() : 0;
  GP_CUDA(cudaEventRecord(p->ev[7], p->own));
  GP_CUDA(cudaStreamWaitEvent(caller, p->ev[7], 0));
  // wait until the GEMM phase is off the GPU: then the next model may start its own
  const unsigned long long want = ++p->gemm_seq_host;
  const volatile unsigned long long* flag = p->gemm_flag_host;
  const auto t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  while (*flag < want) {
    if ((++spins & 0xFFFF) == 0) {
      const cudaError_t qe = cudaStreamQuery(p->own);
      if (qe == cudaSuccess && *flag < want) return fail_plan("plan: the evaluation finished without marking its GEMM phase");
      if (qe != cudaSuccess && qe != cudaErrorNotReady) return gp_fail_cuda(qe, "waiting for the GEMM phase", __LINE__);
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
        return fail_plan("plan: timed out waiting for the GEMM phase");
    }
  }
  if (stats) {
    const long long t4s = now_ns();
    g_tstats.enq1 += t1s - t0s;
    g_tstats.wait_token += t2s - t1s;
    g_tstats.enq2 += t3s - t2s;
    g_tstats.gemm += t4s - t3s;
  }
  return 0;
}

int gpcsd_plan_loglik_grad(void* plan, int R, const double* h_theta, int want_grad, double* h_out, void* stream) {
  if (!plan) return fail_plan("null plan");
  const long seq = ++g_call_seq;
  struct Inflight {
    Inflight() { n = ++g_inflight; }
    ~Inflight() { --g_inflight; }
    int n;
  } guard;
  if (guard.n > 1) g_last_overlap.store(seq);
  const int policy = gemm_token_policy();
  // (not right after a new upload: that evaluation refolds the LFP and is issued kernel by kernel, not as a graph -- its
  //  ~50 launches are best queued in one go behind the upload it waits for anyway)
  const Plan* pp = (const Plan*)plan;
  const bool fresh_upload = (pp->s_split || pp->t_fold) && !pp->yf_valid && pp->N > 0;
  const bool two_phase = !fresh_upload && (policy == 1 || (policy < 0 && seq - g_last_overlap.load() < 4));
  const long my_count = next_eval_count((Plan*)plan);       // every call, whichever path: the same on every rank of an SPMD job
  if (two_phase) {
    PL_CHECK(loglik_grad_two_phase((Plan*)plan, R, h_theta, want_grad, stream, my_count));
  } else {
    PL_CHECK(gpcsd_plan_enqueue(plan, R, h_theta, want_grad, stream));
  }
  if (two_phase && token_stats_on()) {
    const long long t0f = now_ns();
    const int e = gpcsd_plan_finish(plan, R, h_out, stream);
    g_tstats.finish += now_ns() - t0f;
    const long n = ++g_tstats.n;
    if (n % 64 == 0) {
      fprintf(stderr, "[token stats] %ld calls: enqueue-1 %.1f us, wait for token %.1f us, enqueue-2 %.1f us, until GEMM phase done %.1f us, until result %.1f us\n",
              n, 1e-3 * g_tstats.enq1 / n, 1e-3 * g_tstats.wait_token / n, 1e-3 * g_tstats.enq2 / n, 1e-3 * g_tstats.gemm / n,
              1e-3 * g_tstats.finish / n);
    }
    return e;
  }
  return gpcsd_plan_finish(plan, R, h_out, stream);
}

// Kernel-level entry: the same evaluation for ONE hyperparameter vector with caller-supplied
// eigen-factors instead of the eigensolvers -- device arrays QsT [nx][even(nx)], ls [nx], QtT [nt][even(nt)], lt [nt]
// (rows = eigenvectors, as comp_eig_D's np.linalg.eigh columns, utility_functions.py:58-59).
int gpcsd_plan_loglik_grad_factors(void* plan, const double* h_theta, const double* QsT, const double* ls, const double* QtT,
                                   const double* lt, int want_grad, double* h_out, void* stream) {
  if (!plan) return fail_plan("null plan");
  Plan* p = (Plan*)plan;
  if (!p->ws) return fail_plan("plan: no LFP / workspace bound (gpcsd_plan_set_lfp)");
  cudaStream_t caller = (cudaStream_t)stream, st = p->own;
  const size_t nb = (size_t)(p->d.P + 2) * sizeof(double);
  stage_theta(p, 1, h_theta);
  GP_CUDA(cudaEventRecord(p->ev[6], caller));
  GP_CUDA(cudaStreamWaitEvent(st, p->ev[6], 0));
  GP_CUDA(cudaMemcpyAsync(p->theta, p->h_theta, nb, cudaMemcpyHostToDevice, st));
  Factors f{QsT, ls, QtT, lt};
  PL_CHECK(enqueue_body(p, 1, want_grad, &f, st));
  GP_CUDA(cudaEventRecord(p->ev[7], st));
  GP_CUDA(cudaStreamWaitEvent(caller, p->ev[7], 0));
  return gpcsd_plan_finish(plan, 1, h_out, caller);
}

}  // extern "C"
