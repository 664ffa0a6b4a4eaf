// Elementwise / reduction kernels of the GPCSD hot path, the cuSOLVER eigh wrapper and the ABI glue.
// HBM-bound or tiny by construction; see DESIGN.md section 4 for the per-kernel byte counts.
#include <cusolverDn.h>
#include <string.h>

#include <mutex>

#include <stdlib.h>

#include "common.h"
#include "dmma_gemm.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int gp_fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return 1;
}
int gp_fail_cuda(cudaError_t e, const char* what, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in `%s` at line %d", (int)e, cudaGetErrorString(e), what, line);
  return 2;
}
// SMs the calling thread leaves free when it sizes full-GPU grids (set around the GEMM phase of an evaluation while other
// models are in flight: their 8-CTA eigensolver clusters and small kernels then find SMs at once instead of at the next
// kernel boundary of a persistent GEMM)
static thread_local int tl_sm_reserve = 0;
void gp_set_sm_reserve(int n) { tl_sm_reserve = n > 0 ? n : 0; }

int gp_num_sms() {
  static int sms[GP_MAX_DEVICES];                  // per device (a process may drive more than one)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= GP_MAX_DEVICES) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    // GPCSD_SM_BUDGET=n: size every full-GPU grid (persistent GEMM, split-K SYRK) for n SMs, leaving the rest to concurrent
    // latency-bound kernels of ANOTHER model's evaluation (the 8-CTA eigensolver clusters): a persistent kernel that finds
    // 16 SMs occupied runs its last 16 CTAs as a second wave and takes almost twice as long
    if (const char* e = getenv("GPCSD_SM_BUDGET")) {
      const int b = atoi(e);
      if (b >= 8 && b < v) v = b;
    }
    sms[dev] = v;
  }
  const int v = sms[dev] - tl_sm_reserve;
  return v >= 8 ? v : 8;
}

namespace gpcsd {

// ------------------------------------------------------------------------------------------------
// forward-model weights
// ------------------------------------------------------------------------------------------------
__global__ void fwd_weights_1d_kernel(int npts, const double* __restrict__ x, int G, const double* __restrict__ gl_x,
                                      const double* __restrict__ gl_w, double R, double* __restrict__ A,
                                      double* __restrict__ dA, long ldg) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)npts * G) return;
  const int i = (int)(idx / G), gq = (int)(idx % G);
  const double d = (gl_x[gq] - x[i]) / R;
  const double qd = d * d;
  const double s1 = sqrt(qd + 1.0), s0 = sqrt(qd);
  const double w = gl_w[gq];
  A[(long)i * ldg + gq] = w * (s1 - s0);
  // d/dR [sqrt(d^2+1) - |d|], d = r/R  ->  (d^2/sqrt(d^2+1) - |d|) * (-1/R)
  if (dA) dA[(long)i * ldg + gq] = w * (qd / s1 - s0) * (-1.0 / R);
}

__global__ void fwd_weights_2d_kernel(int npts, const double* __restrict__ pts, int ngl1, int ngl2,
                                      const double* __restrict__ gl_x1, const double* __restrict__ gl_w1,
                                      const double* __restrict__ gl_x2, const double* __restrict__ gl_w2, double R,
                                      double eps, double* __restrict__ A, double* __restrict__ dA, long ldg) {
  const int G = ngl1 * ngl2;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)npts * G) return;
  const int i = (int)(idx / G), gq = (int)(idx % G);
  const int g1 = gq / ngl2, g2 = gq % ngl2;
  const double d1 = gl_x1[g1] - pts[2 * i], d2 = gl_x2[g2] - pts[2 * i + 1];
  const double w = sqrt(d1 * d1 + d2 * d2);  // delta_w, covariances.py:131
  const double w2 = w * w;
  const double Re = R + eps;
  const double sR = sqrt(Re * Re + w2);
  const double wp = gl_w1[g1] * gl_w2[g2];
  A[(long)i * ldg + gq] = wp * (log(Re + sR) - log(eps + sqrt(eps * eps + w2)));
  if (dA) dA[(long)i * ldg + gq] = wp / sR;  // d/dR log(Re + sqrt(Re^2+w^2)) = 1/sqrt(Re^2+w^2)
}

// ------------------------------------------------------------------------------------------------
// SE factor matrices
// ------------------------------------------------------------------------------------------------
__global__ void se_matrix_kernel(int na, const double* __restrict__ a, int nb, const double* __restrict__ b, double ell,
                                 double scale, int deriv, double* __restrict__ out, long ld) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)na * nb) return;
  const int i = (int)(idx / nb), j = (int)(idx % nb);
  const double d = a[i] - b[j];
  const double u = d / ell;
  double v = scale * exp(-0.5 * u * u);
  if (deriv) v *= d * d / (ell * ell * ell);
  out[(long)i * ld + j] = v;
}

__global__ void se_grid_to_pts_kernel(int ngl1, int ngl2, const double* __restrict__ gl_x1,
                                      const double* __restrict__ gl_x2, int nz, const double* __restrict__ z,
                                      double ell1, double ell2, double* __restrict__ out, long ld) {
  const long G = (long)ngl1 * ngl2;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * nz) return;
  const int k = (int)(idx / G), gq = (int)(idx % G);
  const double u1 = (gl_x1[gq / ngl2] - z[2 * k]) / ell1, u2 = (gl_x2[gq % ngl2] - z[2 * k + 1]) / ell2;
  out[(long)k * ld + gq] = exp(-0.5 * u1 * u1) * exp(-0.5 * u2 * u2);
}

// ------------------------------------------------------------------------------------------------
// temporal covariance and its gradient contraction
// ------------------------------------------------------------------------------------------------
struct TemporalSpec {
  int ntc;
  int kind[8];
  double ell[8];
  double sigma2[8];
};

__device__ __forceinline__ double kt_term(int kind, double ell, double d) {
  return kind == GPCSD_KIND_SE ? exp(-0.5 * d * d / (ell * ell)) : exp(-fabs(d) / ell);
}

__global__ void kt_build_kernel(int nr, const double* __restrict__ t, int nc, const double* __restrict__ tp,
                                TemporalSpec sp, double* __restrict__ Kt, long ld) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nr * nc) return;
  const int i = (int)(idx / nc), j = (int)(idx % nc);
  const double d = t[i] - tp[j];
  double v = 0.0;
  for (int k = 0; k < sp.ntc; ++k) v += sp.sigma2[k] * kt_term(sp.kind[k], sp.ell[k], d);
  Kt[(long)i * ld + j] = v;
}

// per-block partials ws[block][2*ntc]; second pass sums blocks in order
__global__ void kt_grad_kernel(int nt, const double* __restrict__ t, TemporalSpec sp, const double* __restrict__ G,
                               long ldg, double* __restrict__ ws) {
  __shared__ double red[8];
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  const long total = (long)nt * nt;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / nt), j = (int)(idx % nt);
    const double d = t[i] - t[j];
    const double gv = G[(long)i * ldg + j];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < sp.ntc) {
        const double e = kt_term(sp.kind[k], sp.ell[k], d);
        const double ell = sp.ell[k];
        const double dl = sp.kind[k] == GPCSD_KIND_SE ? d * d / (ell * ell * ell) : fabs(d) / (ell * ell);
        acc[2 * k] += gv * sp.sigma2[k] * e * dl;  // <G, dKt_k/d ell>
        acc[2 * k + 1] += gv * e;                  // <G, dKt_k/d sigma2>
      }
    }
  }
  for (int k = 0; k < 2 * sp.ntc; ++k) {
    const double s = block_sum(acc[k], red);
    if (threadIdx.x == 0) ws[(long)blockIdx.x * 16 + k] = s;
  }
}

__global__ void sum_rows_kernel(const double* __restrict__ ws, int nrows, int stride, int ncols, double* __restrict__ out) {
  const int k = blockIdx.x;
  __shared__ double red[8];
  double a = 0.0;
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) a += ws[(long)r * stride + k];
  const double s = block_sum(a, red);
  if (threadIdx.x == 0 && k < ncols) out[k] = s;
}

// ------------------------------------------------------------------------------------------------
// D, 1/D and their reductions.  One CTA per spatial eigen-index i for the row sums; column sums and
// the two scalars by a second tiny kernel over the stored rows (deterministic order).
// ------------------------------------------------------------------------------------------------
__global__ void eig_D_rows_kernel(int nx, int nt, const double* __restrict__ ls, const double* __restrict__ lt,
                                  const double* __restrict__ sig2n, int n_sig2n, double* __restrict__ rD, long ldrd,
                                  double* __restrict__ rowA, double* __restrict__ rowC, double* __restrict__ rowL) {
  __shared__ double red[8];
  const int i = blockIdx.x;
  const double l = ls[i], s = (n_sig2n == 1) ? sig2n[0] : sig2n[i];
  double a = 0.0, c = 0.0, lg = 0.0;
  for (int j = threadIdx.x; j < nt; j += blockDim.x) {
    const double D = l * lt[j] + s;
    const double r = 1.0 / D;
    rD[(long)i * ldrd + j] = r;
    a += lt[j] * r;
    c += r;
    lg += log(D);
  }
  const double sa = block_sum(a, red);
  const double sc = block_sum(c, red);
  const double sl = block_sum(lg, red);
  if (threadIdx.x == 0) {
    rowA[i] = sa;
    rowC[i] = sc;
    rowL[i] = sl;
  }
}

__global__ void eig_D_cols_kernel(int nx, int nt, const double* __restrict__ ls, const double* __restrict__ rD,
                                  long ldrd, const double* __restrict__ rowC, const double* __restrict__ rowL,
                                  double* __restrict__ colB, double* __restrict__ sums2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nt) {
    double b = 0.0;
    for (int i = 0; i < nx; ++i) b += ls[i] * rD[(long)i * ldrd + j];
    colB[j] = b;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double sl = 0.0, sc = 0.0;
    for (int i = 0; i < nx; ++i) {
      sl += rowL[i];
      sc += rowC[i];
    }
    sums2[0] = sl;
    sums2[1] = sc;
  }
}

// ------------------------------------------------------------------------------------------------
// eigen-basis gradient core
// ------------------------------------------------------------------------------------------------
__global__ void grad_core_kernel(int n, const double* __restrict__ Mm, long ldm, const double* __restrict__ Nm,
                                 long ldnm, const double* __restrict__ lam, const double* __restrict__ s,
                                 const double* __restrict__ rowsum, double ntot, double scale_data, double scale_det,
                                 double* __restrict__ X, long ldx) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  double v = 0.5 * scale_data * Mm[(long)i * ldm + j];
  if (i == j) {
    v += -0.5 * ntot * scale_det * rowsum[i];
  } else if (Nm) {
    const double dl = lam[i] - lam[j];
    if (dl != 0.0) v += 0.5 * scale_data * ((s[i] - s[j]) / dl) * Nm[(long)i * ldnm + j];
  }
  X[(long)i * ldx + j] = v;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void add_diag_kernel(int n, double* K, long ld, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) K[(long)i * ld + i] += v;
}

__global__ void transpose_kernel(int rows, int cols, const double* __restrict__ in, long ldi, double* __restrict__ out,
                                 long ldo) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int rr = r0 + r, cc = c0 + threadIdx.x;
    if (rr < rows && cc < cols) tile[r][threadIdx.x] = in[(long)rr * ldi + cc];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int oc = r0 + threadIdx.x, orow = c0 + r;  // out[orow][oc] = in[oc][orow]
    if (orow < cols && oc < rows) out[(long)orow * ldo + oc] = tile[threadIdx.x][r];
  }
}

__global__ void dot_kernel(int rows, int cols, const double* __restrict__ X, long ldx, const double* __restrict__ Y,
                           long ldy, double* __restrict__ ws) {
  __shared__ double red[8];
  double a = 0.0;
  const long total = (long)rows * cols;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / cols, c = idx % cols;
    a += X[r * ldx + c] * Y[r * ldy + c];
  }
  const double s = block_sum(a, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}

struct PtrList {
  const double* p[8];
};
__global__ void sum_arrays_kernel(long n, int narr, PtrList in, double* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double v = in.p[0][i];
    for (int k = 1; k < narr; ++k) v += in.p[k][i];
    out[i] = v;
  }
}

static int dot_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = 4L * gp_num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------
// cuSOLVER
// ------------------------------------------------------------------------------------------------
// One cuSOLVER handle per CUDA stream (process-wide table): independent eigenproblems run concurrently -- the two
// factors of one model on side streams, independent models on separate host threads/streams -- and a handle
// (with the cuBLAS workspace inside it) must not be shared by work in flight on two streams.  Creating a handle
// costs ~100 ms, so they are cached for the life of the process.
struct SolverSlot {
  int device;
  cudaStream_t stream;
  cusolverDnHandle_t handle;
};
static SolverSlot g_slots[64];
static int g_nslots = 0;
static std::mutex g_slot_mutex;
static int solver_handle(cusolverDnHandle_t* h, cudaStream_t st = nullptr) {
  std::lock_guard<std::mutex> lock(g_slot_mutex);
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < g_nslots; ++i)
    if (g_slots[i].stream == st && g_slots[i].device == dev) {
      *h = g_slots[i].handle;
      return 0;
    }
  // a full table is an error rather than a reason to destroy a handle another thread may still be using
  if (g_nslots == 64) return gp_fail("cuSOLVER handle table full (64 (device, stream) pairs)");
  cusolverDnHandle_t nh = nullptr;
  if (cusolverDnCreate(&nh) != CUSOLVER_STATUS_SUCCESS) return gp_fail("cusolverDnCreate failed");
  if (cusolverDnSetStream(nh, st) != CUSOLVER_STATUS_SUCCESS) return gp_fail("cusolverDnSetStream failed");
  g_slots[g_nslots].device = dev;
  g_slots[g_nslots].stream = st;
  g_slots[g_nslots].handle = nh;
  ++g_nslots;
  *h = nh;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// centrosymmetric split (Cantoni & Butler 1976).  A symmetric Toeplitz matrix -- any stationary temporal
// kernel on a uniform time grid -- satisfies J K J = K, so its eigenvectors are symmetric ([u; Ju]/sqrt2, or
// [u; sqrt2 a; Ju]/sqrt2 for odd n) or skew-symmetric ([v; -Jv]/sqrt2):  the n x n eigenproblem splits
// EXACTLY into two independent half-size ones, S = K11 + K12 J (bordered for odd n) and A = K11 - K12 J.
// ------------------------------------------------------------------------------------------------
__global__ void centro_split_kernel(int n, const double* __restrict__ K, long ldk, double* __restrict__ S, long lds,
                                    double* __restrict__ A, long lda) {
  const int m = n / 2, odd = n & 1, ms = m + odd;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)ms * ms) return;
  const int i = (int)(idx / ms), j = (int)(idx % ms);
  const double r2 = 1.4142135623730951;
  if (i < m && j < m) {
    const double a = K[(long)i * ldk + j], b = K[(long)i * ldk + (n - 1 - j)];
    S[(long)i * lds + j] = a + b;
    A[(long)i * lda + j] = a - b;
  } else if (i == m && j == m) {
    S[(long)i * lds + j] = K[(long)m * ldk + m];
  } else {  // odd border
    const int r = (i == m) ? j : i;
    S[(long)i * lds + j] = r2 * K[(long)r * ldk + m];
  }
}

// QT rows = eigenvectors of K: first ms rows from the symmetric block (UsT, ms x ms), then m rows from the skew
// block (UaT, m x m); W = [Ws, Wa] (not sorted: the Kronecker sums do not depend on the order).
__global__ void centro_assemble_kernel(int n, const double* __restrict__ UsT, long lds, const double* __restrict__ Ws,
                                       const double* __restrict__ UaT, long lda, const double* __restrict__ Wa,
                                       double* __restrict__ QT, long ldq, double* __restrict__ W) {
  const int m = n / 2, odd = n & 1, ms = m + odd;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int a = (int)(idx / n), c = (int)(idx % n);
  const double h = 0.70710678118654752;
  double v;
  if (a < ms) {
    if (c < m) v = h * UsT[(long)a * lds + c];
    else if (odd && c == m) v = UsT[(long)a * lds + m];
    else v = h * UsT[(long)a * lds + (n - 1 - c)];
  } else {
    const int b = a - ms;
    if (c < m) v = h * UaT[(long)b * lda + c];
    else if (odd && c == m) v = 0.0;
    else v = -h * UaT[(long)b * lda + (n - 1 - c)];
  }
  QT[(long)a * ldq + c] = v;
  if (c == 0) W[a] = (a < ms) ? Ws[a] : Wa[a - ms];
}

// Centrosymmetric fold of the time axis of a trial block X[nblk][n][rowlen] (rowlen even, trials contiguous):
//   Xf[b][j]      = (X[b][j] + X[b][n-1-j]) / sqrt 2   j < n/2        (for odd n: Xf[b][n/2] = X[b][n/2])
//   Xf[b][ms + j] = (X[b][j] - X[b][n-1-j]) / sqrt 2   j < n/2,  ms = n/2 + n%2
// With the eigenvectors of a centrosymmetric Kt stored as [symmetric block; skew block] (centro_assemble_kernel) this turns
// Qt^T X_b into two independent products of order ~n/2 (Us^T Xf[b][:ms], Ua^T Xf[b][ms:]): half the flops of the projection.
__global__ void centro_fold_kernel(int n, long rowlen2, long total, const double2* __restrict__ X, double2* __restrict__ Xf) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;        // over [nblk][ms][rowlen/2]
  if (idx >= total) return;
  const int m = n / 2, ms = m + (n & 1);
  const long c = idx % rowlen2, bj = idx / rowlen2;
  const int j = (int)(bj % ms);
  const long b = bj / ms;
  const double2* xb = X + b * n * rowlen2;
  double2* fb = Xf + b * n * rowlen2;
  const double h = 0.70710678118654752;
  if (j < m) {
    const double2 u = xb[(long)j * rowlen2 + c], v = xb[(long)(n - 1 - j) * rowlen2 + c];
    fb[(long)j * rowlen2 + c] = make_double2(h * (u.x + v.x), h * (u.y + v.y));
    fb[(long)(ms + j) * rowlen2 + c] = make_double2(h * (u.x - v.x), h * (u.y - v.y));
  } else {
    fb[(long)j * rowlen2 + c] = xb[(long)j * rowlen2 + c];
  }
}

// Inverse of centro_fold_kernel (the fold matrix is orthogonal): X[b][j] = (Xf[b][j] + Xf[b][ms+j]) / sqrt 2,
// X[b][n-1-j] = (Xf[b][j] - Xf[b][ms+j]) / sqrt 2 for j < n/2, middle row copied for odd n.
__global__ void centro_unfold_kernel(int n, long rowlen2, long total, const double2* __restrict__ Xf, double2* __restrict__ X) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;        // over [nblk][ms][rowlen/2]
  if (idx >= total) return;
  const int m = n / 2, ms = m + (n & 1);
  const long c = idx % rowlen2, bj = idx / rowlen2;
  const int j = (int)(bj % ms);
  const long b = bj / ms;
  const double2* fb = Xf + b * n * rowlen2;
  double2* xb = X + b * n * rowlen2;
  const double h = 0.70710678118654752;
  if (j < m) {
    const double2 u = fb[(long)j * rowlen2 + c], v = fb[(long)(ms + j) * rowlen2 + c];
    xb[(long)j * rowlen2 + c] = make_double2(h * (u.x + v.x), h * (u.y + v.y));
    xb[(long)(n - 1 - j) * rowlen2 + c] = make_double2(h * (u.x - v.x), h * (u.y - v.y));
  } else {
    xb[(long)j * rowlen2 + c] = fb[(long)j * rowlen2 + c];
  }
}

__global__ void copy_matrix_kernel(int n, const double* __restrict__ in, long ldi, double* __restrict__ out, long ldo) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  out[(long)i * ldo + j] = in[(long)i * ldi + j];
}

// Generalisation to any fixed-point-free involution pi of the index set with K[pi(i)][pi(j)] == K[i][j] (e.g. the point
// reflection of a Neuropixels checkerboard about the centre of the integration box): with representatives ra[] and
// partners rb[] = pi(ra[]),  S = K[ra,ra] + K[ra,rb],  A = K[ra,ra] - K[ra,rb]  (order n/2 each).
__global__ void pairsym_split_kernel(int m, const double* __restrict__ K, long ldk, const int* __restrict__ ra,
                                     const int* __restrict__ rb, double* __restrict__ S, long lds, double* __restrict__ A,
                                     long lda) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)m * m) return;
  const int i = (int)(idx / m), j = (int)(idx % m);
  const double a = K[(long)ra[i] * ldk + ra[j]], b = K[(long)ra[i] * ldk + rb[j]];
  S[(long)i * lds + j] = a + b;
  A[(long)i * lda + j] = a - b;
}

__global__ void pairsym_assemble_kernel(int m, const int* __restrict__ ra, const int* __restrict__ rb,
                                        const double* __restrict__ UsT, long lds, const double* __restrict__ Ws,
                                        const double* __restrict__ UaT, long lda, const double* __restrict__ Wa,
                                        double* __restrict__ QT, long ldq, double* __restrict__ W) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2L * m * m) return;
  const int a = (int)(idx / m), c = (int)(idx % m);   // a: eigenvector index 0..2m-1, c: representative index
  const double h = 0.70710678118654752;
  if (a < m) {
    const double v = h * UsT[(long)a * lds + c];
    QT[(long)a * ldq + ra[c]] = v;
    QT[(long)a * ldq + rb[c]] = v;
  } else {
    const double v = h * UaT[(long)(a - m) * lda + c];
    QT[(long)a * ldq + ra[c]] = v;
    QT[(long)a * ldq + rb[c]] = -v;
  }
  if (c == 0) W[a] = (a < m) ? Ws[a] : Wa[a - m];
}

// Fold of the channel axis under the involution pi (see pairsym_split_kernel): X[n][rowlen] -> Xf[n][rowlen],
//   Xf[c] = (X[ra[c]] + X[rb[c]]) / sqrt 2,   Xf[m + c] = (X[ra[c]] - X[rb[c]]) / sqrt 2,   c < m = n/2.
// In this basis the eigenvector matrix assembled by pairsym_assemble_kernel is block diagonal (Us^T, Ua^T): Qs^T Y becomes two
// products of order n/2, and only the two diagonal blocks of the spatial SYRK enter the gradient.
__global__ void pairsym_fold_kernel(int m, long rowlen2, const int* __restrict__ ra, const int* __restrict__ rb,
                                    const double2* __restrict__ X, double2* __restrict__ Xf) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;      // over [m][rowlen/2]
  if (idx >= (long)m * rowlen2) return;
  const long c = idx / rowlen2, col = idx - c * rowlen2;
  const double h = 0.70710678118654752;
  const double2 u = X[(long)ra[c] * rowlen2 + col], v = X[(long)rb[c] * rowlen2 + col];
  Xf[c * rowlen2 + col] = make_double2(h * (u.x + v.x), h * (u.y + v.y));
  Xf[(m + c) * rowlen2 + col] = make_double2(h * (u.x - v.x), h * (u.y - v.y));
}

}  // namespace gpcsd

using namespace gpcsd;

#define GRID1D(n) (unsigned)(((long)(n) + 255) / 256), 256

extern "C" {

int gpcsd_abi_version(void) { return GPCSD_B200_ABI_VERSION; }
const char* gpcsd_last_error(void) { return g_err; }
int gpcsd_num_sms(void) { return gp_num_sms(); }

int gpcsd_fwd_weights_1d(int npts, const double* x, int G, const double* gl_x, const double* gl_w, double R, double* A,
                         double* dA, long ldg, void* stream) {
  if (npts <= 0 || G <= 0) return 0;
  fwd_weights_1d_kernel<<<GRID1D((long)npts * G), 0, (cudaStream_t)stream>>>(npts, x, G, gl_x, gl_w, R, A, dA, ldg);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_fwd_weights_2d(int npts, const double* pts, int ngl1, int ngl2, const double* gl_x1, const double* gl_w1,
                         const double* gl_x2, const double* gl_w2, double R, double eps, double* A, double* dA, long ldg,
                         void* stream) {
  if (npts <= 0 || ngl1 <= 0 || ngl2 <= 0) return 0;
  fwd_weights_2d_kernel<<<GRID1D((long)npts * ngl1 * ngl2), 0, (cudaStream_t)stream>>>(npts, pts, ngl1, ngl2, gl_x1, gl_w1,
                                                                                   gl_x2, gl_w2, R, eps, A, dA, ldg);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_se_matrix(int na, const double* a, int nb, const double* b, double ell, double scale, int deriv, double* out,
                    long ld, void* stream) {
  if (na <= 0 || nb <= 0) return 0;
  se_matrix_kernel<<<GRID1D((long)na * nb), 0, (cudaStream_t)stream>>>(na, a, nb, b, ell, scale, deriv, out, ld);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_se_grid_to_pts(int ngl1, int ngl2, const double* gl_x1, const double* gl_x2, int nz, const double* z,
                         double ell1, double ell2, double* out, long ld, void* stream) {
  if (nz <= 0) return 0;
  se_grid_to_pts_kernel<<<GRID1D((long)ngl1 * ngl2 * nz), 0, (cudaStream_t)stream>>>(ngl1, ngl2, gl_x1, gl_x2, nz, z, ell1,
                                                                                 ell2, out, ld);
  GP_CUDA(cudaGetLastError());
  return 0;
}

static int make_spec(int ntc, const int* h_kind, const double* h_ell, const double* h_sigma2, TemporalSpec* sp) {
  if (ntc < 1 || ntc > 8) return gp_fail("temporal covariance list must have 1..8 entries");
  sp->ntc = ntc;
  for (int k = 0; k < ntc; ++k) {
    if (h_kind[k] != GPCSD_KIND_SE && h_kind[k] != GPCSD_KIND_MATERN) return gp_fail("unknown temporal kernel kind");
    sp->kind[k] = h_kind[k];
    sp->ell[k] = h_ell[k];
    sp->sigma2[k] = h_sigma2[k];
  }
  return 0;
}

int gpcsd_kt_build(int nt_rows, const double* t, int nt_cols, const double* tp, int ntc, const int* h_kind,
                   const double* h_ell, const double* h_sigma2, double* Kt, long ld, void* stream) {
  TemporalSpec sp;
  if (int e = make_spec(ntc, h_kind, h_ell, h_sigma2, &sp)) return e;
  if (nt_rows <= 0 || nt_cols <= 0) return 0;
  kt_build_kernel<<<GRID1D((long)nt_rows * nt_cols), 0, (cudaStream_t)stream>>>(nt_rows, t, nt_cols, tp, sp, Kt, ld);
  GP_CUDA(cudaGetLastError());
  return 0;
}

long gpcsd_kt_grad_ws_doubles(int nt, int ntc) {
  (void)ntc;
  return 16L * dot_blocks((long)nt * nt);
}

int gpcsd_kt_grad(int nt, const double* t, int ntc, const int* h_kind, const double* h_ell, const double* h_sigma2,
                  const double* G, long ldg, double* ws, double* out, void* stream) {
  TemporalSpec sp;
  if (int e = make_spec(ntc, h_kind, h_ell, h_sigma2, &sp)) return e;
  const int nb = dot_blocks((long)nt * nt);
  kt_grad_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(nt, t, sp, G, ldg, ws);
  GP_CUDA(cudaGetLastError());
  sum_rows_kernel<<<2 * ntc, 256, 0, (cudaStream_t)stream>>>(ws, nb, 16, 2 * ntc, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_eig_D(int nx, int nt, const double* ls, const double* lt, const double* sig2n, int n_sig2n, double* rD,
                long ldrd, double* sums2, double* rowA, double* rowC, double* rowL, double* colB, void* stream) {
  if (n_sig2n != 1 && n_sig2n != nx) return gp_fail("sig2n must be a scalar or have one entry per electrode");
  cudaStream_t st = (cudaStream_t)stream;
  eig_D_rows_kernel<<<nx, 256, 0, st>>>(nx, nt, ls, lt, sig2n, n_sig2n, rD, ldrd, rowA, rowC, rowL);
  GP_CUDA(cudaGetLastError());
  eig_D_cols_kernel<<<(nt + 255) / 256, 256, 0, st>>>(nx, nt, ls, rD, ldrd, rowC, rowL, colB, sums2);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_grad_core(int n, const double* Mmat, long ldm, const double* Nmat, long ldnm, const double* lam,
                    const double* s, const double* rowsum, double ntrials_total, double scale_data, double scale_det,
                    double* X, long ldx, void* stream) {
  grad_core_kernel<<<GRID1D((long)n * n), 0, (cudaStream_t)stream>>>(n, Mmat, ldm, Nmat, ldnm, lam, s, rowsum,
                                                                 ntrials_total, scale_data, scale_det, X, ldx);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_add_diag(int n, double* K, long ld, double v, void* stream) {
  add_diag_kernel<<<GRID1D(n), 0, (cudaStream_t)stream>>>(n, K, ld, v);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_transpose(int rows, int cols, const double* in, long ldi, double* out, long ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(rows, cols, in, ldi, out, ldo);
  GP_CUDA(cudaGetLastError());
  return 0;
}

long gpcsd_dot_ws_doubles(long n) { return dot_blocks(n); }

int gpcsd_dot(int rows, int cols, const double* X, long ldx, const double* Y, long ldy, double* ws, double* out,
              void* stream) {
  const int nb = dot_blocks((long)rows * cols);
  dot_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(rows, cols, X, ldx, Y, ldy, ws);
  GP_CUDA(cudaGetLastError());
  sum_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ws, nb, 1, 1, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_sum_arrays(long n, int narr, const double* const* h_in, double* out, void* stream) {
  if (narr < 1 || narr > 8) return gp_fail("sum_arrays: 1..8 inputs");
  PtrList pl;
  for (int k = 0; k < narr; ++k) pl.p[k] = h_in[k];
  long nb = (n + 255) / 256;
  const long cap = 16L * gp_num_sms();
  if (nb > cap) nb = cap;
  sum_arrays_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(n, narr, pl, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_sum_vec(long n, const double* in, double* out, void* stream) {
  sum_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(in, (int)n, 1, 1, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

// ---- batched small-order path --------------------------------------------------------------------------------
// cusolverDnXsyevBatched keeps every matrix of order <= 128 inside one CTA (measured on B200: two 128x128 problems in
// 0.78 ms, 64 problems of order 50 in 0.39 ms) while a single syevd of order 128 costs 2.1 ms of launch/sync latency;
// above 128 (or with batch == 1) it is no faster than syevd.  Used for the two halves of the split temporal factor when
// nt <= 256 and for spatial factors of order <= 128 (batch 2 with the matrix duplicated).
static cusolverDnParams_t g_params = nullptr;
static int solver_params(cusolverDnParams_t* p) {
  std::lock_guard<std::mutex> lock(g_slot_mutex);
  if (!g_params && cusolverDnCreateParams(&g_params) != CUSOLVER_STATUS_SUCCESS) return gp_fail("cusolverDnCreateParams failed");
  *p = g_params;
  return 0;
}

long gpcsd_eigh_batched_ws_bytes(int n, long ld, int batch) {
  cusolverDnHandle_t h;
  cusolverDnParams_t prm;
  if (solver_handle(&h) || solver_params(&prm)) return -1;
  size_t wd = 0, wh = 0;
  cusolverStatus_t s = cusolverDnXsyevBatched_bufferSize(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F,
                                                         nullptr, ld, CUDA_R_64F, nullptr, CUDA_R_64F, &wd, &wh, batch);
  if (s != CUSOLVER_STATUS_SUCCESS || wh != 0) {
    gp_fail("cusolverDnXsyevBatched_bufferSize failed (or asks for host workspace)");
    return -1;
  }
  return (long)wd;
}

int gpcsd_eigh_batched(int n, int batch, double* A, long ld, double* W, void* ws, long ws_bytes, int* info, void* stream) {
  cusolverDnHandle_t h;
  cusolverDnParams_t prm;
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = solver_handle(&h, st)) return e;
  if (int e = solver_params(&prm)) return e;
  // in: `batch` symmetric matrices [n][ld] stacked with stride n*ld; out: eigenvectors as ROWS (column-major Q == row-major Q^T)
  cusolverStatus_t s = cusolverDnXsyevBatched(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, ld,
                                              CUDA_R_64F, W, CUDA_R_64F, ws, (size_t)ws_bytes, nullptr, 0, info, batch);
  if (s != CUSOLVER_STATUS_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cusolverDnXsyevBatched failed with status %d", (int)s);
    return 3;
  }
  return 0;
}

long gpcsd_eigh_ws_doubles(int n, long ldq) {
  cusolverDnHandle_t h;
  if (solver_handle(&h)) return -1;
  int lwork = 0;
  cusolverStatus_t s = cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, nullptr,
                                                   (int)ldq, nullptr, &lwork);
  if (s != CUSOLVER_STATUS_SUCCESS) {
    gp_fail("cusolverDnDsyevd_bufferSize failed");
    return -1;
  }
  return (long)lwork;
}

int gpcsd_centro_split(int n, const double* K, long ldk, double* S, long lds, double* A, long lda, void* stream) {
  if (n < 2) return gp_fail("centro_split: n must be >= 2");
  const long ms = n / 2 + (n & 1);
  centro_split_kernel<<<GRID1D(ms * ms), 0, (cudaStream_t)stream>>>(n, K, ldk, S, lds, A, lda);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_centro_assemble(int n, const double* UsT, long lds, const double* Ws, const double* UaT, long lda,
                          const double* Wa, double* QT, long ldq, double* W, void* stream) {
  centro_assemble_kernel<<<GRID1D((long)n * n), 0, (cudaStream_t)stream>>>(n, UsT, lds, Ws, UaT, lda, Wa, QT, ldq, W);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_centro_fold(int nblk, int n, long rowlen, const double* X, double* Xf, void* stream) {
  if (rowlen & 1L) return gp_fail("centro_fold: row length must be even");
  if (((uintptr_t)X | (uintptr_t)Xf) & 15) return gp_fail("centro_fold: pointers must be 16-byte aligned");
  if (nblk <= 0 || n <= 0 || rowlen <= 0) return 0;
  const long ms = n / 2 + (n & 1), total = (long)nblk * ms * (rowlen / 2);
  centro_fold_kernel<<<GRID1D(total), 0, (cudaStream_t)stream>>>(n, rowlen / 2, total, (const double2*)X, (double2*)Xf);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_centro_unfold(int nblk, int n, long rowlen, const double* Xf, double* X, void* stream) {
  if (rowlen & 1L) return gp_fail("centro_unfold: row length must be even");
  if (((uintptr_t)X | (uintptr_t)Xf) & 15) return gp_fail("centro_unfold: pointers must be 16-byte aligned");
  if (nblk <= 0 || n <= 0 || rowlen <= 0) return 0;
  const long ms = n / 2 + (n & 1), total = (long)nblk * ms * (rowlen / 2);
  centro_unfold_kernel<<<GRID1D(total), 0, (cudaStream_t)stream>>>(n, rowlen / 2, total, (const double2*)Xf, (double2*)X);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_pairsym_split(int n, const double* K, long ldk, const int* ra, const int* rb, double* S, long lds, double* A,
                        long lda, void* stream) {
  if (n < 2 || (n & 1)) return gp_fail("pairsym_split: n must be even and >= 2");
  const long m = n / 2;
  pairsym_split_kernel<<<GRID1D(m * m), 0, (cudaStream_t)stream>>>((int)m, K, ldk, ra, rb, S, lds, A, lda);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_pairsym_assemble(int n, const int* ra, const int* rb, const double* UsT, long lds, const double* Ws,
                           const double* UaT, long lda, const double* Wa, double* QT, long ldq, double* W, void* stream) {
  if (n < 2 || (n & 1)) return gp_fail("pairsym_assemble: n must be even and >= 2");
  const long m = n / 2;
  pairsym_assemble_kernel<<<GRID1D(2 * m * m), 0, (cudaStream_t)stream>>>((int)m, ra, rb, UsT, lds, Ws, UaT, lda, Wa, QT, ldq, W);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_pairsym_fold(int n, const int* ra, const int* rb, long rowlen, const double* X, double* Xf, void* stream) {
  if (n < 2 || (n & 1)) return gp_fail("pairsym_fold: n must be even and >= 2");
  if (rowlen & 1L) return gp_fail("pairsym_fold: row length must be even");
  if (((uintptr_t)X | (uintptr_t)Xf) & 15) return gp_fail("pairsym_fold: pointers must be 16-byte aligned");
  const long m = n / 2;
  pairsym_fold_kernel<<<GRID1D(m * (rowlen / 2)), 0, (cudaStream_t)stream>>>((int)m, rowlen / 2, ra, rb, (const double2*)X,
                                                                             (double2*)Xf);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_eigh(int n, const double* K, long ldk, double* QT, long ldq, double* W, double* ws, long ws_doubles, int* info,
               void* stream) {
  cusolverDnHandle_t h;
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = solver_handle(&h, st)) return e;
  copy_matrix_kernel<<<GRID1D((long)n * n), 0, st>>>(n, K, ldk, QT, ldq);
  GP_CUDA(cudaGetLastError());
  // K symmetric: row-major == column-major on input; output eigenvectors are the COLUMNS of a column-major
  // matrix, i.e. the ROWS of QT viewed row-major (QT = Q^T).
  cusolverStatus_t s = cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, QT, (int)ldq, W, ws,
                                        (int)ws_doubles, info);
  if (s != CUSOLVER_STATUS_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cusolverDnDsyevd failed with status %d", (int)s);
    return 3;
  }
  return 0;
}

}  // extern "C"
