// Callers either side of the hot path, on the device:
//   * forward-model operators  fwd_model_1d / fwd_model_2d  (forward_models.py:20-39, 57-81) as weight matrices for ONE
//     gpcsd_dgemm instead of the reference's Python double loop over time x location;
//   * sample_prior (gpcsd1d.py:295-309, gpcsd2d.py:336-360): blocked Cholesky of the small factors, counter-based
//     Philox4x32-10 normal generator, fused additive noise -- configs[4]-scale synthetic LFP never touches the host;
//   * the per-trial evoked-shift objective of auditory_lfp/fit_mean_function.py:311-321: shifted-mean residual, per-trial
//     quadratic forms and the analytic shift gradient, for ALL trials of a batch at once.
#include "common.h"
#include "dmma_gemm.cuh"

namespace gpcsd {

// ------------------------------------------------------------------------------------------------
// forward-model operators
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double trapz_weight(const double* __restrict__ x, int n, int k) {
  // numpy.trapz / scipy.integrate.trapz weights on a non-uniform grid: 0.5*(x[k+1]-x[k-1]) inside, half intervals at the ends
  if (n < 2) return 0.0;
  const double lo = (k > 0) ? x[k - 1] : x[0], hi = (k < n - 1) ? x[k + 1] : x[n - 1];
  return 0.5 * (hi - lo);
}

__global__ void fwd_operator_1d_kernel(int nz, const double* __restrict__ z, int nx, const double* __restrict__ x, double R,
                                       double scale, double* __restrict__ W, long ld) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nz * nx) return;
  const int i = (int)(idx / nx), k = (int)(idx % nx);
  const double d = (z[i] - x[k]) / R, qd = d * d;
  W[(long)i * ld + k] = scale * (sqrt(qd + 1.0) - sqrt(qd)) * trapz_weight(x, nx, k);   // b_fwd_1d, forward_models.py:16
}

__global__ void fwd_operator_2d_kernel(int nz, const double* __restrict__ z, int nx1, const double* __restrict__ x1, int nx2,
                                       const double* __restrict__ x2, double R, double eps, double* __restrict__ W, long ld) {
  const long G = (long)nx1 * nx2;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nz * G) return;
  const int i = (int)(idx / G), gq = (int)(idx % G);
  const int a = gq / nx2, b = gq % nx2;
  const double d1 = z[2 * i] - x1[a], d2 = z[2 * i + 1] - x2[b];
  const double w2 = d1 * d1 + d2 * d2, Re = R + eps;
  const double wt = log(Re + sqrt(Re * Re + w2)) - log(eps + sqrt(eps * eps + w2));      // b_fwd_2d, forward_models.py:53
  W[(long)i * ld + gq] = wt * trapz_weight(x1, nx1, a) * trapz_weight(x2, nx2, b);
}

// ------------------------------------------------------------------------------------------------
// blocked right-looking Cholesky (lower), block size 32: diagonal block in one CTA, panel solve one row per thread,
// trailing update as 32x32 tiles of the lower triangle.  info: 0 ok, k+1 = first non-positive pivot (numpy raises
// LinAlgError there; the Python layer does the same).
// ------------------------------------------------------------------------------------------------
constexpr int CB = 32;

__global__ void chol_diag_kernel(int n, int k0, double* __restrict__ L, long ld, int* __restrict__ info) {
  __shared__ double s[CB][CB + 1];
  const int nb = min(CB, n - k0);
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (tx < nb && ty < nb) s[ty][tx] = L[(long)(k0 + ty) * ld + k0 + tx];
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    if (tx == 0 && ty == 0) {
      const double d = s[k][k];
      if (!(d > 0.0)) {
        if (*info == 0) *info = k0 + k + 1;
        s[k][k] = nan("");
      } else {
        s[k][k] = sqrt(d);
      }
    }
    __syncthreads();
    if (ty == 0 && tx > k && tx < nb) s[tx][k] /= s[k][k];
    __syncthreads();
    if (tx > k && ty > k && tx <= ty && ty < nb) s[ty][tx] -= s[ty][k] * s[tx][k];
    __syncthreads();
  }
  if (tx < nb && ty < nb) L[(long)(k0 + ty) * ld + k0 + tx] = (tx <= ty) ? s[ty][tx] : 0.0;
}

// rows below the diagonal block: solve X L11^T = A21 by forward substitution, one row per thread; zero the block row above
__global__ void chol_panel_kernel(int n, int k0, double* __restrict__ L, long ld) {
  __shared__ double l11[CB][CB + 1];
  const int nb = min(CB, n - k0);
  for (int e = threadIdx.x; e < CB * CB; e += blockDim.x) {
    const int r = e / CB, c = e % CB;
    l11[r][c] = (r < nb && c < nb) ? L[(long)(k0 + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  const int i = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) x[c] = (c < nb) ? L[(long)i * ld + k0 + c] : 0.0;
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    if (c < nb) {
      double v = x[c];
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < c) v -= x[k] * l11[c][k];
      x[c] = v / l11[c][c];
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c)
    if (c < nb) L[(long)i * ld + k0 + c] = x[c];
}

// A22 -= L21 L21^T on the lower-triangular 32x32 tiles of the trailing matrix; the strictly upper part of the block
// column is zeroed by chol_zero_upper_kernel once at the end
__global__ void chol_update_kernel(int n, int k0, double* __restrict__ L, long ld) {
  const int nb = min(CB, n - k0);
  const int base = k0 + nb;
  int t = blockIdx.x, tm = 0;
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  const int tn = t - tm * (tm + 1) / 2;
  __shared__ double a[CB][CB + 1], b[CB][CB + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ri = base + tm * CB + ty, rj = base + tn * CB + ty;
  a[ty][tx] = (ri < n && tx < nb) ? L[(long)ri * ld + k0 + tx] : 0.0;
  b[ty][tx] = (rj < n && tx < nb) ? L[(long)rj * ld + k0 + tx] : 0.0;
  __syncthreads();
  const int i = base + tm * CB + ty, j = base + tn * CB + tx;
  if (i < n && j < n && j <= i) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < CB; ++k) acc += a[ty][k] * b[tx][k];
    L[(long)i * ld + j] -= acc;
  }
}

__global__ void chol_zero_upper_kernel(int n, double* __restrict__ L, long ld) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  if (j > i) L[(long)i * ld + j] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) counter-based generator -> standard normals by Box-Muller.
// Counter = (pair index lo, pair index hi, stream id, 0), key = (seed lo, seed hi): element 2p and 2p+1 of the output come
// from counter p, whatever the launch geometry -- the same (seed, stream) always yields the same array.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void philox_normal_pair(unsigned long long pair, unsigned long long seed, uint32_t stream_id,
                                                   double& z0, double& z1) {
  uint32_t r[4];
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  // two 53-bit uniforms in (0, 1): ((hi >> 5) * 2^26 + (lo >> 6) + 0.5) / 2^53
  const double u1 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

// out[i] (mode 0) or out[i] += sd * z_i (mode 1) for the logical index i = row*ncols + col of a [nrows][ld] array whose first
// ncols columns are live (padding columns are left untouched)
__global__ void randn_kernel(long nrows, long ncols, long ld, unsigned long long seed, uint32_t stream_id, double sd, int accumulate,
                             double* __restrict__ out) {
  const long npairs = (nrows * ncols + 1) / 2;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long)gridDim.x * blockDim.x) {
    double z0, z1;
    philox_normal_pair((unsigned long long)p, seed, stream_id, z0, z1);
    const long i0 = 2 * p, i1 = 2 * p + 1;
    const long a0 = (i0 / ncols) * ld + (i0 % ncols);
    if (accumulate) out[a0] += sd * z0; else out[a0] = sd * z0;
    if (i1 < nrows * ncols) {
      const long a1 = (i1 / ncols) * ld + (i1 % ncols);
      if (accumulate) out[a1] += sd * z1; else out[a1] = sd * z1;
    }
  }
}

__global__ void philox_raw_kernel(long ncounters, unsigned long long seed, uint32_t stream_id, uint32_t* __restrict__ out) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncounters) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)p, (uint32_t)((unsigned long long)p >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  for (int k = 0; k < 4; ++k) out[4 * p + k] = r[k];
}

// ------------------------------------------------------------------------------------------------
// per-trial evoked-shift objective (auditory_lfp/fit_mean_function.py:311-321)
//   resid_r = Y_r - mu_0 - sum_s interp(mu_s, t + tau[r][s])     (scipy interp1d, linear, fill_value="extrapolate")
//   nll_r   = 1/2 sum_ij alpha_r,ij^2 / D_ij + 1/2 sum_s ((tau[r][s] - mutau)/sigtau)^2,   alpha_r = Qs^T resid_r Qt
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int interval_of(const double* __restrict__ t, int nt, double q, int uniform, double t0, double inv_dt) {
  // index k in [0, nt-2] of the grid interval used for q (end intervals extrapolate)
  int k;
  if (uniform) {
    k = (int)floor((q - t0) * inv_dt);
  } else {
    int lo = 0, hi = nt - 1;              // largest k with t[k] <= q
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (t[mid] <= q) lo = mid; else hi = mid;
    }
    k = lo;
  }
  return max(0, min(nt - 2, k));
}

// R[i][j][r] = Y[i][j][r] - mu[0][i][j] - sum_s lerp(mu[s][i][:], t_j + tau[r][s]).  One thread per (i, j, r), r fastest.
__global__ void shift_residual_kernel(int nx, int nt, int N, long ldn, const double* __restrict__ Y, int nseg,
                                      const double* __restrict__ mu, const double* __restrict__ t, int uniform,
                                      const double* __restrict__ tau, double* __restrict__ Rout) {
  const long total = (long)nx * nt * ldn;
  const double t0 = t[0], inv_dt = (nt > 1) ? 1.0 / (t[1] - t[0]) : 0.0;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % ldn);
    const long ij = idx / ldn;
    if (r >= N) {
      Rout[idx] = 0.0;
      continue;
    }
    const int j = (int)(ij % nt), i = (int)(ij / nt);
    double v = Y[idx] - mu[(long)i * nt + j];
    for (int s = 0; s < nseg; ++s) {
      const double q = t[j] + tau[(long)r * nseg + s];
      const int k = interval_of(t, nt, q, uniform, t0, inv_dt);
      const double* m = mu + ((long)(s + 1) * nx + i) * nt;
      const double slope = (m[k + 1] - m[k]) / (t[k + 1] - t[k]);
      v -= m[k] + slope * (q - t[k]);
    }
    Rout[idx] = v;
  }
}

// quad[r] = sum_ij B[i][j][r]^2 / rD[i][j]   (= sum alpha^2 / D with B = alpha / D).  Grid: blocks over trial columns x row
// chunks; partial sums per row chunk in ws[chunk][ldn], fixed-order second pass.
__global__ void quad_per_trial_kernel(int nrows, int nt, long ldn, int N, const double* __restrict__ B, const double* __restrict__ rD,
                                      long ldrd, int rows_per_chunk, double* __restrict__ ws) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  if (r >= N) return;
  const int row0 = chunk * rows_per_chunk, row1 = min(nrows, row0 + rows_per_chunk);
  double acc = 0.0;
  for (int row = row0; row < row1; ++row) {
    const int i = row / nt, j = row - i * nt;
    const double b = B[(long)row * ldn + r];
    acc += b * b / rD[(long)i * ldrd + j];
  }
  ws[(long)chunk * ldn + r] = acc;
}

__global__ void colsum_chunks_kernel(int nchunks, long ldn, int N, int ncomp, const double* __restrict__ ws, double* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;     // over [N][ncomp] packed as ws[chunk][comp][ldn]
  if (e >= (long)N * ncomp) return;
  const int r = (int)(e % N), c = (int)(e / N);
  double acc = 0.0;
  for (int k = 0; k < nchunks; ++k) acc += ws[((long)k * ncomp + c) * ldn + r];
  out[(long)r * ncomp + c] = acc;
}

// g[r][s] = - sum_ij V[i][j][r] * d/dtau lerp(mu_s[i][:], t_j + tau[r][s]),  V = K^{-1} resid = Qs B Qt^T
__global__ void shift_grad_kernel(int nx, int nt, int N, long ldn, const double* __restrict__ V, int nseg,
                                  const double* __restrict__ mu, const double* __restrict__ t, int uniform,
                                  const double* __restrict__ tau, int rows_per_chunk, double* __restrict__ ws) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  if (r >= N) return;
  const int nrows = nx * nt;
  const int row0 = chunk * rows_per_chunk, row1 = min(nrows, row0 + rows_per_chunk);
  const double t0 = t[0], inv_dt = (nt > 1) ? 1.0 / (t[1] - t[0]) : 0.0;
  for (int s = 0; s < nseg; ++s) {
    const double ts = tau[(long)r * nseg + s];
    double acc = 0.0;
    for (int row = row0; row < row1; ++row) {
      const int i = row / nt, j = row - i * nt;
      const int k = interval_of(t, nt, t[j] + ts, uniform, t0, inv_dt);
      const double* m = mu + ((long)(s + 1) * nx + i) * nt;
      acc -= V[(long)row * ldn + r] * (m[k + 1] - m[k]) / (t[k + 1] - t[k]);
    }
    ws[((long)chunk * nseg + s) * ldn + r] = acc;
  }
}

}  // namespace gpcsd

using namespace gpcsd;

#define GRID1D(n) (unsigned)(((long)(n) + 255) / 256), 256

static int capped_blocks(long n, int per_sm) {
  long b = (n + 255) / 256;
  const long cap = (long)per_sm * gp_num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" {

int gpcsd_fwd_operator_1d(int nz, const double* z, int nx, const double* x, double R, double scale, double* W, long ld,
                          void* stream) {
  if (nz <= 0 || nx <= 0) return 0;
  fwd_operator_1d_kernel<<<GRID1D((long)nz * nx), 0, (cudaStream_t)stream>>>(nz, z, nx, x, R, scale, W, ld);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_fwd_operator_2d(int nz, const double* z, int nx1, const double* x1, int nx2, const double* x2, double R, double eps,
                          double* W, long ld, void* stream) {
  if (nz <= 0 || nx1 <= 0 || nx2 <= 0) return 0;
  fwd_operator_2d_kernel<<<GRID1D((long)nz * nx1 * nx2), 0, (cudaStream_t)stream>>>(nz, z, nx1, x1, nx2, x2, R, eps, W, ld);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_cholesky(int n, double* L, long ld, int* info, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GP_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  for (int k0 = 0; k0 < n; k0 += CB) {
    const int nb = (n - k0 < CB) ? n - k0 : CB;
    chol_diag_kernel<<<1, dim3(CB, CB), 0, st>>>(n, k0, L, ld, info);
    GP_CUDA(cudaGetLastError());
    const int m = n - k0 - nb;
    if (m <= 0) break;
    chol_panel_kernel<<<(m + 127) / 128, 128, 0, st>>>(n, k0, L, ld);
    GP_CUDA(cudaGetLastError());
    const long t1 = (m + CB - 1) / CB;
    chol_update_kernel<<<(unsigned)(t1 * (t1 + 1) / 2), dim3(CB, CB), 0, st>>>(n, k0, L, ld);
    GP_CUDA(cudaGetLastError());
  }
  chol_zero_upper_kernel<<<GRID1D((long)n * n), 0, st>>>(n, L, ld);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_randn(long nrows, long ncols, long ld, unsigned long long seed, unsigned int stream_id, double sd, int accumulate,
                double* out, void* stream) {
  if (nrows <= 0 || ncols <= 0) return 0;
  if (ld < ncols) return gp_fail("randn: ld < ncols");
  const long npairs = (nrows * ncols + 1) / 2;
  randn_kernel<<<capped_blocks(npairs, 16), 256, 0, (cudaStream_t)stream>>>(nrows, ncols, ld, seed, stream_id, sd, accumulate, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_philox_raw(long ncounters, unsigned long long seed, unsigned int stream_id, unsigned int* out, void* stream) {
  if (ncounters <= 0) return 0;
  philox_raw_kernel<<<GRID1D(ncounters), 0, (cudaStream_t)stream>>>(ncounters, seed, stream_id, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_shift_residual(int nx, int nt, int ntrials, long ldn, const double* Y, int nseg, const double* mu, const double* t,
                         int t_uniform, const double* tau, double* Rout, void* stream) {
  if (nt < 2) return gp_fail("shift_residual: needs at least two time points");
  const long total = (long)nx * nt * ldn;
  if (total <= 0) return 0;
  shift_residual_kernel<<<capped_blocks(total, 32), 256, 0, (cudaStream_t)stream>>>(nx, nt, ntrials, ldn, Y, nseg, mu, t, t_uniform,
                                                                                  tau, Rout);
  GP_CUDA(cudaGetLastError());
  return 0;
}

static int trial_chunks(int nrows, int ntrials, int* rows_per_chunk) {
  // enough (column block x row chunk) CTAs to fill the GPU a few times over; row chunks of whole multiples of 8 rows
  const long colblocks = (ntrials + 127) / 128;
  long want = (4L * gp_num_sms() + colblocks - 1) / colblocks;
  if (want < 1) want = 1;
  if (want > nrows) want = nrows;
  int rpc = (int)((nrows + want - 1) / want);
  if (rpc < 1) rpc = 1;
  *rows_per_chunk = rpc;
  return (nrows + rpc - 1) / rpc;
}

long gpcsd_per_trial_ws_doubles(int nx, int nt, int ntrials, long ldn, int ncomp) {
  int rpc;
  const int nch = trial_chunks(nx * nt, ntrials, &rpc);
  return (long)nch * (ncomp < 1 ? 1 : ncomp) * ldn;
}

int gpcsd_quad_per_trial(int nx, int nt, int ntrials, long ldn, const double* B, const double* rD, long ldrd, double* ws,
                         double* out, void* stream) {
  if (ntrials <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rpc;
  const int nch = trial_chunks(nx * nt, ntrials, &rpc);
  quad_per_trial_kernel<<<dim3((ntrials + 127) / 128, nch), 128, 0, st>>>(nx * nt, nt, ldn, ntrials, B, rD, ldrd, rpc, ws);
  GP_CUDA(cudaGetLastError());
  colsum_chunks_kernel<<<GRID1D(ntrials), 0, st>>>(nch, ldn, ntrials, 1, ws, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

int gpcsd_shift_grad(int nx, int nt, int ntrials, long ldn, const double* V, int nseg, const double* mu, const double* t,
                     int t_uniform, const double* tau, double* ws, double* out, void* stream) {
  if (ntrials <= 0 || nseg <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rpc;
  const int nch = trial_chunks(nx * nt, ntrials, &rpc);
  shift_grad_kernel<<<dim3((ntrials + 127) / 128, nch), 128, 0, st>>>(nx, nt, ntrials, ldn, V, nseg, mu, t, t_uniform, tau, rpc, ws);
  GP_CUDA(cudaGetLastError());
  colsum_chunks_kernel<<<GRID1D((long)ntrials * nseg), 0, st>>>(nch, ldn, ntrials, nseg, ws, out);
  GP_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
