// Numerical core of the divide-and-conquer symmetric tridiagonal eigensolver (Cuppen's method with Gu-Eisenstat
// eigenvector stabilisation), written so the SAME source compiles for the device (gpcsd_eig.cu: one warp cooperates on one
// secular root / one z component) and for the host (tests/dc_host_harness.cpp: one "lane"), which is how the numerics are
// validated on machines without a GPU.
//
// Problem at every merge node: eigen-decomposition of  diag(dl) + rho * w w^T  (dl ascending, k entries left after
// deflation).  Conventions shared by all functions:
//   * root i lies in (dl[i], dl[i+1]) for i < k-1 and in (dl[k-1], dl[k-1] + rho*|w|^2] for i = k-1;
//   * a root is stored as (org, mu): lambda_i = dl[org] + mu with org the NEAREST pole, so every difference
//     dl[j] - lambda_i = (dl[j] - dl[org]) - mu is computed to high relative accuracy (the property the Gu-Eisenstat
//     recomputation of w needs for numerically orthogonal eigenvectors).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define DC_HD __host__ __device__ __forceinline__
#else
#define DC_HD inline
#endif

namespace gpcsd {
namespace dc {

constexpr double EPS = 1.1102230246251565e-16;   // 2^-53 (LAPACK dlamch('E'))
constexpr int MAX_ITER = 80;

// ---- lane policies: how many cooperating lanes evaluate one sum ---------------------------------------------------------
struct OneLane {
  static constexpr int L = 1;
  DC_HD static int lane() { return 0; }
  DC_HD static double sum(double x) { return x; }
  DC_HD static double prod(double x) { return x; }
};
#ifdef __CUDACC__
struct WarpLanes {
  static constexpr int L = 32;
  __device__ __forceinline__ static int lane() { return threadIdx.x & 31; }
  __device__ __forceinline__ static double sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  }
  __device__ __forceinline__ static double prod(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x *= __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  }
};
#endif

// ---- tree geometry: ceil(log2 n) levels over leaves of size 0 or 1 ------------------------------------------------------
DC_HD int num_levels(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}
// first index of node p when [0, n) is cut into `nodes` nearly equal consecutive ranges
DC_HD int node_start(int n, int nodes, int p) { return (int)(((long long)p * n) / nodes); }
// node containing position g
DC_HD int node_of(int n, int nodes, int g) {
  int p = (int)(((long long)g * nodes) / n);
  if (node_start(n, nodes, p + 1) <= g) ++p;
  return p;
}

// ---- secular function ---------------------------------------------------------------------------------------------------
// psi = sum_{j <= isplit} w_j^2 / (dl_j - lambda), phi = sum_{j > isplit}, and the derivative sums; lambda = dl[org] + mu.
template <class X>
DC_HD void secular_eval(int k, int isplit, int org, double mu, const double* dl, const double* w, double& psi, double& phi,
                        double& dpsi, double& dphi) {
  const double dlo = dl[org];
  double ps = 0.0, ph = 0.0, dps = 0.0, dph = 0.0;
  for (int j = X::lane(); j < k; j += X::L) {
    const double del = (dl[j] - dlo) - mu;
    const double t = w[j] / del;
    const double wt = w[j] * t, tt = t * t;
    if (j <= isplit) {
      ps += wt;
      dps += tt;
    } else {
      ph += wt;
      dph += tt;
    }
  }
  psi = X::sum(ps);
  phi = X::sum(ph);
  dpsi = X::sum(dps);
  dphi = X::sum(dph);
}

// next point strictly inside the bracket (lo, hi): arithmetic midpoint, or the geometric one when the ends have the same
// sign and differ by orders of magnitude (roots that sit extremely close to their pole)
DC_HD double bracket_mid(double lo, double hi) {
  if (lo > 0.0 && hi > 16.0 * lo) return sqrt(lo) * sqrt(hi);
  if (hi < 0.0 && lo < 16.0 * hi) return -sqrt(-lo) * sqrt(-hi);
  if (lo == 0.0 && hi > 0.0) return hi * 9.765625e-4;     // 2^-10: walk towards the pole geometrically
  if (hi == 0.0 && lo < 0.0) return lo * 9.765625e-4;
  return 0.5 * (lo + hi);
}

// Root i of 1/rho + sum_j w_j^2 / (dl_j - lambda) = 0.  All lanes of the policy call this together with identical arguments
// and receive identical results.
template <class X>
DC_HD void secular_root(int k, int i, const double* dl, const double* w, double rho, double& mu_out, int& org_out) {
  const double rhoinv = 1.0 / rho;
  if (k == 1) {
    org_out = 0;
    mu_out = rho * w[0] * w[0];
    return;
  }
  const bool last = (i == k - 1);
  // two poles of the rational model: pl (left) and pr (right); psi covers j <= isplit
  const int pl = last ? k - 2 : i, pr = last ? k - 1 : i + 1, isplit = pl;
  int org;
  double lo, hi, mu;
  double psi, phi, dpsi, dphi;
  if (!last) {
    const double gap = dl[pr] - dl[pl], half = 0.5 * gap;
    secular_eval<X>(k, isplit, pl, half, dl, w, psi, phi, dpsi, dphi);
    const double fm = rhoinv + psi + phi;
    const double wl2 = w[pl] * w[pl], wr2 = w[pr] * w[pr];
    // value of everything but the two nearest poles at the midpoint
    const double c0 = fm - wl2 / (-half) - wr2 / half;
    if (fm > 0.0) {            // root in the left half: origin = left pole
      org = pl; lo = 0.0; hi = half;
      const double A = c0 * gap + wl2 + wr2, B = wl2 * gap;
      const double s = sqrt(fabs(A * A - 4.0 * B * c0));
      mu = (A > 0.0) ? 2.0 * B / (A + s) : (A - s) / (2.0 * c0);
    } else {                   // right half: origin = right pole, mu < 0
      org = pr; lo = -half; hi = 0.0;
      const double A = -c0 * gap + wl2 + wr2, B = wr2 * gap;
      const double s = sqrt(fabs(A * A + 4.0 * B * c0));
      mu = (A > 0.0) ? -2.0 * B / (A + s) : (A - s) / (2.0 * c0);
    }
  } else {
    double wsq = 0.0;
    for (int j = X::lane(); j < k; j += X::L) wsq += w[j] * w[j];
    wsq = X::sum(wsq);
    org = pr; lo = 0.0; hi = rho * wsq;
    const double half = 0.5 * hi;
    secular_eval<X>(k, isplit, org, half, dl, w, psi, phi, dpsi, dphi);
    const double fm = rhoinv + psi + phi;
    if (fm > 0.0) hi = half; else lo = half;
    const double g = dl[pr] - dl[pl];
    const double wl2 = w[pl] * w[pl], wr2 = w[pr] * w[pr];
    const double c0 = fm - wl2 / (-g - half) - wr2 / (-half);
    const double A = -c0 * g + wl2 + wr2, B = wr2 * g;
    const double s = sqrt(fabs(A * A + 4.0 * B * c0));
    mu = (A < 0.0) ? 2.0 * B / (s - A) : (A + s) / (2.0 * c0);
  }
  if (!(mu > lo && mu < hi)) mu = bracket_mid(lo, hi);

  for (int it = 0; it < MAX_ITER; ++it) {
    secular_eval<X>(k, isplit, org, mu, dl, w, psi, phi, dpsi, dphi);
    const double f = rhoinv + psi + phi;
    const double dw = dpsi + dphi;
    if (f > 0.0) hi = mu; else lo = mu;
    if (!(fabs(f) > EPS * (16.0 * (rhoinv + fabs(psi) + fabs(phi)) + fabs(mu) * dw))) break;   // also exits on NaN
    // two-pole rational interpolation ("middle way"): psi ~ s + S/(pl - x), phi ~ r + R/(pr - x)
    const double D0 = (dl[pl] - dl[org]) - mu, D1 = (dl[pr] - dl[org]) - mu;
    const double C = f - D0 * dpsi - D1 * dphi;
    const double A = (D0 + D1) * f - D0 * D1 * dw;
    const double B = D0 * D1 * f;
    double eta;
    const double s = sqrt(fabs(A * A - 4.0 * B * C));
    if (C == 0.0) {
      eta = B / A;
    } else if (!last) {
      eta = (A <= 0.0) ? (A - s) / (2.0 * C) : 2.0 * B / (A + s);
    } else {
      eta = (A >= 0.0) ? (A + s) / (2.0 * C) : 2.0 * B / (A - s);
    }
    if (!(f * eta < 0.0)) eta = -f / dw;          // wrong direction (or NaN): Newton step
    double munew = mu + eta;
    if (!(munew > lo && munew < hi)) munew = bracket_mid(lo, hi);
    if (!(munew > lo && munew < hi) || munew == mu) break;       // bracket exhausted
    mu = munew;
  }
  mu_out = mu;
  org_out = org;
}

// dl[j] - lambda_i from the stored (org, mu) representation of root i
DC_HD double delta_ji(const double* dl, int j, int org_i, double mu_i) { return (dl[j] - dl[org_i]) - mu_i; }

// Gu-Eisenstat: the w for which the COMPUTED roots are the exact eigenvalues of diag(dl) + rho w w^T (up to a common factor
// that the eigenvector normalisation removes):  w_j^2 = prod_i (lambda_i - dl_j) / prod_{i != j} (dl_i - dl_j).
template <class X>
DC_HD double zhat_component(int k, int j, const double* dl, const double* w, const double* mu, const int* org) {
  double p = 1.0;
  for (int i = X::lane(); i < k; i += X::L) {
    const double del = delta_ji(dl, j, org[i], mu[i]);
    p *= (i == j) ? fabs(del) : fabs(del / (dl[j] - dl[i]));
  }
  p = X::prod(p);
  return copysign(sqrt(p), w[j]);
}

// 1 / |u_i| for the eigenvector u_i[j] = zhat[j] / (dl[j] - lambda_i)
template <class X>
DC_HD double inv_norm(int k, int org_i, double mu_i, const double* dl, const double* zhat) {
  double s = 0.0;
  for (int j = X::lane(); j < k; j += X::L) {
    const double t = zhat[j] / delta_ji(dl, j, org_i, mu_i);
    s += t * t;
  }
  return 1.0 / sqrt(X::sum(s));
}

// ---- deflation (LAPACK dlaed2 logic) ------------------------------------------------------------------------------------
// One caller per merge node.  Position-indexed arrays are addressed [a, a+m); d, z are indexed by eigenvector ROW (global).
//   in : srt[a+s] = row with the s-th smallest d;  d[row], z[row] (|z| = 1 over the node), rho, and the node's
//        dmax = max |d|, zmax = max |z|
//   out: k; row_out[a+r], dl[a+r], w[a+r] for the k kept entries (ascending), row_out[a+pos], dl[a+pos] (final eigenvalue)
//        for the deflated ones, pos = k..m-1; Givens rotations (rows rp, rn; c, s) in application order, nrot of them.
// d and z are updated in place by the rotations.
DC_HD void deflate(int a, int m, const int* srt, double* d, double* z, double rho, double dmax, double zmax, int* row_out,
                   double* dl, double* w, int* rot_p, int* rot_n, double* rot_c, double* rot_s, int& k_out, int& nrot_out) {
  const double tol = 8.0 * EPS * fmax(dmax, zmax);
  int k = 0, k2 = m, nrot = 0;
  if (!(rho * zmax > tol)) {
    for (int s = 0; s < m; ++s) {
      const int r = srt[a + s];
      row_out[a + s] = r;
      dl[a + s] = d[r];
    }
    k_out = 0;
    nrot_out = 0;
    return;
  }
  int pj = -1;
  for (int s = 0; s < m; ++s) {
    const int nj = srt[a + s];
    if (!(rho * fabs(z[nj]) > tol)) {          // negligible coupling: eigenpair unchanged
      --k2;
      row_out[a + k2] = nj;
      dl[a + k2] = d[nj];
      continue;
    }
    if (pj < 0) {
      pj = nj;
      continue;
    }
    const double zs = z[pj], zc = z[nj], t = d[nj] - d[pj];
    const double h2 = zc * zc + zs * zs;
    if (fabs(t * zc * zs) <= tol * h2) {       // |t c s| <= tol: rotate the pair so that z[pj] = 0 and deflate pj
      const double tau = sqrt(h2), c = zc / tau, sn = -zs / tau;
      z[nj] = tau;
      z[pj] = 0.0;
      rot_p[a + nrot] = pj;
      rot_n[a + nrot] = nj;
      rot_c[a + nrot] = c;
      rot_s[a + nrot] = sn;
      ++nrot;
      const double dp = d[pj] * c * c + d[nj] * sn * sn;
      d[nj] = d[pj] * sn * sn + d[nj] * c * c;
      d[pj] = dp;
      --k2;
      row_out[a + k2] = pj;
      dl[a + k2] = dp;
      pj = nj;
    } else {
      row_out[a + k] = pj;
      dl[a + k] = d[pj];
      w[a + k] = z[pj];
      ++k;
      pj = nj;
    }
  }
  if (pj >= 0) {
    row_out[a + k] = pj;
    dl[a + k] = d[pj];
    w[a + k] = z[pj];
    ++k;
  }
  k_out = k;
  nrot_out = nrot;
}

}  // namespace dc
}  // namespace gpcsd
