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
constexpr double STEP_TOL = 1e-9;

// reciprocal: on the device the IEEE-rounded MUFU.RCP64H + Newton sequence (~8 instructions) instead of the ~30-instruction
// division; every quotient below is a * rcp(b), within 1.5 ulp, which is all the Gu-Eisenstat argument needs
DC_HD double rcp(double x) {
#ifdef __CUDA_ARCH__
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// sqrt for x >= 0: x * rsqrt(x) on the device (MUFU.RSQ64H + Newton, ~10 instructions instead of ~30)
DC_HD double fsqrt(double x) {
#ifdef __CUDA_ARCH__
  return (x > 0.0) ? x * rsqrt(x) : 0.0;
#else
  return sqrt(x);
#endif
}

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
DC_HD int node_start(int n, int nodes, int p) { return (p * n) / nodes; }      // n, nodes < 2^15
// node containing position g
DC_HD int node_of(int n, int nodes, int g) {
  int p = (g * nodes) / n;
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
  // four poles per trip: the reciprocals (MUFU seed + Newton steps, ~80 cycles of dependent latency each) are independent
  // and overlap; a one-pole loop with a runtime trip count is a serial chain of them
  for (int j0 = X::lane(); j0 < k; j0 += 4 * X::L) {
    double t[4], wj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * X::L;
      const bool in = j < k;
      wj[u] = in ? w[j] : 0.0;
      const double del = in ? (dl[j] - dlo) - mu : 1.0;
      t[u] = wj[u] * rcp(del);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * X::L;
      const double wt = wj[u] * t[u], tt = t[u] * t[u];
      if (j <= isplit) {
        ps += wt;
        dps += tt;
      } else {
        ph += wt;
        dph += tt;
      }
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
  if (lo > 0.0 && hi > 16.0 * lo) return fsqrt(lo) * fsqrt(hi);
  if (hi < 0.0 && lo < 16.0 * hi) return -fsqrt(-lo) * fsqrt(-hi);
  if (lo == 0.0 && hi > 0.0) return hi * 9.765625e-4;     // 2^-10: walk towards the pole geometrically
  if (hi == 0.0 && lo < 0.0) return lo * 9.765625e-4;
  return 0.5 * (lo + hi);
}

// Root i of 1/rho + sum_j w_j^2 / (dl_j - lambda) = 0.  All lanes of the policy call this together with identical arguments
// and receive identical results.
template <class X>
DC_HD void secular_root(int k, int i, const double* dl, const double* w, double rho, double& mu_out, int& org_out) {
  const double rhoinv = rcp(rho);
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
    const double rhalf = rcp(half);
    const double c0 = fm + wl2 * rhalf - wr2 * rhalf;
    if (fm > 0.0) {            // root in the left half: origin = left pole
      org = pl; lo = 0.0; hi = half;
      const double A = c0 * gap + wl2 + wr2, B = wl2 * gap;
      const double s = fsqrt(fabs(A * A - 4.0 * B * c0));
      mu = (A > 0.0) ? 2.0 * B * rcp(A + s) : (A - s) * rcp(2.0 * c0);
    } else {                   // right half: origin = right pole, mu < 0
      org = pr; lo = -half; hi = 0.0;
      const double A = -c0 * gap + wl2 + wr2, B = wr2 * gap;
      const double s = fsqrt(fabs(A * A + 4.0 * B * c0));
      mu = (A > 0.0) ? -2.0 * B * rcp(A + s) : (A - s) * rcp(2.0 * c0);
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
    const double c0 = fm + wl2 * rcp(g + half) + wr2 * rcp(half);
    const double A = -c0 * g + wl2 + wr2, B = wr2 * g;
    const double s = fsqrt(fabs(A * A + 4.0 * B * c0));
    mu = (A < 0.0) ? 2.0 * B * rcp(s - A) : (A + s) * rcp(2.0 * c0);
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
    const double s = fsqrt(fabs(A * A - 4.0 * B * C));
    if (C == 0.0) {
      eta = B * rcp(A);
    } else if (!last) {
      eta = (A <= 0.0) ? (A - s) * rcp(2.0 * C) : 2.0 * B * rcp(A + s);
    } else {
      eta = (A >= 0.0) ? (A + s) * rcp(2.0 * C) : 2.0 * B * rcp(A - s);
    }
    if (!(f * eta < 0.0)) eta = -f * rcp(dw);     // wrong direction (or NaN): Newton step
    double munew = mu + eta;
    const bool inside = (munew > lo && munew < hi);
    if (!inside) munew = bracket_mid(lo, hi);
    if (!(munew > lo && munew < hi) || munew == mu) break;       // bracket exhausted
    mu = munew;
    // the rational iteration converges at least quadratically: after a step this small relative to the distance from the
    // origin pole the next residual is below rounding, so the confirming evaluation is skipped
    if (inside && fabs(eta) <= STEP_TOL * fabs(mu)) break;
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
  for (int i0 = X::lane(); i0 < k; i0 += 4 * X::L) {      // four factors per trip (independent reciprocals)
    double f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * X::L;
      if (i < k) {
        const double del = delta_ji(dl, j, org[i], mu[i]);
        f[u] = (i == j) ? fabs(del) : fabs(del * rcp(dl[j] - dl[i]));
      } else {
        f[u] = 1.0;
      }
    }
    p *= (f[0] * f[1]) * (f[2] * f[3]);
  }
  p = X::prod(p);
  return copysign(fsqrt(p), w[j]);
}

// 1 / |u_i| for the eigenvector u_i[j] = zhat[j] / (dl[j] - lambda_i)
template <class X>
DC_HD double inv_norm(int k, int org_i, double mu_i, const double* dl, const double* zhat) {
  double s = 0.0;
  for (int j = X::lane(); j < k; j += X::L) {
    const double t = zhat[j] * rcp(delta_ji(dl, j, org_i, mu_i));
    s += t * t;
  }
  return rcp(fsqrt(X::sum(s)));
}

// ---- deflation (LAPACK dlaed2 logic) ------------------------------------------------------------------------------------
// One caller per merge node; a serial scan whose loop-carried state (the last kept candidate) stays in registers and whose
// inputs are read from contiguous, sorted arrays (prefetched one element ahead).  Position-indexed arrays are addressed
// [a, a+m).
//   in : srt[a+s] = eigenvector row with the s-th smallest d, dS[a+s], zS[a+s] its d and z (|z| = 1 over the node), rho, and
//        the node's dmax = max |d|, zmax = max |z|
//   out: k; row_out[a+r], dl[a+r], w[a+r] for the k kept entries (ascending), row_out[a+pos], dl[a+pos] (final eigenvalue)
//        for the deflated ones, pos = k..m-1; Givens rotations (rows rp, rn; c, s) in application order, nrot of them.
DC_HD void deflate(int a, int m, const int* srt, const double* dS, const double* zS, double rho, double dmax, double zmax,
                   int* row_out, double* dl, double* w, int* rot_p, int* rot_n, double* rot_c, double* rot_s, int& k_out,
                   int& nrot_out) {
  const double tol = 8.0 * EPS * fmax(dmax, zmax);
  int k = 0, k2 = m, nrot = 0;
  if (!(rho * zmax > tol)) {
    for (int s = 0; s < m; ++s) {
      row_out[a + s] = srt[a + s];
      dl[a + s] = dS[a + s];
    }
    k_out = 0;
    nrot_out = 0;
    return;
  }
  int pj = -1;
  double dp = 0.0, zp = 0.0;                  // d and z of the candidate pj (as modified by earlier rotations)
  double dnx = dS[a], znx = zS[a];
  int rnx = srt[a];
  for (int s = 0; s < m; ++s) {
    const double dc = dnx, zc = znx;
    const int nj = rnx;
    if (s + 1 < m) {
      dnx = dS[a + s + 1];
      znx = zS[a + s + 1];
      rnx = srt[a + s + 1];
    }
    if (!(rho * fabs(zc) > tol)) {            // negligible coupling: eigenpair unchanged
      --k2;
      row_out[a + k2] = nj;
      dl[a + k2] = dc;
      continue;
    }
    if (pj < 0) {
      pj = nj;
      dp = dc;
      zp = zc;
      continue;
    }
    const double t = dc - dp;
    const double h2 = zc * zc + zp * zp;
    if (fabs(t * zc * zp) <= tol * h2) {      // |t c s| <= tol: rotate the pair so that z[pj] = 0 and deflate pj
      const double tau = fsqrt(h2), rt = rcp(tau), c = zc * rt, sn = -zp * rt;
      rot_p[a + nrot] = pj;
      rot_n[a + nrot] = nj;
      rot_c[a + nrot] = c;
      rot_s[a + nrot] = sn;
      ++nrot;
      const double dpn = dp * c * c + dc * sn * sn;
      --k2;
      row_out[a + k2] = pj;
      dl[a + k2] = dpn;
      dp = dp * sn * sn + dc * c * c;
      zp = tau;
      pj = nj;
    } else {
      row_out[a + k] = pj;
      dl[a + k] = dp;
      w[a + k] = zp;
      ++k;
      pj = nj;
      dp = dc;
      zp = zc;
    }
  }
  if (pj >= 0) {
    row_out[a + k] = pj;
    dl[a + k] = dp;
    w[a + k] = zp;
    ++k;
  }
  k_out = k;
  nrot_out = nrot;
}

}  // namespace dc
}  // namespace gpcsd
