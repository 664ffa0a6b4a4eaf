// Host build of the divide-and-conquer tridiagonal eigensolver: the SAME numerical core (gpcsd_b200/csrc/dc_core.h) that the
// CUDA kernel in gpcsd_eig.cu uses, driven sequentially with one "lane".  Test infrastructure only (tests/test_dc_host.py
// compiles it with g++ and compares against LAPACK); nothing in the product links it.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../gpcsd_b200/csrc/dc_core.h"

using namespace gpcsd::dc;

// d[n], e[n] (e[i] = T[i][i-1], e[0] ignored) -> W[n] ascending, QT[n][n] rows = eigenvectors.  iters_out (optional): total
// secular iterations are not tracked here; returns the number of non-deflated roots summed over all merges.
extern "C" long dc_host_eig(int n, const double* d_in, const double* e_in, double* W, double* QT) {
  std::vector<double> d(n), e(n, 0.0);
  double scale = 0.0;
  for (int i = 0; i < n; ++i) {
    scale = std::max(scale, std::fabs(d_in[i]));
    if (i > 0) scale = std::max(scale, std::fabs(e_in[i]));
  }
  if (scale == 0.0) scale = 1.0;
  for (int i = 0; i < n; ++i) {
    d[i] = d_in[i] / scale;
    e[i] = (i > 0) ? e_in[i] / scale : 0.0;
  }
  // tear every off-diagonal: leaves of size 1
  for (int i = 0; i < n; ++i) {
    const double el = (i > 0) ? std::fabs(e[i]) : 0.0, er = (i + 1 < n) ? std::fabs(e[i + 1]) : 0.0;
    d[i] = d[i] - el - er;
  }
  std::vector<double> Qa((size_t)n * n, 0.0), Qb((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) Qa[(size_t)i * n + i] = 1.0;
  std::vector<double> dnew(n), z(n), dS(n), zS(n), dl(n), w(n), rot_c(n), rot_s(n), mu(n), zh(n);
  std::vector<int> srt(n), row(n), rot_p(n), rot_n(n), org(n);
  long kept = 0;
  const int levels = num_levels(n);
  for (int L = 1; L <= levels; ++L) {
    const int nodes = 1 << (levels - L);
    for (int p = 0; p < nodes; ++p) {
      const int a = node_start(n, 2 * nodes, 2 * p), c = node_start(n, 2 * nodes, 2 * p + 1),
                b = node_start(n, 2 * nodes, 2 * p + 2);
      const int m = b - a;
      if (m == 0) continue;
      if (c == a || c == b) {          // one child is empty: nothing to merge
        for (int g = a; g < b; ++g) {
          dnew[g] = d[g];
          for (int j = a; j < b; ++j) Qb[(size_t)g * n + j] = Qa[(size_t)g * n + j];
        }
        continue;
      }
      const double rho = 2.0 * std::fabs(e[c]), sgn = (e[c] < 0.0) ? -1.0 : 1.0;
      for (int g = a; g < b; ++g)
        z[g] = ((g < c) ? Qa[(size_t)g * n + (c - 1)] : sgn * Qa[(size_t)g * n + c]) * M_SQRT1_2;
      for (int g = a; g < b; ++g) {    // counting sort
        int r = 0;
        for (int h = a; h < b; ++h) r += (d[h] < d[g]) || (d[h] == d[g] && h < g);
        srt[a + r] = g;
        dS[a + r] = d[g];
        zS[a + r] = z[g];
      }
      int k, nrot;
      double dmax = 0.0, zmax = 0.0;
      for (int g = a; g < b; ++g) {
        dmax = std::max(dmax, std::fabs(d[g]));
        zmax = std::max(zmax, std::fabs(z[g]));
      }
      deflate(a, m, srt.data(), dS.data(), zS.data(), rho, dmax, zmax, row.data(), dl.data(), w.data(), rot_p.data(), rot_n.data(),
              rot_c.data(), rot_s.data(), k, nrot);
      for (int q = 0; q < nrot; ++q) {
        double* xp = &Qa[(size_t)rot_p[a + q] * n];
        double* xn = &Qa[(size_t)rot_n[a + q] * n];
        const double cc = rot_c[a + q], ss = rot_s[a + q];
        for (int j = a; j < b; ++j) {
          const double vp = xp[j], vn = xn[j];
          xp[j] = cc * vp + ss * vn;
          xn[j] = cc * vn - ss * vp;
        }
      }
      kept += k;
      for (int i = 0; i < k; ++i) secular_root<OneLane>(k, i, &dl[a], &w[a], rho, mu[a + i], org[a + i]);
      for (int j = 0; j < k; ++j) zh[a + j] = zhat_component<OneLane>(k, j, &dl[a], &w[a], &mu[a], &org[a]);
      for (int i = 0; i < k; ++i) {
        const double inv = inv_norm<OneLane>(k, org[a + i], mu[a + i], &dl[a], &zh[a]);
        dnew[a + i] = dl[a + org[a + i]] + mu[a + i];
        for (int col = a; col < b; ++col) {
          double acc = 0.0;
          for (int j = 0; j < k; ++j)
            acc += zh[a + j] * rcp(delta_ji(&dl[a], j, org[a + i], mu[a + i])) * Qa[(size_t)row[a + j] * n + col];
          Qb[(size_t)(a + i) * n + col] = acc * inv;
        }
      }
      for (int pos = k; pos < m; ++pos) {
        dnew[a + pos] = dl[a + pos];
        for (int col = a; col < b; ++col) Qb[(size_t)(a + pos) * n + col] = Qa[(size_t)row[a + pos] * n + col];
      }
    }
    std::swap(Qa, Qb);
    std::swap(d, dnew);
  }
  for (int g = 0; g < n; ++g) {
    int r = 0;
    for (int h = 0; h < n; ++h) r += (d[h] < d[g]) || (d[h] == d[g] && h < g);
    W[r] = d[g] * scale;
    for (int j = 0; j < n; ++j) QT[(size_t)r * n + j] = Qa[(size_t)g * n + j];
  }
  return kept;
}
