// dense_host.hpp -- tiny dense kernels that stay on the host because they are latency-bound scalar work on matrices
// of order <= a few hundred (projected Rayleigh-Ritz matrices, b x b Gram factors, the nbPart x nbPart connectivity
// matrix of src/geneo.cpp:1182-1201 that the reference hands to EPS "lapack").  Everything of order n_i or N runs on
// the GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace geneo {

// Upper Cholesky G = R^T R (row-major n x n, in place: R in the upper triangle, lower zeroed).
// Returns the index of the first non-positive pivot (relative to relTol * max diag), or -1 on success.
inline int chol_upper(int n, double* g, double relTol = 1e-14) {
  double dmax = 0.;
  for (int i = 0; i < n; i++) dmax = std::max(dmax, std::fabs(g[i * n + i]));
  for (int j = 0; j < n; j++) {
    double s = g[j * n + j];
    for (int k = 0; k < j; k++) s -= g[k * n + j] * g[k * n + j];
    if (!(s > relTol * dmax)) return j;
    const double r = std::sqrt(s);
    g[j * n + j] = r;
    for (int i = j + 1; i < n; i++) {
      double t = g[j * n + i];
      for (int k = 0; k < j; k++) t -= g[k * n + j] * g[k * n + i];
      g[j * n + i] = t / r;
    }
    for (int i = 0; i < j; i++) g[j * n + i] = 0.;
  }
  return -1;
}

// inverse of an upper-triangular matrix (row-major), out-of-place
inline void triu_inverse(int n, const double* r, double* inv) {
  std::fill(inv, inv + (size_t)n * n, 0.);
  for (int j = 0; j < n; j++) {
    inv[j * n + j] = 1. / r[j * n + j];
    for (int i = j - 1; i >= 0; i--) {
      double s = 0.;
      for (int k = i + 1; k <= j; k++) s += r[i * n + k] * inv[k * n + j];
      inv[i * n + j] = -s / r[i * n + i];
    }
  }
}

// Symmetric eigen-decomposition (Householder tridiagonalisation + implicit QL).  a: row-major n x n symmetric input,
// on exit column j of a (a[i*n+j]) is the eigenvector of w[j]; eigenvalues ascending.
inline void sym_eig(int n, double* a, double* w) {
  if (n == 0) return;
  std::vector<double> e(n, 0.);
  double* d = w;
  // The input is symmetric, so the tridiagonalisation and the accumulation of the Householder transformations may run on
  // the TRANSPOSED layout: V(k, j) with k running is then contiguous -- every O(n^3) inner loop below streams memory
  // (the projected matrices of the block Lanczos solver reach n ~ 1000-1600 when a subdomain needs hundreds of pairs:
  // 5x faster than the strided walk at n = 800).
  auto V = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };
  for (int j = 0; j < n; j++) d[j] = V(n - 1, j);
  for (int i = n - 1; i > 0; i--) {
    double scale = 0., h = 0.;
    for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
    if (scale == 0.) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) { d[j] = V(i - 1, j); V(i, j) = 0.; V(j, i) = 0.; }
    } else {
      for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1], g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.;
      for (int j = 0; j < i; j++) {
        f = d[j];
        V(j, i) = f;
        g = e[j] + V(j, j) * f;
        for (int k = j + 1; k <= i - 1; k++) { g += V(k, j) * d[k]; e[k] += V(k, j) * f; }
        e[j] = g;
      }
      f = 0.;
      for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; k++) V(k, j) -= (f * e[k] + g * d[k]);
        d[j] = V(i - 1, j);
        V(i, j) = 0.;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    V(n - 1, i) = V(i, i);
    V(i, i) = 1.;
    const double h = d[i + 1];
    if (h != 0.) {
      for (int k = 0; k <= i; k++) d[k] = V(k, i + 1) / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.;
        for (int k = 0; k <= i; k++) g += V(k, i + 1) * V(k, j);
        for (int k = 0; k <= i; k++) V(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) V(k, i + 1) = 0.;
  }
  for (int j = 0; j < n; j++) { d[j] = V(n - 1, j); V(n - 1, j) = 0.; }
  V(n - 1, n - 1) = 1.;
  e[0] = 0.;
  // implicit QL.  The rotations combine two COLUMNS of V: they run on the transpose (two contiguous rows, vectorisable),
  // which is what makes the O(n^3) accumulation of the eigenvectors cheap for the projected matrices of the block
  // Lanczos solver (n up to a few hundred, once per step).
  std::vector<double> vt(a, a + (size_t)n * n);  // vt[j * n + i] = V(i, j): the storage of `a` already is that transpose
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.;
  double f = 0., tst1 = 0.;
  const double eps = std::pow(2., -52.);
  for (int l = 0; l < n; l++) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
    if (m > l) {
      int iter = 0;
      do {
        iter++;
        double g = d[l];
        double p = (d[l + 1] - g) / (2. * e[l]);
        double r = std::hypot(p, 1.);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1., c2 = c, c3 = c, s = 0., s2 = 0.;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          {
            double* __restrict__ vi = &vt[(size_t)i * n];
            double* __restrict__ vi1 = &vt[(size_t)(i + 1) * n];
            for (int k = 0; k < n; k++) {
              const double hh = vi1[k];
              vi1[k] = s * vi[k] + c * hh;
              vi[k] = c * vi[k] - s * hh;
            }
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.;
  }
  // sort ascending (selection sort on the rows of the transpose), then back to column layout
  for (int i = 0; i < n - 1; i++) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; j++)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      std::swap_ranges(&vt[(size_t)i * n], &vt[(size_t)i * n] + n, &vt[(size_t)k * n]);
    }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) a[(size_t)i * n + j] = vt[(size_t)j * n + i];  // row-major out: column j = eigenvector j
}

// Eigenvalues of a symmetric matrix plus SELECTED ROWS of its eigenvector matrix: the Rayleigh-Ritz step of the block
// Lanczos solver only needs the last b components of every Ritz vector (residual estimate ||R y_bottom||), the whole
// vectors only at a restart and at the end.  Householder tridiagonalisation (4/3 n^3, the only cubic part), the needed rows
// of Q by applying the reflectors to unit vectors (O(n^2) per row), implicit QL with the rotations applied to those rows
// only (O(n^2) per row) -- against ~7 n^3 for the full decomposition.
//   a      row-major n x n symmetric (destroyed)
//   rows   indices of the wanted rows (nrows of them)
//   w      n eigenvalues, ascending
//   yrows  nrows x n row-major: yrows[t * n + j] = component rows[t] of the eigenvector of w[j]
inline void sym_eig_rows(int n, double* a, const int* rows, int nrows, double* w, double* yrows) {
  if (n == 0) return;
  std::vector<double> d(n), e(n, 0.), tau(n, 0.), p(n), v(n);
  auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };  // lower triangle is referenced
  for (int i = 0; i + 1 < n; i++) {
    // reflector annihilating A[i+2.., i]
    const int m = n - i - 1;  // length of x = A[i+1.., i]
    double alpha = A(i + 1, i), xn = 0.;
    for (int r = i + 2; r < n; r++) xn += A(r, i) * A(r, i);
    if (xn == 0.) { tau[i] = 0.; e[i] = alpha; d[i] = A(i, i); continue; }
    const double beta = -std::copysign(std::sqrt(alpha * alpha + xn), alpha);
    tau[i] = (beta - alpha) / beta;
    const double sc = 1. / (alpha - beta);
    for (int r = i + 2; r < n; r++) A(r, i) *= sc;  // v[1..] stored below the sub-diagonal, v[0] = 1 implicit
    e[i] = beta;
    d[i] = A(i, i);
    v[0] = 1.;
    for (int r = 1; r < m; r++) v[r] = A(i + 1 + r, i);
    // p = tau * A22 v (symmetric, lower triangle, row-wise sweeps)
    std::fill(p.begin(), p.begin() + m, 0.);
    for (int r = 0; r < m; r++) {
      const double* row = &A(i + 1 + r, i + 1);
      double s1 = 0.;
      const double vr = v[r];
      for (int c = 0; c < r; c++) { s1 += row[c] * v[c]; p[c] += row[c] * vr; }
      p[r] += s1 + row[r] * vr;
    }
    double pv = 0.;
    for (int r = 0; r < m; r++) { p[r] *= tau[i]; pv += p[r] * v[r]; }
    const double hh = 0.5 * tau[i] * pv;
    for (int r = 0; r < m; r++) p[r] -= hh * v[r];  // w
    for (int r = 0; r < m; r++) {                    // A22 -= v w^T + w v^T (lower triangle)
      double* row = &A(i + 1 + r, i + 1);
      const double vr = v[r], wr = p[r];
      for (int c = 0; c <= r; c++) row[c] -= vr * p[c] + wr * v[c];
    }
  }
  d[n - 1] = A(n - 1, n - 1);
  // wanted rows of Q = H_0 H_1 ... H_{n-3}: row r = e_r^T H_0 H_1 ...  (H_i = I - tau_i v_i v_i^T acts on indices i+1..n-1)
  std::vector<double> zt((size_t)n * nrows);  // transposed: zt[j * nrows + t] = Q[rows[t], j]
  {
    std::vector<double> x(n);
    for (int t = 0; t < nrows; t++) {
      std::fill(x.begin(), x.end(), 0.);
      x[rows[t]] = 1.;
      for (int i = 0; i + 1 < n; i++) {
        if (tau[i] == 0.) continue;
        double s1 = x[i + 1];
        for (int r = i + 2; r < n; r++) s1 += x[r] * A(r, i);
        s1 *= tau[i];
        if (s1 == 0.) continue;
        x[i + 1] -= s1;
        for (int r = i + 2; r < n; r++) x[r] -= s1 * A(r, i);
      }
      for (int j = 0; j < n; j++) zt[(size_t)j * nrows + t] = x[j];
    }
  }
  // implicit QL on (d, e); rotation (i, i+1) combines columns i, i+1 of the row block = rows i, i+1 of zt
  e[n - 1] = 0.;
  double f = 0., tst1 = 0.;
  const double eps = std::pow(2., -52.);
  for (int l = 0; l < n; l++) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
    if (m > l) {
      int iter = 0;
      do {
        iter++;
        double g = d[l];
        double pp = (d[l + 1] - g) / (2. * e[l]);
        double r = std::hypot(pp, 1.);
        if (pp < 0) r = -r;
        d[l] = e[l] / (pp + r);
        d[l + 1] = e[l] * (pp + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        pp = d[m];
        double c = 1., c2 = c, c3 = c, s = 0., s2 = 0.;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * pp;
          r = std::hypot(pp, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = pp / r;
          pp = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          double* __restrict__ vi = &zt[(size_t)i * nrows];
          double* __restrict__ vi1 = &zt[(size_t)(i + 1) * nrows];
          for (int k = 0; k < nrows; k++) {
            const double t1 = vi1[k];
            vi1[k] = s * vi[k] + c * t1;
            vi[k] = c * vi[k] - s * t1;
          }
        }
        pp = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * pp;
        d[l] = c * pp;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.;
  }
  std::vector<int> ord(n);
  for (int i = 0; i < n; i++) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int x, int y) { return d[x] < d[y]; });
  for (int j = 0; j < n; j++) {
    w[j] = d[ord[j]];
    for (int t = 0; t < nrows; t++) yrows[(size_t)t * n + j] = zt[(size_t)ord[j] * nrows + t];
  }
}

}  // namespace geneo
