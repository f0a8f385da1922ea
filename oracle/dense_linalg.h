// ORACLE — test infrastructure only (see jet.h header).
//
// Dense SPD factorisation used by the oracle's exact Schur solve. The reference hands the
// reduced camera system to Ceres (SPARSE_SCHUR + Eigen SimplicialLDLT by default,
// sfm/bundle_adjustment/bundle_adjustment.h:98-104); an exact Cholesky of the same matrix
// is the same linear algebra up to rounding.
#ifndef ORACLE_DENSE_LINALG_H_
#define ORACLE_DENSE_LINALG_H_

#include <algorithm>
#include <cmath>
#include <vector>

namespace oracle {

// In-place blocked Cholesky of the lower triangle of a row-major n x n matrix (leading
// dimension n). Returns false if a non-positive pivot is met.
inline bool CholeskyLower(double* A, int n) {
  const int NB = 96;
  std::vector<double> Pt;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = std::min(NB, n - k0);
    // Diagonal block, unblocked.
    for (int j = k0; j < k0 + kb; ++j) {
      double d = A[(size_t)j * n + j];
      for (int k = k0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      d = std::sqrt(d);
      A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k0 + kb; ++i) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / d;
      }
    }
    const int r0 = k0 + kb;
    const int m = n - r0;
    if (m <= 0) break;
    // Panel: L_ik = A_ik L_kk^-T, one independent forward substitution per row.
#pragma omp parallel for schedule(static)
    for (int i = r0; i < n; ++i) {
      double* row = A + (size_t)i * n;
      for (int j = k0; j < k0 + kb; ++j) {
        double s = row[j];
        const double* lj = A + (size_t)j * n;
        for (int k = k0; k < j; ++k) s -= row[k] * lj[k];
        row[j] = s / lj[j];
      }
    }
    // Transposed copy of the panel so the trailing update is an axpy over contiguous j.
    Pt.resize((size_t)kb * m);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < kb; ++k) Pt[(size_t)k * m + i] = A[(size_t)(r0 + i) * n + k0 + k];
    // Trailing update of the lower triangle: A_ij -= sum_k L_ik L_jk, j <= i.
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < m; ++i) {
      double* crow = A + (size_t)(r0 + i) * n + r0;
      const double* lrow = A + (size_t)(r0 + i) * n + k0;
      for (int k = 0; k < kb; ++k) {
        const double lik = lrow[k];
        const double* p = Pt.data() + (size_t)k * m;
        for (int j = 0; j <= i; ++j) crow[j] -= lik * p[j];
      }
    }
  }
  return true;
}

// Solve L L^T x = b in place (row-major lower factor).
inline void CholeskySolveLower(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    const double* row = L + (size_t)i * n;
    for (int k = 0; k < i; ++k) s -= row[k] * b[k];
    b[i] = s / row[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    b[i] /= L[(size_t)i * n + i];
    const double bi = b[i];
    const double* row = L + (size_t)i * n;
    for (int k = 0; k < i; ++k) b[k] -= row[k] * bi;
  }
}

// Small dense SPD solve (n <= 4) by Cholesky; returns false if not positive definite.
inline bool SmallCholeskySolve(int n, const double* A, const double* b, double* x, double* Linv_out = nullptr) {
  double L[16];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        L[i * n + i] = std::sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  (void)Linv_out;
  double y[4];
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
  return true;
}

}  // namespace oracle
#endif  // ORACLE_DENSE_LINALG_H_
