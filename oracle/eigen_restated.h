// ORACLE — test infrastructure only (see jet.h header).
//
// Restatement of the Eigen 3.4 dense decompositions the reference's minimal solvers call:
//   Eigen::FullPivLU     (kernel(), solve())  five_point_relative_pose.cc:239-243, 261-263
//   Eigen::JacobiSVD 3x3 (full U, V)          essential_matrix_utils.cc:64-65
//   Eigen::EigenSolver   (real, N x N)        five_point_relative_pose.cc:275-277
// Eigen is an un-vendored, un-pinned dependency absent from /root/reference: these follow the
// published Eigen 3.4 algorithms (complete pivoting with first-maximum tie break in column-major scan
// order; two-sided Jacobi sweeps; Householder Hessenberg reduction + Francis double-shift QR with the
// eigenvalues read off the diagonal of T top to bottom; JAMA-style back-substitution for eigenvectors).
// PARITY UNPINNED against Eigen itself; they are checked against numpy (tests/test_oracle_ransac.py).
#ifndef ORACLE_EIGEN_RESTATED_H_
#define ORACLE_EIGEN_RESTATED_H_

#include <algorithm>
#include <cmath>
#include <limits>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// FullPivLU of an R x C matrix, row-major storage m[r*C + c].
template <int R, int C>
struct FullPivLU {
  static constexpr int SZ = R < C ? R : C;
  double lu[R * C];
  int row_tr[SZ], col_tr[SZ];
  int q[C];  // permutationQ().indices()
  int nonzero_pivots;
  double maxpivot;

  void compute(const double* a) {
    for (int i = 0; i < R * C; ++i) lu[i] = a[i];
    nonzero_pivots = SZ;
    maxpivot = 0.0;
    for (int k = 0; k < SZ; ++k) {
      // biggest |coeff| in the bottom-right corner; Eigen's visitor scans column by column, first max wins
      int br = k, bc = k;
      double biggest = -1.0;
      for (int c = k; c < C; ++c)
        for (int r = k; r < R; ++r) {
          const double v = std::fabs(lu[r * C + c]);
          if (v > biggest) { biggest = v; br = r; bc = c; }
        }
      if (biggest == 0.0) {
        nonzero_pivots = k;
        for (int i = k; i < SZ; ++i) { row_tr[i] = i; col_tr[i] = i; }
        break;
      }
      if (biggest > maxpivot) maxpivot = biggest;
      row_tr[k] = br; col_tr[k] = bc;
      if (k != br) for (int c = 0; c < C; ++c) std::swap(lu[k * C + c], lu[br * C + c]);
      if (k != bc) for (int r = 0; r < R; ++r) std::swap(lu[r * C + k], lu[r * C + bc]);
      if (k < R - 1) for (int r = k + 1; r < R; ++r) lu[r * C + k] /= lu[k * C + k];
      if (k < SZ - 1)
        for (int r = k + 1; r < R; ++r)
          for (int c = k + 1; c < C; ++c) lu[r * C + c] -= lu[r * C + k] * lu[k * C + c];
    }
    for (int i = 0; i < C; ++i) q[i] = i;
    for (int k = 0; k < SZ; ++k) std::swap(q[k], q[col_tr[k]]);
  }
  double threshold() const { return std::numeric_limits<double>::epsilon() * SZ; }
  int rank() const {
    const double pre = std::fabs(maxpivot) * threshold();
    int r = 0;
    for (int i = 0; i < nonzero_pivots; ++i) r += std::fabs(lu[i * C + i]) > pre;
    return r;
  }
  int dimensionOfKernel() const { return C - rank(); }

  // kernel(): C x dimker, row-major ker[r*dimker + k]. Only the generic case where the non-negligible
  // pivots are the leading ones is exercised by the 5-point solver; the general permutation is kept.
  void kernel(double* ker) const {
    const int rk = rank(), dimker = C - rk;
    if (dimker == 0) return;
    int pivots[SZ];
    const double pre = std::fabs(maxpivot) * threshold();
    int p = 0;
    for (int i = 0; i < nonzero_pivots; ++i) if (std::fabs(lu[i * C + i]) > pre) pivots[p++] = i;
    double m[SZ * C];
    for (int i = 0; i < rk; ++i) {
      for (int c = 0; c < C; ++c) m[i * C + c] = 0.0;
      for (int c = i; c < C; ++c) m[i * C + c] = lu[pivots[i] * C + c];
    }
    for (int i = 0; i < rk; ++i)
      for (int c = 0; c < i; ++c) m[i * C + c] = 0.0;
    for (int i = 0; i < rk; ++i)
      if (pivots[i] != i) for (int r = 0; r < rk; ++r) std::swap(m[r * C + i], m[r * C + pivots[i]]);
    // upper-triangular solve m[:, :rk] X = m[:, rk:rk+dimker], in place (back substitution)
    for (int k = 0; k < dimker; ++k)
      for (int i = rk - 1; i >= 0; --i) {
        double s = m[i * C + rk + k];
        for (int j = i + 1; j < rk; ++j) s -= m[i * C + j] * m[j * C + rk + k];
        m[i * C + rk + k] = s / m[i * C + i];
      }
    for (int i = rk - 1; i >= 0; --i)
      if (pivots[i] != i) for (int r = 0; r < rk; ++r) std::swap(m[r * C + i], m[r * C + pivots[i]]);
    for (int i = 0; i < rk; ++i)
      for (int k = 0; k < dimker; ++k) ker[q[i] * dimker + k] = -m[i * C + rk + k];
    for (int i = rk; i < C; ++i)
      for (int k = 0; k < dimker; ++k) ker[q[i] * dimker + k] = 0.0;
    for (int k = 0; k < dimker; ++k) ker[q[rk + k] * dimker + k] = 1.0;
  }

  // solve(): square case (R == C), NR right-hand sides, row-major rhs[r*NR + k] -> x[r*NR + k]
  template <int NR>
  void solve(const double* rhs, double* x) const {
    static_assert(R == C, "square only");
    const int rk = rank();
    if (rk == 0) { for (int i = 0; i < C * NR; ++i) x[i] = 0.0; return; }
    double c[R * NR];
    for (int i = 0; i < R * NR; ++i) c[i] = rhs[i];
    for (int k = 0; k < SZ; ++k)
      if (row_tr[k] != k) for (int j = 0; j < NR; ++j) std::swap(c[k * NR + j], c[row_tr[k] * NR + j]);
    for (int j = 0; j < NR; ++j) {
      for (int i = 0; i < R; ++i) {  // unit lower
        double s = c[i * NR + j];
        for (int t = 0; t < i; ++t) s -= lu[i * C + t] * c[t * NR + j];
        c[i * NR + j] = s;
      }
      for (int i = rk - 1; i >= 0; --i) {  // upper, leading rk x rk block
        double s = c[i * NR + j];
        for (int t = i + 1; t < rk; ++t) s -= lu[i * C + t] * c[t * NR + j];
        c[i * NR + j] = s / lu[i * C + i];
      }
    }
    for (int i = 0; i < rk; ++i) for (int j = 0; j < NR; ++j) x[q[i] * NR + j] = c[i * NR + j];
    for (int i = rk; i < C; ++i) for (int j = 0; j < NR; ++j) x[q[i] * NR + j] = 0.0;
  }
};

// ---------------------------------------------------------------------------------------------
// JacobiSVD of a 3x3 matrix (row-major), full U and V, singular values sorted descending.
struct JacobiRot { double c, s; };
inline bool MakeJacobi(double x, double y, double z, JacobiRot* j) {
  const double deno = 2.0 * std::fabs(y);
  if (deno < std::numeric_limits<double>::min()) { j->c = 1.0; j->s = 0.0; return false; }
  const double tau = (x - z) / deno;
  const double w = std::sqrt(tau * tau + 1.0);
  const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0.0 ? 1.0 : -1.0;
  const double n = 1.0 / std::sqrt(t * t + 1.0);
  j->s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  j->c = n;
  return true;
}
template <int N>
inline void RotLeftN(double* M, int p, int q, JacobiRot j) {  // rows p, q: x' = c x + s y, y' = -s x + c y
  for (int i = 0; i < N; ++i) {
    const double x = M[p * N + i], y = M[q * N + i];
    M[p * N + i] = j.c * x + j.s * y;
    M[q * N + i] = -j.s * x + j.c * y;
  }
}
template <int N>
inline void RotRightN(double* M, int p, int q, JacobiRot j) {  // cols p, q with j^T: x' = c x - s y, y' = s x + c y
  for (int i = 0; i < N; ++i) {
    const double x = M[i * N + p], y = M[i * N + q];
    M[i * N + p] = j.c * x - j.s * y;
    M[i * N + q] = j.s * x + j.c * y;
  }
}
// Square JacobiSVD (no QR preconditioner is applied by Eigen when rows == cols), row-major N x N.
template <int N>
inline void JacobiSVDN(const double* A, double* U, double* S, double* V) {
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  const double considerAsZero = std::numeric_limits<double>::min();
  double scale = 0.0;
  for (int i = 0; i < N * N; ++i) scale = std::max(scale, std::fabs(A[i]));
  if (scale == 0.0) scale = 1.0;
  double W[N * N];
  for (int i = 0; i < N * N; ++i) { W[i] = A[i] / scale; U[i] = V[i] = (i / N == i % N) ? 1.0 : 0.0; }
  double maxDiag = 0.0;
  for (int i = 0; i < N; ++i) maxDiag = std::max(maxDiag, std::fabs(W[i * N + i]));
  bool finished = false;
  int sweeps = 0;
  while (!finished && sweeps++ < 1000) {
    finished = true;
    for (int p = 1; p < N; ++p)
      for (int q = 0; q < p; ++q) {
        const double threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::fabs(W[p * N + q]) > threshold || std::fabs(W[q * N + p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd on [[W_pp, W_pq], [W_qp, W_qq]]
          const double m00 = W[p * N + p], m01 = W[p * N + q], m10 = W[q * N + p], m11 = W[q * N + q];
          JacobiRot rot1;
          const double t = m00 + m11, d = m10 - m01;
          if (std::fabs(d) < std::numeric_limits<double>::min()) { rot1.s = 0.0; rot1.c = 1.0; }
          else { const double u = t / d, tmp = std::sqrt(1.0 + u * u); rot1.s = 1.0 / tmp; rot1.c = u / tmp; }
          const double n00 = rot1.c * m00 + rot1.s * m10, n01 = rot1.c * m01 + rot1.s * m11;
          const double n11 = -rot1.s * m01 + rot1.c * m11;
          JacobiRot jr;
          MakeJacobi(n00, n01, n11, &jr);
          const JacobiRot jrt{jr.c, -jr.s};
          const JacobiRot jl{rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
          RotLeftN<N>(W, p, q, jl);
          RotRightN<N>(U, p, q, JacobiRot{jl.c, -jl.s});
          RotRightN<N>(W, p, q, jr);
          RotRightN<N>(V, p, q, jr);
          maxDiag = std::max(maxDiag, std::max(std::fabs(W[p * N + p]), std::fabs(W[q * N + q])));
        }
      }
  }
  for (int i = 0; i < N; ++i) {
    const double a = W[i * N + i];
    S[i] = std::fabs(a);
    if (a < 0.0) for (int r = 0; r < N; ++r) U[r * N + i] = -U[r * N + i];
  }
  for (int i = 0; i < N; ++i) S[i] *= scale;
  for (int i = 0; i < N; ++i) {
    int pos = i;
    double best = S[i];
    for (int k = i + 1; k < N; ++k) if (S[k] > best) { best = S[k]; pos = k; }
    if (best == 0.0) break;
    if (pos != i) {
      std::swap(S[i], S[pos]);
      for (int r = 0; r < N; ++r) { std::swap(U[r * N + i], U[r * N + pos]); std::swap(V[r * N + i], V[r * N + pos]); }
    }
  }
}
inline void JacobiSVD3(const double* A, double* U, double* S, double* V) { JacobiSVDN<3>(A, U, S, V); }

// ---------------------------------------------------------------------------------------------
// EigenSolver for a real N x N matrix (row-major). eig_re/eig_im: eigenvalues in Eigen's order; vec: for
// every REAL eigenvalue j, column j (vec[r*N + j]) is the unit-norm eigenvector; columns of complex pairs
// are left zero (the reference only uses real ones, five_point_relative_pose.cc:283-291).
template <int N>
struct EigenSolverReal {
  double T[N * N], Uq[N * N];
  double eig_re[N], eig_im[N];
  double vec[N * N];
  bool ok;

  static void MakeHouseholder(const double* v, int n, double* ess, double* tau, double* beta) {
    double tailSq = 0.0;
    for (int i = 1; i < n; ++i) tailSq += v[i] * v[i];
    const double c0 = v[0];
    if (tailSq <= std::numeric_limits<double>::min()) {
      *tau = 0.0; *beta = c0;
      for (int i = 0; i < n - 1; ++i) ess[i] = 0.0;
    } else {
      double b = std::sqrt(c0 * c0 + tailSq);
      if (c0 >= 0.0) b = -b;
      for (int i = 0; i < n - 1; ++i) ess[i] = v[1 + i] / (c0 - b);
      *tau = (b - c0) / b;
      *beta = b;
    }
  }
  // block(r0, c0, nr, nc).applyHouseholderOnTheLeft(ess (nr-1), tau)
  void HouseLeft(double* M, int r0, int c0, int nr, int nc, const double* ess, double tau) {
    if (nr == 1) { for (int c = 0; c < nc; ++c) M[r0 * N + c0 + c] *= 1.0 - tau; return; }
    if (tau == 0.0) return;
    for (int c = 0; c < nc; ++c) {
      double tmp = 0.0;
      for (int r = 1; r < nr; ++r) tmp += ess[r - 1] * M[(r0 + r) * N + c0 + c];
      tmp += M[r0 * N + c0 + c];
      M[r0 * N + c0 + c] -= tau * tmp;
      for (int r = 1; r < nr; ++r) M[(r0 + r) * N + c0 + c] -= tau * ess[r - 1] * tmp;
    }
  }
  void HouseRight(double* M, int r0, int c0, int nr, int nc, const double* ess, double tau) {
    if (nc == 1) { for (int r = 0; r < nr; ++r) M[(r0 + r) * N + c0] *= 1.0 - tau; return; }
    if (tau == 0.0) return;
    for (int r = 0; r < nr; ++r) {
      double tmp = 0.0;
      for (int c = 1; c < nc; ++c) tmp += M[(r0 + r) * N + c0 + c] * ess[c - 1];
      tmp += M[(r0 + r) * N + c0];
      M[(r0 + r) * N + c0] -= tau * tmp;
      for (int c = 1; c < nc; ++c) M[(r0 + r) * N + c0 + c] -= tau * tmp * ess[c - 1];
    }
  }

  void compute(const double* A, bool want_vectors = true) {
    ok = true;
    double scale = 0.0;
    for (int i = 0; i < N * N; ++i) scale = std::max(scale, std::fabs(A[i]));
    for (int i = 0; i < N * N; ++i) { vec[i] = 0.0; Uq[i] = (i / N == i % N) ? 1.0 : 0.0; }
    if (scale < std::numeric_limits<double>::min()) {
      for (int i = 0; i < N * N; ++i) T[i] = 0.0;
      for (int i = 0; i < N; ++i) { eig_re[i] = 0.0; eig_im[i] = 0.0; }
      return;
    }
    for (int i = 0; i < N * N; ++i) T[i] = A[i] / scale;
    // --- HessenbergDecomposition: Householder reflectors H_0 .. H_{N-3}; Q = H_0 H_1 ...
    double hco[N], hess[N][N];
    for (int i = 0; i < N - 1; ++i) {
      const int rem = N - i - 1;
      double v[N], ess[N], tau, beta;
      for (int r = 0; r < rem; ++r) v[r] = T[(i + 1 + r) * N + i];
      MakeHouseholder(v, rem, ess, &tau, &beta);
      T[(i + 1) * N + i] = beta;
      for (int r = 1; r < rem; ++r) T[(i + 1 + r) * N + i] = ess[r - 1];
      hco[i] = tau;
      for (int r = 0; r < rem - 1; ++r) hess[i][r] = ess[r];
      HouseLeft(T, i + 1, i + 1, rem, rem, ess, tau);
      HouseRight(T, 0, i + 1, N, rem, ess, tau);
    }
    // Q accumulated by applying the reflectors on the right of the identity in order (Q = H_0 H_1 ...)
    for (int i = 0; i < N - 1; ++i) {
      const int rem = N - i - 1;
      HouseRight(Uq, 0, i + 1, N, rem, hess[i], hco[i]);
    }
    // matrixH: zero below the first sub-diagonal
    for (int r = 2; r < N; ++r) for (int c = 0; c < r - 1; ++c) T[r * N + c] = 0.0;
    // --- RealSchur::computeFromHessenberg
    const double eps = std::numeric_limits<double>::epsilon();
    int iu = N - 1, iter = 0, totalIter = 0;
    const int maxIters = 40 * N;
    double exshift = 0.0, norm = 0.0;
    for (int j = 0; j < N; ++j) for (int r = 0; r < std::min(N, j + 2); ++r) norm += std::fabs(T[r * N + j]);
    const double considerAsZero = std::max(norm * eps * eps, std::numeric_limits<double>::min());
    if (norm != 0.0) {
      while (iu >= 0) {
        int il = iu;
        while (il > 0) {
          double s = std::fabs(T[(il - 1) * N + il - 1]) + std::fabs(T[il * N + il]);
          s = std::max(s * eps, considerAsZero);
          if (std::fabs(T[il * N + il - 1]) <= s) break;
          il--;
        }
        if (il == iu) {
          T[iu * N + iu] += exshift;
          if (iu > 0) T[iu * N + iu - 1] = 0.0;
          iu--; iter = 0;
        } else if (il == iu - 1) {
          SplitOffTwoRows(iu, exshift);
          iu -= 2; iter = 0;
        } else {
          double sh[3], v[3] = {0, 0, 0};
          ComputeShift(iu, iter, &exshift, sh);
          ++iter; ++totalIter;
          if (totalIter > maxIters) break;
          int im;
          for (im = iu - 2; im >= il; --im) {
            const double Tmm = T[im * N + im], r = sh[0] - Tmm, s = sh[1] - Tmm;
            v[0] = (r * s - sh[2]) / T[(im + 1) * N + im] + T[im * N + im + 1];
            v[1] = T[(im + 1) * N + im + 1] - Tmm - r - s;
            v[2] = T[(im + 2) * N + im + 1];
            if (im == il) break;
            const double lhs = T[im * N + im - 1] * (std::fabs(v[1]) + std::fabs(v[2]));
            const double rhs = v[0] * (std::fabs(T[(im - 1) * N + im - 1]) + std::fabs(Tmm) + std::fabs(T[(im + 1) * N + im + 1]));
            if (std::fabs(lhs) < eps * rhs) break;
          }
          FrancisStep(il, im, iu, v);
        }
      }
    }
    if (totalIter > maxIters) ok = false;
    for (int i = 0; i < N * N; ++i) T[i] *= scale;
    // --- eigenvalues from T (top to bottom)
    int i = 0;
    while (i < N) {
      if (i == N - 1 || T[(i + 1) * N + i] == 0.0) {
        eig_re[i] = T[i * N + i]; eig_im[i] = 0.0;
        if (!std::isfinite(eig_re[i])) { ok = false; return; }
        ++i;
      } else {
        const double p = 0.5 * (T[i * N + i] - T[(i + 1) * N + i + 1]);
        double t0 = T[(i + 1) * N + i], t1 = T[i * N + i + 1];
        const double maxval = std::max(std::fabs(p), std::max(std::fabs(t0), std::fabs(t1)));
        t0 /= maxval; t1 /= maxval;
        const double p0 = p / maxval;
        const double z = maxval * std::sqrt(std::fabs(p0 * p0 + t0 * t1));
        eig_re[i] = T[(i + 1) * N + i + 1] + p; eig_im[i] = z;
        eig_re[i + 1] = T[(i + 1) * N + i + 1] + p; eig_im[i + 1] = -z;
        if (!std::isfinite(eig_re[i]) || !std::isfinite(z)) { ok = false; return; }
        i += 2;
      }
    }
    if (want_vectors) ComputeRealEigenvectors();
  }

  void SplitOffTwoRows(int iu, double exshift) {
    const double p = 0.5 * (T[(iu - 1) * N + iu - 1] - T[iu * N + iu]);
    const double q = p * p + T[iu * N + iu - 1] * T[(iu - 1) * N + iu];
    T[iu * N + iu] += exshift;
    T[(iu - 1) * N + iu - 1] += exshift;
    if (q >= 0.0) {
      const double z = std::sqrt(std::fabs(q));
      double c, s;
      MakeGivens(p >= 0.0 ? p + z : p - z, T[iu * N + iu - 1], &c, &s);
      // rightCols(size-iu+1).applyOnTheLeft(iu-1, iu, rot.adjoint()): x' = c x - s y, y' = s x + c y
      for (int col = iu - 1; col < N; ++col) {
        const double x = T[(iu - 1) * N + col], y = T[iu * N + col];
        T[(iu - 1) * N + col] = c * x - s * y;
        T[iu * N + col] = s * x + c * y;
      }
      // topRows(iu+1).applyOnTheRight(iu-1, iu, rot): x' = c x - s y, y' = s x + c y
      for (int r = 0; r <= iu; ++r) {
        const double x = T[r * N + iu - 1], y = T[r * N + iu];
        T[r * N + iu - 1] = c * x - s * y;
        T[r * N + iu] = s * x + c * y;
      }
      T[iu * N + iu - 1] = 0.0;
      for (int r = 0; r < N; ++r) {
        const double x = Uq[r * N + iu - 1], y = Uq[r * N + iu];
        Uq[r * N + iu - 1] = c * x - s * y;
        Uq[r * N + iu] = s * x + c * y;
      }
    }
    if (iu > 1) T[(iu - 1) * N + iu - 2] = 0.0;
  }
  static void MakeGivens(double p, double q, double* c, double* s) {
    if (q == 0.0) { *c = p < 0.0 ? -1.0 : 1.0; *s = 0.0; }
    else if (p == 0.0) { *c = 0.0; *s = q < 0.0 ? 1.0 : -1.0; }
    else if (std::fabs(p) > std::fabs(q)) {
      const double t = q / p; double u = std::sqrt(1.0 + t * t); if (p < 0.0) u = -u;
      *c = 1.0 / u; *s = -t * (*c);
    } else {
      const double t = p / q; double u = std::sqrt(1.0 + t * t); if (q < 0.0) u = -u;
      *s = -1.0 / u; *c = -t * (*s);
    }
  }
  void ComputeShift(int iu, int iter, double* exshift, double* sh) {
    sh[0] = T[iu * N + iu]; sh[1] = T[(iu - 1) * N + iu - 1]; sh[2] = T[iu * N + iu - 1] * T[(iu - 1) * N + iu];
    if (iter == 10) {
      *exshift += sh[0];
      for (int i = 0; i <= iu; ++i) T[i * N + i] -= sh[0];
      const double s = std::fabs(T[iu * N + iu - 1]) + std::fabs(T[(iu - 1) * N + iu - 2]);
      sh[0] = 0.75 * s; sh[1] = 0.75 * s; sh[2] = -0.4375 * s * s;
    }
    if (iter == 30) {
      double s = (sh[1] - sh[0]) / 2.0;
      s = s * s + sh[2];
      if (s > 0.0) {
        s = std::sqrt(s);
        if (sh[1] < sh[0]) s = -s;
        s = s + (sh[1] - sh[0]) / 2.0;
        s = sh[0] - sh[2] / s;
        *exshift += s;
        for (int i = 0; i <= iu; ++i) T[i * N + i] -= s;
        sh[0] = sh[1] = sh[2] = 0.964;
      }
    }
  }
  void FrancisStep(int il, int im, int iu, const double* first) {
    for (int k = im; k <= iu - 2; ++k) {
      const bool firstIteration = k == im;
      double v[3];
      if (firstIteration) { v[0] = first[0]; v[1] = first[1]; v[2] = first[2]; }
      else { v[0] = T[k * N + k - 1]; v[1] = T[(k + 1) * N + k - 1]; v[2] = T[(k + 2) * N + k - 1]; }
      double ess[2], tau, beta;
      MakeHouseholder(v, 3, ess, &tau, &beta);
      if (beta != 0.0) {
        if (firstIteration && k > il) T[k * N + k - 1] = -T[k * N + k - 1];
        else if (!firstIteration) T[k * N + k - 1] = beta;
        HouseLeft(T, k, k, 3, N - k, ess, tau);
        HouseRight(T, 0, k, std::min(iu, k + 3) + 1, 3, ess, tau);
        HouseRight(Uq, 0, k, N, 3, ess, tau);
      }
    }
    double v[2] = {T[(iu - 1) * N + iu - 2], T[iu * N + iu - 2]};
    double ess[1], tau, beta;
    MakeHouseholder(v, 2, ess, &tau, &beta);
    if (beta != 0.0) {
      T[(iu - 1) * N + iu - 2] = beta;
      HouseLeft(T, iu - 1, iu - 1, 2, N - iu + 1, ess, tau);
      HouseRight(T, 0, iu - 1, iu + 1, 2, ess, tau);
      HouseRight(Uq, 0, iu - 1, N, 2, ess, tau);
    }
    for (int i = im + 2; i <= iu; ++i) {
      T[i * N + i - 2] = 0.0;
      if (i > im + 2) T[i * N + i - 3] = 0.0;
    }
  }
  // EigenSolver::doComputeEigenvectors restricted to real eigenvalues (the complex branch only writes
  // columns that no real eigenvector reads).
  void ComputeRealEigenvectors() {
    const double eps = std::numeric_limits<double>::epsilon();
    double norm = 0.0;
    for (int j = 0; j < N; ++j) for (int c = std::max(j - 1, 0); c < N; ++c) norm += std::fabs(T[j * N + c]);
    if (norm == 0.0) return;
    double M[N * N];
    for (int i = 0; i < N * N; ++i) M[i] = T[i];
    for (int n = N - 1; n >= 0; --n) {
      const double p = eig_re[n], q = eig_im[n];
      if (q != 0.0) continue;
      double lastr = 0.0, lastw = 0.0;
      int l = n;
      M[n * N + n] = 1.0;
      for (int i = n - 1; i >= 0; --i) {
        const double w = M[i * N + i] - p;
        double r = 0.0;
        for (int k = l; k <= n; ++k) r += M[i * N + k] * M[k * N + n];
        if (eig_im[i] < 0.0) { lastw = w; lastr = r; }
        else {
          l = i;
          if (eig_im[i] == 0.0) {
            if (w != 0.0) M[i * N + n] = -r / w; else M[i * N + n] = -r / (eps * norm);
          } else {
            const double x = M[i * N + i + 1], y = M[(i + 1) * N + i];
            const double denom = (eig_re[i] - p) * (eig_re[i] - p) + eig_im[i] * eig_im[i];
            const double t = (x * lastr - lastw * r) / denom;
            M[i * N + n] = t;
            if (std::fabs(x) > std::fabs(lastw)) M[(i + 1) * N + n] = (-r - w * t) / x;
            else M[(i + 1) * N + n] = (-lastr - y * t) / lastw;
          }
          const double t = std::fabs(M[i * N + n]);
          if ((eps * t) * t > 1.0) for (int k = i; k < N; ++k) M[k * N + n] /= t;
        }
      }
      // back transformation + normalisation
      double col[N], nrm = 0.0;
      for (int r = 0; r < N; ++r) {
        double s = 0.0;
        for (int k = 0; k <= n; ++k) s += Uq[r * N + k] * M[k * N + n];
        col[r] = s; nrm += s * s;
      }
      nrm = std::sqrt(nrm);
      for (int r = 0; r < N; ++r) vec[r * N + n] = col[r] / nrm;
    }
  }
};

}  // namespace oracle
#endif  // ORACLE_EIGEN_RESTATED_H_
