// Device-side small dense decompositions used by the minimal solvers (one thread = one problem).
//
// The reference calls Eigen::FullPivLU, Eigen::JacobiSVD and Eigen::EigenSolver inside its minimal solvers
// (sfm/pose/five_point_relative_pose.cc:239-277, sfm/pose/essential_matrix_utils.cc:64-65). RANSAC results
// depend on the ORDER in which those routines return solutions, so the same published algorithms are used
// here: complete pivoting with first-maximum tie break in column-major scan order, two-sided Jacobi sweeps,
// Householder Hessenberg reduction + Francis double-shift QR with eigenvalues read off the Schur form top to
// bottom, back-substitution for the real eigenvectors. This translation unit is compiled with -fmad=false:
// every +,-,*,/,sqrt is a separately rounded IEEE operation.
#ifndef THB_SMALL_LINALG_CUH_
#define THB_SMALL_LINALG_CUH_

#include <cuda_runtime.h>

namespace thb {
namespace sl {

constexpr double kEps = 2.220446049250313e-16;
constexpr double kMin = 2.2250738585072014e-308;

__device__ __forceinline__ void dswap(double& a, double& b) { const double t = a; a = b; b = t; }
__device__ __forceinline__ void iswap(int& a, int& b) { const int t = a; a = b; b = t; }

// Complete-pivoting LU of an R x C row-major matrix (in place).
template <int R, int C>
struct FullPivLU {
  static constexpr int SZ = R < C ? R : C;
  double* lu;  // caller-provided storage, R*C
  int row_tr[SZ], col_tr[SZ], q[C];
  int nonzero_pivots;
  double maxpivot;

  __device__ void compute() {
    nonzero_pivots = SZ;
    maxpivot = 0.0;
    for (int k = 0; k < SZ; ++k) {
      int br = k, bc = k;
      double biggest = -1.0;
      for (int c = k; c < C; ++c)
        for (int r = k; r < R; ++r) {
          const double v = fabs(lu[r * C + c]);
          if (v > biggest) { biggest = v; br = r; bc = c; }
        }
      if (biggest == 0.0) {
        nonzero_pivots = k;
        for (int i = k; i < SZ; ++i) { row_tr[i] = i; col_tr[i] = i; }
        break;
      }
      if (biggest > maxpivot) maxpivot = biggest;
      row_tr[k] = br; col_tr[k] = bc;
      if (k != br) for (int c = 0; c < C; ++c) dswap(lu[k * C + c], lu[br * C + c]);
      if (k != bc) for (int r = 0; r < R; ++r) dswap(lu[r * C + k], lu[r * C + bc]);
      if (k < R - 1) for (int r = k + 1; r < R; ++r) lu[r * C + k] /= lu[k * C + k];
      if (k < SZ - 1)
        for (int r = k + 1; r < R; ++r)
          for (int c = k + 1; c < C; ++c) lu[r * C + c] -= lu[r * C + k] * lu[k * C + c];
    }
    for (int i = 0; i < C; ++i) q[i] = i;
    for (int k = 0; k < SZ; ++k) iswap(q[k], q[col_tr[k]]);
  }
  __device__ int rank() const {
    const double pre = fabs(maxpivot) * (kEps * SZ);
    int r = 0;
    for (int i = 0; i < nonzero_pivots; ++i) r += fabs(lu[i * C + i]) > pre;
    return r;
  }
};

// kernel of a full-row-rank R x 9 matrix: ker[9][9-R] (row-major). d holds the computed decomposition.
template <int R>
__device__ inline void kernel_rx9(const FullPivLU<R, 9>& d, double* ker) {
  constexpr int C = 9, rk = R, dimker = 9 - R;
  double m[rk * C];
  for (int i = 0; i < rk; ++i) {
    for (int c = 0; c < i; ++c) m[i * C + c] = 0.0;
    for (int c = i; c < C; ++c) m[i * C + c] = d.lu[i * C + c];
  }
  for (int k = 0; k < dimker; ++k)
    for (int i = rk - 1; i >= 0; --i) {
      double s = m[i * C + rk + k];
      for (int j = i + 1; j < rk; ++j) s -= m[i * C + j] * m[j * C + rk + k];
      m[i * C + rk + k] = s / m[i * C + i];
    }
  for (int i = 0; i < rk; ++i)
    for (int k = 0; k < dimker; ++k) ker[d.q[i] * dimker + k] = -m[i * C + rk + k];
  for (int i = rk; i < C; ++i)
    for (int k = 0; k < dimker; ++k) ker[d.q[i] * dimker + k] = 0.0;
  for (int k = 0; k < dimker; ++k) ker[d.q[rk + k] * dimker + k] = 1.0;
}

// x = A^-1 rhs for the 10x10 decomposition with 10 right-hand sides (row-major), rank-revealing like Eigen.
__device__ inline void solve_10x10(const FullPivLU<10, 10>& d, double* c /* in: rhs, scratch */, double* x) {
  constexpr int N = 10;
  const int rk = d.rank();
  if (rk == 0) { for (int i = 0; i < N * N; ++i) x[i] = 0.0; return; }
  for (int k = 0; k < N; ++k)
    if (d.row_tr[k] != k) for (int j = 0; j < N; ++j) dswap(c[k * N + j], c[d.row_tr[k] * N + j]);
  for (int j = 0; j < N; ++j) {
    for (int i = 0; i < N; ++i) {
      double s = c[i * N + j];
      for (int t = 0; t < i; ++t) s -= d.lu[i * N + t] * c[t * N + j];
      c[i * N + j] = s;
    }
    for (int i = rk - 1; i >= 0; --i) {
      double s = c[i * N + j];
      for (int t = i + 1; t < rk; ++t) s -= d.lu[i * N + t] * c[t * N + j];
      c[i * N + j] = s / d.lu[i * N + i];
    }
  }
  for (int i = 0; i < rk; ++i) for (int j = 0; j < N; ++j) x[d.q[i] * N + j] = c[i * N + j];
  for (int i = rk; i < N; ++i) for (int j = 0; j < N; ++j) x[d.q[i] * N + j] = 0.0;
}

// ---- two-sided Jacobi SVD of a 3x3 (row-major), full U and V, singular values sorted descending ----
struct Rot { double c, s; };
__device__ __forceinline__ void make_jacobi(double x, double y, double z, Rot* j) {
  const double deno = 2.0 * fabs(y);
  if (deno < kMin) { j->c = 1.0; j->s = 0.0; return; }
  const double tau = (x - z) / deno;
  const double w = sqrt(tau * tau + 1.0);
  const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0.0 ? 1.0 : -1.0;
  const double n = 1.0 / sqrt(t * t + 1.0);
  j->s = -sign_t * (y / fabs(y)) * fabs(t) * n;
  j->c = n;
}
template <int N>
__device__ __forceinline__ void rot_left(double* M, int p, int q, Rot j) {
  for (int i = 0; i < N; ++i) {
    const double x = M[p * N + i], y = M[q * N + i];
    M[p * N + i] = j.c * x + j.s * y;
    M[q * N + i] = -j.s * x + j.c * y;
  }
}
template <int N>
__device__ __forceinline__ void rot_right(double* M, int p, int q, Rot j) {
  for (int i = 0; i < N; ++i) {
    const double x = M[i * N + p], y = M[i * N + q];
    M[i * N + p] = j.c * x - j.s * y;
    M[i * N + q] = j.s * x + j.c * y;
  }
}
// Square two-sided Jacobi SVD, row-major N x N; W is caller-provided N*N scratch (may alias nothing else).
// WANT_U = false skips the accumulation of U (Eigen's ComputeFullV-only decomposition): W, S and V are unaffected, U may be nullptr.
template <int N, bool WANT_U = true>
__device__ inline void jacobi_svd(const double* A, double* W, double* U, double* S, double* V) {
  const double precision = 2.0 * kEps;
  double scale = 0.0;
  for (int i = 0; i < N * N; ++i) scale = fmax(scale, fabs(A[i]));
  if (scale == 0.0) scale = 1.0;
  for (int i = 0; i < N * N; ++i) { W[i] = A[i] / scale; V[i] = (i / N == i % N) ? 1.0 : 0.0; if (WANT_U) U[i] = V[i]; }
  double maxDiag = 0.0;
  for (int i = 0; i < N; ++i) maxDiag = fmax(maxDiag, fabs(W[i * N + i]));
  bool finished = false;
  int sweeps = 0;
  while (!finished && sweeps++ < 1000) {
    finished = true;
    for (int p = 1; p < N; ++p)
      for (int q = 0; q < p; ++q) {
        const double threshold = fmax(kMin, precision * maxDiag);
        if (fabs(W[p * N + q]) > threshold || fabs(W[q * N + p]) > threshold) {
          finished = false;
          const double m00 = W[p * N + p], m01 = W[p * N + q], m10 = W[q * N + p], m11 = W[q * N + q];
          Rot rot1;
          const double t = m00 + m11, d = m10 - m01;
          if (fabs(d) < kMin) { rot1.s = 0.0; rot1.c = 1.0; }
          else { const double u = t / d, tmp = sqrt(1.0 + u * u); rot1.s = 1.0 / tmp; rot1.c = u / tmp; }
          const double n00 = rot1.c * m00 + rot1.s * m10, n01 = rot1.c * m01 + rot1.s * m11;
          const double n11 = -rot1.s * m01 + rot1.c * m11;
          Rot jr;
          make_jacobi(n00, n01, n11, &jr);
          const Rot jrt{jr.c, -jr.s};
          const Rot jl{rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
          rot_left<N>(W, p, q, jl);
          if (WANT_U) rot_right<N>(U, p, q, Rot{jl.c, -jl.s});
          rot_right<N>(W, p, q, jr);
          rot_right<N>(V, p, q, jr);
          maxDiag = fmax(maxDiag, fmax(fabs(W[p * N + p]), fabs(W[q * N + q])));
        }
      }
  }
  for (int i = 0; i < N; ++i) {
    const double a = W[i * N + i];
    S[i] = fabs(a);
    if (WANT_U && a < 0.0) for (int r = 0; r < N; ++r) U[r * N + i] = -U[r * N + i];
  }
  for (int i = 0; i < N; ++i) S[i] *= scale;
  for (int i = 0; i < N; ++i) {
    int pos = i;
    double best = S[i];
    for (int k = i + 1; k < N; ++k) if (S[k] > best) { best = S[k]; pos = k; }
    if (best == 0.0) break;
    if (pos != i) {
      dswap(S[i], S[pos]);
      for (int r = 0; r < N; ++r) { if (WANT_U) dswap(U[r * N + i], U[r * N + pos]); dswap(V[r * N + i], V[r * N + pos]); }
    }
  }
}
__device__ inline void jacobi_svd3(const double* A, double* U, double* S, double* V) {
  double W[9];
  jacobi_svd<3>(A, W, U, S, V);
}

// ---- real nonsymmetric eigen-decomposition, N x N row-major ---------------------------------------
// T, Uq, M: caller-provided N*N scratch. On return eig_re/eig_im hold the eigenvalues in Schur-form order and,
// for every real eigenvalue j, vec_tail4[j][0..3] = the last four components of its unit-norm eigenvector
// (all the five-point solver needs, five_point_relative_pose.cc:288-290).
template <int N>
struct EigenReal {
  double *T, *Uq, *M;
  double eig_re[N], eig_im[N];
  bool ok;

  __device__ static void householder(const double* v, int n, double* ess, double* tau, double* beta) {
    double tailSq = 0.0;
    for (int i = 1; i < n; ++i) tailSq += v[i] * v[i];
    const double c0 = v[0];
    if (tailSq <= kMin) {
      *tau = 0.0; *beta = c0;
      for (int i = 0; i < n - 1; ++i) ess[i] = 0.0;
    } else {
      double b = sqrt(c0 * c0 + tailSq);
      if (c0 >= 0.0) b = -b;
      for (int i = 0; i < n - 1; ++i) ess[i] = v[1 + i] / (c0 - b);
      *tau = (b - c0) / b;
      *beta = b;
    }
  }
  __device__ static void house_left(double* X, int r0, int c0, int nr, int nc, const double* ess, double tau) {
    if (nr == 1) { for (int c = 0; c < nc; ++c) X[r0 * N + c0 + c] *= 1.0 - tau; return; }
    if (tau == 0.0) return;
    for (int c = 0; c < nc; ++c) {
      double tmp = 0.0;
      for (int r = 1; r < nr; ++r) tmp += ess[r - 1] * X[(r0 + r) * N + c0 + c];
      tmp += X[r0 * N + c0 + c];
      X[r0 * N + c0 + c] -= tau * tmp;
      for (int r = 1; r < nr; ++r) X[(r0 + r) * N + c0 + c] -= tau * ess[r - 1] * tmp;
    }
  }
  __device__ static void house_right(double* X, int r0, int c0, int nr, int nc, const double* ess, double tau) {
    if (nc == 1) { for (int r = 0; r < nr; ++r) X[(r0 + r) * N + c0] *= 1.0 - tau; return; }
    if (tau == 0.0) return;
    for (int r = 0; r < nr; ++r) {
      double tmp = 0.0;
      for (int c = 1; c < nc; ++c) tmp += X[(r0 + r) * N + c0 + c] * ess[c - 1];
      tmp += X[(r0 + r) * N + c0];
      X[(r0 + r) * N + c0] -= tau * tmp;
      for (int c = 1; c < nc; ++c) X[(r0 + r) * N + c0 + c] -= tau * tmp * ess[c - 1];
    }
  }
  __device__ static void givens(double p, double q, double* c, double* s) {
    if (q == 0.0) { *c = p < 0.0 ? -1.0 : 1.0; *s = 0.0; }
    else if (p == 0.0) { *c = 0.0; *s = q < 0.0 ? 1.0 : -1.0; }
    else if (fabs(p) > fabs(q)) {
      const double t = q / p; double u = sqrt(1.0 + t * t); if (p < 0.0) u = -u;
      *c = 1.0 / u; *s = -t * (*c);
    } else {
      const double t = p / q; double u = sqrt(1.0 + t * t); if (q < 0.0) u = -u;
      *s = -1.0 / u; *c = -t * (*s);
    }
  }

  // A is read from T (caller fills T with the matrix).
  __device__ void compute(double (*vec_tail4)[4], bool want_vectors = true) {
    ok = true;
    double scale = 0.0;
    for (int i = 0; i < N * N; ++i) scale = fmax(scale, fabs(T[i]));
    for (int i = 0; i < N * N; ++i) Uq[i] = (i / N == i % N) ? 1.0 : 0.0;
    if (want_vectors) for (int j = 0; j < N; ++j) for (int k = 0; k < 4; ++k) vec_tail4[j][k] = 0.0;
    if (scale < kMin) {
      for (int i = 0; i < N; ++i) { eig_re[i] = 0.0; eig_im[i] = 0.0; }
      return;
    }
    for (int i = 0; i < N * N; ++i) T[i] = T[i] / scale;
    // Hessenberg reduction; the reflectors are kept in M (row i: essential part) to accumulate Q afterwards
    double hco[N];
    for (int i = 0; i < N - 1; ++i) {
      const int rem = N - i - 1;
      double v[N], ess[N], tau, beta;
      for (int r = 0; r < rem; ++r) v[r] = T[(i + 1 + r) * N + i];
      householder(v, rem, ess, &tau, &beta);
      T[(i + 1) * N + i] = beta;
      for (int r = 1; r < rem; ++r) T[(i + 1 + r) * N + i] = ess[r - 1];
      hco[i] = tau;
      for (int r = 0; r < rem - 1; ++r) M[i * N + r] = ess[r];
      house_left(T, i + 1, i + 1, rem, rem, ess, tau);
      house_right(T, 0, i + 1, N, rem, ess, tau);
    }
    for (int i = 0; i < N - 1; ++i) house_right(Uq, 0, i + 1, N, N - i - 1, M + i * N, hco[i]);
    for (int r = 2; r < N; ++r) for (int c = 0; c < r - 1; ++c) T[r * N + c] = 0.0;
    // Francis double-shift QR
    int iu = N - 1, iter = 0, totalIter = 0;
    const int maxIters = 40 * N;
    double exshift = 0.0, norm = 0.0;
    for (int j = 0; j < N; ++j) for (int r = 0; r < (N < j + 2 ? N : j + 2); ++r) norm += fabs(T[r * N + j]);
    const double considerAsZero = fmax(norm * kEps * kEps, kMin);
    if (norm != 0.0) {
      while (iu >= 0) {
        int il = iu;
        while (il > 0) {
          double s = fabs(T[(il - 1) * N + il - 1]) + fabs(T[il * N + il]);
          s = fmax(s * kEps, considerAsZero);
          if (fabs(T[il * N + il - 1]) <= s) break;
          il--;
        }
        if (il == iu) {
          T[iu * N + iu] += exshift;
          if (iu > 0) T[iu * N + iu - 1] = 0.0;
          iu--; iter = 0;
        } else if (il == iu - 1) {
          split_off_two_rows(iu, exshift);
          iu -= 2; iter = 0;
        } else {
          double sh[3], v[3] = {0.0, 0.0, 0.0};
          compute_shift(iu, iter, &exshift, sh);
          ++iter; ++totalIter;
          if (totalIter > maxIters) break;
          int im;
          for (im = iu - 2; im >= il; --im) {
            const double Tmm = T[im * N + im], r = sh[0] - Tmm, s = sh[1] - Tmm;
            v[0] = (r * s - sh[2]) / T[(im + 1) * N + im] + T[im * N + im + 1];
            v[1] = T[(im + 1) * N + im + 1] - Tmm - r - s;
            v[2] = T[(im + 2) * N + im + 1];
            if (im == il) break;
            const double lhs = T[im * N + im - 1] * (fabs(v[1]) + fabs(v[2]));
            const double rhs = v[0] * (fabs(T[(im - 1) * N + im - 1]) + fabs(Tmm) + fabs(T[(im + 1) * N + im + 1]));
            if (fabs(lhs) < kEps * rhs) break;
          }
          francis_step(il, im, iu, v);
        }
      }
    }
    if (totalIter > maxIters) ok = false;
    for (int i = 0; i < N * N; ++i) T[i] *= scale;
    int i = 0;
    while (i < N) {
      if (i == N - 1 || T[(i + 1) * N + i] == 0.0) {
        eig_re[i] = T[i * N + i]; eig_im[i] = 0.0;
        if (!isfinite(eig_re[i])) { ok = false; return; }
        ++i;
      } else {
        const double p = 0.5 * (T[i * N + i] - T[(i + 1) * N + i + 1]);
        double t0 = T[(i + 1) * N + i], t1 = T[i * N + i + 1];
        const double maxval = fmax(fabs(p), fmax(fabs(t0), fabs(t1)));
        t0 /= maxval; t1 /= maxval;
        const double p0 = p / maxval;
        const double z = maxval * sqrt(fabs(p0 * p0 + t0 * t1));
        eig_re[i] = T[(i + 1) * N + i + 1] + p; eig_im[i] = z;
        eig_re[i + 1] = T[(i + 1) * N + i + 1] + p; eig_im[i + 1] = -z;
        if (!isfinite(eig_re[i]) || !isfinite(z)) { ok = false; return; }
        i += 2;
      }
    }
    if (want_vectors) real_eigenvectors(vec_tail4);
  }

  __device__ void split_off_two_rows(int iu, double exshift) {
    const double p = 0.5 * (T[(iu - 1) * N + iu - 1] - T[iu * N + iu]);
    const double q = p * p + T[iu * N + iu - 1] * T[(iu - 1) * N + iu];
    T[iu * N + iu] += exshift;
    T[(iu - 1) * N + iu - 1] += exshift;
    if (q >= 0.0) {
      const double z = sqrt(fabs(q));
      double c, s;
      givens(p >= 0.0 ? p + z : p - z, T[iu * N + iu - 1], &c, &s);
      for (int col = iu - 1; col < N; ++col) {
        const double x = T[(iu - 1) * N + col], y = T[iu * N + col];
        T[(iu - 1) * N + col] = c * x - s * y;
        T[iu * N + col] = s * x + c * y;
      }
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
  __device__ void compute_shift(int iu, int iter, double* exshift, double* sh) {
    sh[0] = T[iu * N + iu]; sh[1] = T[(iu - 1) * N + iu - 1]; sh[2] = T[iu * N + iu - 1] * T[(iu - 1) * N + iu];
    if (iter == 10) {
      *exshift += sh[0];
      for (int i = 0; i <= iu; ++i) T[i * N + i] -= sh[0];
      const double s = fabs(T[iu * N + iu - 1]) + fabs(T[(iu - 1) * N + iu - 2]);
      sh[0] = 0.75 * s; sh[1] = 0.75 * s; sh[2] = -0.4375 * s * s;
    }
    if (iter == 30) {
      double s = (sh[1] - sh[0]) / 2.0;
      s = s * s + sh[2];
      if (s > 0.0) {
        s = sqrt(s);
        if (sh[1] < sh[0]) s = -s;
        s = s + (sh[1] - sh[0]) / 2.0;
        s = sh[0] - sh[2] / s;
        *exshift += s;
        for (int i = 0; i <= iu; ++i) T[i * N + i] -= s;
        sh[0] = sh[1] = sh[2] = 0.964;
      }
    }
  }
  __device__ void francis_step(int il, int im, int iu, const double* first) {
    for (int k = im; k <= iu - 2; ++k) {
      const bool firstIteration = k == im;
      double v[3];
      if (firstIteration) { v[0] = first[0]; v[1] = first[1]; v[2] = first[2]; }
      else { v[0] = T[k * N + k - 1]; v[1] = T[(k + 1) * N + k - 1]; v[2] = T[(k + 2) * N + k - 1]; }
      double ess[2], tau, beta;
      householder(v, 3, ess, &tau, &beta);
      if (beta != 0.0) {
        if (firstIteration && k > il) T[k * N + k - 1] = -T[k * N + k - 1];
        else if (!firstIteration) T[k * N + k - 1] = beta;
        house_left(T, k, k, 3, N - k, ess, tau);
        house_right(T, 0, k, (iu < k + 3 ? iu : k + 3) + 1, 3, ess, tau);
        house_right(Uq, 0, k, N, 3, ess, tau);
      }
    }
    double v[2] = {T[(iu - 1) * N + iu - 2], T[iu * N + iu - 2]};
    double ess[1], tau, beta;
    householder(v, 2, ess, &tau, &beta);
    if (beta != 0.0) {
      T[(iu - 1) * N + iu - 2] = beta;
      house_left(T, iu - 1, iu - 1, 2, N - iu + 1, ess, tau);
      house_right(T, 0, iu - 1, iu + 1, 2, ess, tau);
      house_right(Uq, 0, iu - 1, N, 2, ess, tau);
    }
    for (int i = im + 2; i <= iu; ++i) {
      T[i * N + i - 2] = 0.0;
      if (i > im + 2) T[i * N + i - 3] = 0.0;
    }
  }
  __device__ void real_eigenvectors(double (*vec_tail4)[4]) {
    double norm = 0.0;
    for (int j = 0; j < N; ++j) for (int c = (j - 1 > 0 ? j - 1 : 0); c < N; ++c) norm += fabs(T[j * N + c]);
    if (norm == 0.0) return;
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
            if (w != 0.0) M[i * N + n] = -r / w; else M[i * N + n] = -r / (kEps * norm);
          } else {
            const double x = M[i * N + i + 1], y = M[(i + 1) * N + i];
            const double denom = (eig_re[i] - p) * (eig_re[i] - p) + eig_im[i] * eig_im[i];
            const double t = (x * lastr - lastw * r) / denom;
            M[i * N + n] = t;
            if (fabs(x) > fabs(lastw)) M[(i + 1) * N + n] = (-r - w * t) / x;
            else M[(i + 1) * N + n] = (-lastr - y * t) / lastw;
          }
          const double t = fabs(M[i * N + n]);
          if ((kEps * t) * t > 1.0) for (int k = i; k < N; ++k) M[k * N + n] /= t;
        }
      }
      double col[N], nrm = 0.0;
      for (int r = 0; r < N; ++r) {
        double s = 0.0;
        for (int k = 0; k <= n; ++k) s += Uq[r * N + k] * M[k * N + n];
        col[r] = s; nrm += s * s;
      }
      nrm = sqrt(nrm);
      for (int k = 0; k < 4; ++k) vec_tail4[n][k] = col[(N >= 4 ? N - 4 : 0) + (N >= 4 ? k : 0)] / nrm;
    }
  }
};

}  // namespace sl
}  // namespace thb
#endif  // THB_SMALL_LINALG_CUH_
