// Hot path 2: RANSAC two-view verification, whole loop on the device. Two schedules of the same loop, bit-identical results:
// k_ransac (one persistent CTA per image pair, all phases fused; LO-RANSAC; large batches) and k_rs_* (one kernel per phase over
// all active pairs, batches up to 2 048 pairs). Estimators: relative pose (five-point), absolute pose (P3P), homography (4-point);
// samplers: RandomSampler (RANSAC, LMED) and ProsacSampler (PROSAC); quality measurements: inlier support, MLE, LMED (median). two_view.cuh builds EstimateTwoViewInfo / VerifyMatches on top.
//
// Stands behind theia::EstimateRelativePose (sfm/estimators/estimate_relative_pose.cc:159-172) =
// SampleConsensusEstimator<RelativePoseEstimator>::Estimate (solvers/sample_consensus_estimator.h:299-415)
// with RandomSampler (solvers/random_sampler.cc:53-72) on std::mt19937 (util/random.cc:46-84).
//
// The reference loop is sequential (persistent sampler permutation, strict-< best update in model order,
// adaptive iteration bound). It is replayed exactly: iterations are processed in batches of BI = 128 —
//   draw   : one thread advances a bit-exact mt19937 + libstdc++ uniform_int_distribution (Lemire) and the
//            partial Fisher-Yates permutation, 5 indices per iteration
//   solve  : one THREAD per hypothesis runs the five-point solver and the essential-matrix decomposition + cheirality
//            vote (~1.5 M cycles of dependent FP64 / local-memory latency each, profiles/r02_ransac_schedule_ab.txt); the
//            candidate models go to a scratch area in global memory (L2)
//   score  : one WARP per model scores every correspondence (cheirality-gated Sampson), read through L1
//            (a pair's 64 KB stays cached between models), lane-strided, warp-shuffle reduction; a model
//            whose partial cost already reaches the best cost known at batch start is abandoned (it can
//            never win the strict-< test: exactness is preserved)
//   scan   : one thread walks the (iteration, model) costs in order, updates the best model and the adaptive
//            bound, and discards everything past the terminating iteration
// This file is compiled with -fmad=false (see small_linalg.cuh); the scoring formulas use explicit fma() in a
// fixed order so that they are both fast and bit-reproducible against the CPU oracle.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "small_linalg.cuh"
#include "ba_device.cuh"

namespace thb {
namespace {

constexpr int RT = 128;   // threads per CTA (three CTAs per SM: the relative-pose solver wants ~170 registers)
constexpr int NW = RT / 32;
constexpr int BI = 128;   // iterations per batch = hypotheses solved in parallel (one per thread)
constexpr int MAXM = 10;  // five-point solutions per sample
#ifndef RANSAC_CTAS_PER_SM
#define RANSAC_CTAS_PER_SM 3
#endif

struct Model { double E[9], R[9], p[3]; };

__constant__ int kT11[4][4] = {{0, 1, 3, 6}, {1, 2, 4, 7}, {3, 4, 5, 8}, {6, 7, 8, 9}};
__constant__ int kT21[10][4] = {{0, 1, 4, 10}, {1, 2, 5, 11}, {2, 3, 6, 12}, {4, 5, 7, 13}, {5, 6, 8, 14},
                                {7, 8, 9, 15}, {10, 11, 13, 16}, {11, 12, 14, 17}, {13, 14, 15, 18}, {16, 17, 18, 19}};

__device__ __forceinline__ void mul11(const double* a, const double* b, double* out) {
  for (int k = 0; k < 10; ++k) out[k] = 0.0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[kT11[i][j]] += a[i] * b[j];
}
__device__ __forceinline__ void mul21(const double* a, const double* b, double* out) {
  for (int k = 0; k < 20; ++k) out[k] = 0.0;
  for (int i = 0; i < 10; ++i) for (int j = 0; j < 4; ++j) out[kT21[i][j]] += a[i] * b[j];
}

// theia::FivePointRelativePose, minimal case (five_point_relative_pose.cc:212-293). E_out: up to 10 row-major 3x3.
__device__ int five_point(const double* x1, const double* x2, double* E_out) {
  double ns[9 * 4];
  {
    double epi[5 * 9];
    for (int i = 0; i < 5; ++i) {
      const double ax = x1[2 * i], ay = x1[2 * i + 1], bx = x2[2 * i], by = x2[2 * i + 1];
      double* r = epi + 9 * i;
      r[0] = bx * ax; r[1] = by * ax; r[2] = ax; r[3] = bx * ay; r[4] = by * ay; r[5] = ay; r[6] = bx; r[7] = by; r[8] = 1.0;
    }
    sl::FullPivLU<5, 9> lu;
    lu.lu = epi;
    lu.compute();
    if (9 - lu.rank() != 4) return 0;
    sl::kernel_rx9<5>(lu, ns);
  }
  const double* E[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) E[i][j] = ns + 4 * (i + 3 * j);
  double C[10 * 20];
  {
    double eet[3][3][10], tr[10], tmp[10], t20[20];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < 10; ++k) eet[i][j][k] = 0.0;
        for (int m = 0; m < 3; ++m) { mul11(E[i][m], E[j][m], tmp); for (int k = 0; k < 10; ++k) eet[i][j][k] += tmp[k]; }
        for (int k = 0; k < 10; ++k) eet[i][j][k] *= 2.0;
      }
    for (int k = 0; k < 10; ++k) tr[k] = eet[0][0][k] + eet[1][1][k] + eet[2][2][k];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double* row = C + 20 * (3 * i + j);
        for (int k = 0; k < 20; ++k) row[k] = 0.0;
        for (int m = 0; m < 3; ++m) { mul21(eet[i][m], E[m][j], t20); for (int k = 0; k < 20; ++k) row[k] += t20[k]; }
        mul21(tr, E[i][j], t20);
        for (int k = 0; k < 20; ++k) row[k] -= 0.5 * t20[k];
      }
    double a[10], b[10], d[10];
    double* row = C + 20 * 9;
    for (int k = 0; k < 20; ++k) row[k] = 0.0;
    const int idx[3][4][2] = {{{0, 1}, {1, 2}, {0, 2}, {1, 1}}, {{0, 2}, {1, 0}, {0, 0}, {1, 2}}, {{0, 0}, {1, 1}, {0, 1}, {1, 0}}};
    for (int c = 0; c < 3; ++c) {
      mul11(E[idx[c][0][0]][idx[c][0][1]], E[idx[c][1][0]][idx[c][1][1]], a);
      mul11(E[idx[c][2][0]][idx[c][2][1]], E[idx[c][3][0]][idx[c][3][1]], b);
      for (int k = 0; k < 10; ++k) d[k] = a[k] - b[k];
      mul21(d, E[2][c], t20);
      for (int k = 0; k < 20; ++k) row[k] += t20[k];
    }
  }
  // per-thread work matrices live in local memory and their footprint decides the L1 hit rate of the solve phase:
  // X reuses C, which is dead between its split into L | Rm and its reuse as eigen-solver scratch
  double L[100], Rm[100];
  double* X = C;
  for (int r = 0; r < 10; ++r) for (int c = 0; c < 10; ++c) { L[r * 10 + c] = C[r * 20 + c]; Rm[r * 10 + c] = C[r * 20 + 10 + c]; }
  {
    sl::FullPivLU<10, 10> clu;
    clu.lu = L;
    clu.compute();
    sl::solve_10x10(clu, Rm, X);
  }
  // action matrix in L (reused as T), scratch Rm (Uq) and C (M)
  for (int i = 0; i < 100; ++i) L[i] = 0.0;
  const int src[6] = {0, 1, 2, 4, 5, 7};
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 10; ++c) L[r * 10 + c] = X[src[r] * 10 + c];
  L[6 * 10 + 0] = -1.0; L[7 * 10 + 1] = -1.0; L[8 * 10 + 3] = -1.0; L[9 * 10 + 6] = -1.0;
  sl::EigenReal<10> es;
  es.T = L; es.Uq = Rm; es.M = C;
  double tail[10][4];
  es.compute(tail);
  int n = 0;
  for (int i = 0; i < 10; ++i) {
    if (es.eig_im[i] != 0.0) continue;
    double e9[9];
    for (int r = 0; r < 9; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += ns[r * 4 + k] * tail[i][k];
      e9[r] = s;
    }
    double* Eo = E_out + 9 * n;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Eo[r * 3 + c] = e9[c * 3 + r];
    ++n;
  }
  return n;
}

__device__ __forceinline__ double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// theia::IsTriangulatedPointInFrontOfCameras (sfm/triangulation/triangulation.cc:216-232); explicit fma order
// shared with the oracle.
__device__ __forceinline__ bool in_front(double x1, double y1, double x2, double y2, const double* R, const double* pos) {
  const double d2x = fma(R[0], x2, fma(R[3], y2, R[6]));
  const double d2y = fma(R[1], x2, fma(R[4], y2, R[7]));
  const double d2z = fma(R[2], x2, fma(R[5], y2, R[8]));
  const double dir1_sq = fma(x1, x1, fma(y1, y1, 1.0));
  const double dir2_sq = fma(d2x, d2x, fma(d2y, d2y, d2z * d2z));
  const double dir1_dir2 = fma(x1, d2x, fma(y1, d2y, d2z));
  const double dir1_pos = fma(x1, pos[0], fma(y1, pos[1], pos[2]));
  const double dir2_pos = fma(d2x, pos[0], fma(d2y, pos[1], d2z * pos[2]));
  return fma(dir2_sq, dir1_pos, -(dir1_dir2 * dir2_pos)) > 0.0 && fma(dir1_dir2, dir1_pos, -(dir1_sq * dir2_pos)) > 0.0;
}

// theia::SquaredSampsonDistance (sfm/pose/util.cc:56-69)
__device__ __forceinline__ double sampson(const double* F, double x0, double x1, double y0, double y1) {
  const double ex0 = fma(F[0], x0, fma(F[1], x1, F[2]));
  const double ex1 = fma(F[3], x0, fma(F[4], x1, F[5]));
  const double ex2 = fma(F[6], x0, fma(F[7], x1, F[8]));
  const double num = fma(y0, ex0, fma(y1, ex1, ex2));
  const double dy0 = fma(y0, F[0], fma(y1, F[3], F[6]));
  const double dy1 = fma(y0, F[1], fma(y1, F[4], F[7]));
  const double den = fma(dy0, dy0, fma(dy1, dy1, fma(ex0, ex0, ex1 * ex1)));
  return num * num / den;
}

// theia::DecomposeEssentialMatrix + GetBestPoseFromEssentialMatrix (sfm/pose/essential_matrix_utils.cc:57-80,
// 109-149) over the sample's n correspondences (x1,y1,x2,y2 quadruples).
__device__ int best_pose(const double* E, const double* corr, int n, double* Rbest, double* pbest) {
  double U[9], S[3], V[9];
  sl::jacobi_svd3(E, U, S, V);
  if (det3(U) < 0) for (int r = 0; r < 3; ++r) U[r * 3 + 2] *= -1.0;
  if (det3(V) < 0) for (int r = 0; r < 3; ++r) V[r * 3 + 2] *= -1.0;
  double Ud[9], Udt[9], R[2][9];
  for (int r = 0; r < 3; ++r) {
    Ud[r * 3 + 0] = -U[r * 3 + 1]; Ud[r * 3 + 1] = U[r * 3 + 0]; Ud[r * 3 + 2] = U[r * 3 + 2];
    Udt[r * 3 + 0] = U[r * 3 + 1]; Udt[r * 3 + 1] = -U[r * 3 + 0]; Udt[r * 3 + 2] = U[r * 3 + 2];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      R[0][r * 3 + c] = Ud[r * 3 + 0] * V[c * 3 + 0] + Ud[r * 3 + 1] * V[c * 3 + 1] + Ud[r * 3 + 2] * V[c * 3 + 2];
      R[1][r * 3 + c] = Udt[r * 3 + 0] * V[c * 3 + 0] + Udt[r * 3 + 1] * V[c * 3 + 1] + Udt[r * 3 + 2] * V[c * 3 + 2];
    }
  double t[3] = {U[2], U[5], U[8]};
  const double tn = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  for (int k = 0; k < 3; ++k) t[k] /= tn;
  int best = -1, best_count = -1;
  double pc[4][3];
  for (int i = 0; i < 4; ++i) {
    const double* Ri = R[i / 2];
    const double sgn = (i % 2 == 0) ? 1.0 : -1.0;
    for (int k = 0; k < 3; ++k) pc[i][k] = -(Ri[0 * 3 + k] * (sgn * t[0]) + Ri[1 * 3 + k] * (sgn * t[1]) + Ri[2 * 3 + k] * (sgn * t[2]));
    int count = 0;
    for (int m = 0; m < n; ++m) count += in_front(corr[4 * m], corr[4 * m + 1], corr[4 * m + 2], corr[4 * m + 3], Ri, pc[i]) ? 1 : 0;
    if (count > best_count) { best_count = count; best = i; }
  }
  for (int k = 0; k < 9; ++k) Rbest[k] = R[best / 2][k];
  for (int k = 0; k < 3; ++k) pbest[k] = pc[best][k];
  return best_count;
}

// ---- polynomial roots: theia::FindPolynomialRootsCompanionMatrix (math/find_polynomial_roots_companion_matrix.cc:
// 89-233: Parlett-Reinsch balancing :89-137, companion matrix :139-151) and the closed forms of
// math/polynomial.cc:176-231 for degree 1 and 2 -----------------------------------------------------------------
template <int D>
__device__ int companion_roots(const double* monic_tail, double* re, double* im) {
  double C[D * D], Off[D * D], Uq[D * D], M[D * D];
  for (int i = 0; i < D * D; ++i) C[i] = 0.0;
  for (int i = 1; i < D; ++i) C[i * D + i - 1] = 1.0;
  for (int i = 0; i < D; ++i) C[i * D + D - 1] = -monic_tail[D - 1 - i];
  for (int i = 0; i < D * D; ++i) Off[i] = C[i];
  for (int i = 0; i < D; ++i) Off[i * D + i] = 0.0;
  const double gamma = 0.9;
  bool changed;
  do {
    changed = false;
    for (int i = 0; i < D; ++i) {
      double row_norm = 0.0, col_norm = 0.0;
      for (int k = 0; k < D; ++k) { row_norm += fabs(Off[i * D + k]); col_norm += fabs(Off[k * D + i]); }
      int exponent = 0;
      frexp(row_norm / col_norm, &exponent);
      exponent /= 2;
      if (exponent != 0) {
        const double scaled_col = ldexp(col_norm, exponent), scaled_row = ldexp(row_norm, -exponent);
        if (scaled_col + scaled_row < gamma * (col_norm + row_norm)) {
          changed = true;
          const double fr = ldexp(1.0, -exponent), fc = ldexp(1.0, exponent);
          for (int k = 0; k < D; ++k) Off[i * D + k] *= fr;
          for (int k = 0; k < D; ++k) Off[k * D + i] *= fc;
        }
      }
    }
  } while (changed);
  for (int i = 0; i < D; ++i) Off[i * D + i] = C[i * D + i];
  sl::EigenReal<D> es;
  es.T = Off; es.Uq = Uq; es.M = M;
  es.compute(nullptr, false);
  if (!es.ok) return 0;
  for (int i = 0; i < D; ++i) { re[i] = es.eig_re[i]; im[i] = es.eig_im[i]; }
  return D;
}
__device__ int poly_roots(const double* poly_in, int size, double* re, double* im) {
  int lead = 0;
  while (lead < size - 1 && poly_in[lead] == 0.0) ++lead;
  const double* p = poly_in + lead;
  const int degree = size - lead - 1;
  if (degree == 0) return 0;
  if (degree == 1) { re[0] = -p[1] / p[0]; im[0] = 0.0; return 1; }
  if (degree == 2) {
    const double a = p[0], b = p[1], c = p[2];
    const double Dd = b * b - 4 * a * c, sqrt_D = sqrt(fabs(Dd));
    im[0] = im[1] = 0.0;
    if (Dd >= 0) {
      if (b >= 0) { re[0] = (-b - sqrt_D) / (2.0 * a); re[1] = (2.0 * c) / (-b - sqrt_D); }
      else { re[0] = (2.0 * c) / (-b + sqrt_D); re[1] = (-b + sqrt_D) / (2.0 * a); }
      return 2;
    }
    re[0] = -b / (2.0 * a); re[1] = -b / (2.0 * a);
    im[0] = sqrt_D / (2.0 * a); im[1] = -sqrt_D / (2.0 * a);
    return 2;
  }
  double tail[4];
  for (int i = 0; i < degree; ++i) tail[i] = p[1 + i] / p[0];
  if (degree == 3) return companion_roots<3>(tail, re, im);
  return companion_roots<4>(tail, re, im);
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void normalize3(double* a) { const double n = sqrt(dot3(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; }

// theia::PoseFromThreePoints (Kneip P3P, sfm/pose/perspective_three_point.cc:182-291; quartic :58-130, back-substitution
// :135-178). The real parts of all roots are kept, as the reference does.
__device__ int p3p(const double* feat, const double* world, double* Rs, double* ts) {
  double f[3][3], w[3][3];
  for (int i = 0; i < 3; ++i) {
    f[i][0] = feat[2 * i]; f[i][1] = feat[2 * i + 1]; f[i][2] = 1.0;
    normalize3(f[i]);
    for (int k = 0; k < 3; ++k) w[i][k] = world[3 * i + k];
  }
  double w10[3], w20[3], cr[3];
  for (int k = 0; k < 3; ++k) { w10[k] = w[1][k] - w[0][k]; w20[k] = w[2][k] - w[0][k]; }
  cross3(w10, w20, cr);
  if (dot3(cr, cr) < 1e-6) return 0;
  double Tc[3][3];
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = 0; k < 3; ++k) Tc[0][k] = f[0][k];
    cross3(f[0], f[1], Tc[2]); normalize3(Tc[2]);
    cross3(Tc[2], Tc[0], Tc[1]);
    if (pass == 1) break;
    if (dot3(Tc[2], f[2]) > 0) {
      for (int k = 0; k < 3; ++k) { const double t = f[0][k]; f[0][k] = f[1][k]; f[1][k] = t; }
      for (int k = 0; k < 3; ++k) { const double t = w[0][k]; w[0][k] = w[1][k]; w[1][k] = t; }
      for (int k = 0; k < 3; ++k) { w10[k] = w[1][k] - w[0][k]; w20[k] = w[2][k] - w[0][k]; }
    } else {
      break;
    }
  }
  double ip[3];
  for (int r = 0; r < 3; ++r) ip[r] = dot3(Tc[r], f[2]);
  double Nw[3][3];
  for (int k = 0; k < 3; ++k) Nw[0][k] = w10[k];
  normalize3(Nw[0]);
  cross3(Nw[0], w20, Nw[2]); normalize3(Nw[2]);
  cross3(Nw[2], Nw[0], Nw[1]);
  double iw[3];
  for (int r = 0; r < 3; ++r) iw[r] = dot3(Nw[r], w20);
  const double d_12 = sqrt(dot3(w10, w10));
  const double f_1 = ip[0] / ip[2], f_2 = ip[1] / ip[2], p_1 = iw[0], p_2 = iw[1];
  const double cos_beta = dot3(f[0], f[1]);
  double b = 1.0 / (1.0 - cos_beta * cos_beta) - 1.0;
  b = cos_beta < 0 ? -sqrt(b) : sqrt(b);
  const double f_1_pw2 = f_1 * f_1, f_2_pw2 = f_2 * f_2, p_1_pw2 = p_1 * p_1, p_1_pw3 = p_1_pw2 * p_1, p_1_pw4 = p_1_pw3 * p_1;
  const double p_2_pw2 = p_2 * p_2, p_2_pw3 = p_2_pw2 * p_2, p_2_pw4 = p_2_pw3 * p_2, d_12_pw2 = d_12 * d_12, b_pw2 = b * b;
  double co[5];
  co[0] = -f_2_pw2 * p_2_pw4 - p_2_pw4 * f_1_pw2 - p_2_pw4;
  co[1] = 2.0 * p_2_pw3 * d_12 * b + 2.0 * f_2_pw2 * p_2_pw3 * d_12 * b - 2.0 * f_2 * p_2_pw3 * f_1 * d_12;
  co[2] = -f_2_pw2 * p_2_pw2 * p_1_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 + f_2_pw2 * p_2_pw4 +
          p_2_pw4 * f_1_pw2 + 2.0 * p_1 * p_2_pw2 * d_12 + 2.0 * f_1 * f_2 * p_1 * p_2_pw2 * d_12 * b - p_2_pw2 * p_1_pw2 * f_1_pw2 +
          2.0 * p_1 * p_2_pw2 * f_2_pw2 * d_12 - p_2_pw2 * d_12_pw2 * b_pw2 - 2.0 * p_1_pw2 * p_2_pw2;
  co[3] = 2.0 * p_1_pw2 * p_2 * d_12 * b + 2.0 * f_2 * p_2_pw3 * f_1 * d_12 - 2.0 * f_2_pw2 * p_2_pw3 * d_12 * b - 2.0 * p_1 * p_2 * d_12_pw2 * b;
  co[4] = -2 * f_2 * p_2_pw2 * f_1 * p_1 * d_12 * b + f_2_pw2 * p_2_pw2 * d_12_pw2 + 2.0 * p_1_pw3 * d_12 - p_1_pw2 * d_12_pw2 +
          f_2_pw2 * p_2_pw2 * p_1_pw2 - p_1_pw4 - 2.0 * f_2_pw2 * p_2_pw2 * p_1 * d_12 + p_2_pw2 * f_1_pw2 * p_1_pw2 +
          f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2;
  double re[4], im[4];
  const int nroots = poly_roots(co, 5, re, im);
  for (int s = 0; s < nroots; ++s) {
    const double cos_theta = re[s];
    const double cot_alpha = (-f_1 * p_1 / f_2 - cos_theta * p_2 + d_12 * b) / (-f_1 * cos_theta * p_2 / f_2 + p_1 - d_12);
    const double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
    const double sin_alpha = sqrt(1.0 / (cot_alpha * cot_alpha + 1.0));
    double cos_alpha = sqrt(1.0 - sin_alpha * sin_alpha);
    if (cot_alpha < 0) cos_alpha = -cos_alpha;
    const double k = sin_alpha * b + cos_alpha;
    const double c_nu[3] = {d_12 * cos_alpha * k, cos_theta * d_12 * sin_alpha * k, sin_theta * d_12 * sin_alpha * k};
    double trans[3];
    for (int c = 0; c < 3; ++c) trans[c] = w[0][c] + (Nw[0][c] * c_nu[0] + Nw[1][c] * c_nu[1] + Nw[2][c] * c_nu[2]);
    const double Q[3][3] = {{-cos_alpha, -sin_alpha * cos_theta, -sin_alpha * sin_theta},
                            {sin_alpha, -cos_alpha * cos_theta, -cos_alpha * sin_theta},
                            {0, -sin_theta, cos_theta}};
    double QN[3][3], R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) QN[r][c] = Q[r][0] * Nw[0][c] + Q[r][1] * Nw[1][c] + Q[r][2] * Nw[2][c];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = Tc[0][r] * QN[0][c] + Tc[1][r] * QN[1][c] + Tc[2][r] * QN[2][c];
    for (int k2 = 0; k2 < 9; ++k2) Rs[9 * s + k2] = R[k2];
    for (int r = 0; r < 3; ++r) ts[3 * s + r] = -(R[r * 3] * trans[0] + R[r * 3 + 1] * trans[1] + R[r * 3 + 2] * trans[2]);
  }
  return nroots;
}

// theia::NormalizeImagePoints (sfm/pose/util.cc:81-111)
__device__ void normalize_image_points(const double* pts, int n, int stride, double* out, double* T) {
  double cx = 0.0, cy = 0.0;
  for (int i = 0; i < n; ++i) { cx += pts[i * stride]; cy += pts[i * stride + 1]; }
  cx /= n; cy /= n;
  double sq = 0.0;
  for (int i = 0; i < n; ++i) { const double dx = pts[i * stride] - cx, dy = pts[i * stride + 1] - cy; sq += dx * dx; sq += dy * dy; }
  const double rms = sqrt(sq / n);
  const double nf = sqrt(2.0) / rms;
  T[0] = nf; T[1] = 0; T[2] = -1.0 * nf * cx; T[3] = 0; T[4] = nf; T[5] = -1.0 * nf * cy; T[6] = 0; T[7] = 0; T[8] = 1;
  for (int i = 0; i < n; ++i) {
    const double x = pts[i * stride], y = pts[i * stride + 1];
    const double hx = T[0] * x + T[1] * y + T[2], hy = T[3] * x + T[4] * y + T[5], hw = T[6] * x + T[7] * y + T[8];
    out[2 * i] = hx / hw; out[2 * i + 1] = hy / hw;
  }
}
__device__ __forceinline__ void inverse3(const double* M, double* inv) {
  const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  const double det = M[0] * c00 + M[1] * c01 + M[2] * c02, id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = (M[2] * M[7] - M[1] * M[8]) * id; inv[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  inv[3] = c01 * id; inv[4] = (M[0] * M[8] - M[2] * M[6]) * id; inv[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  inv[6] = c02 * id; inv[7] = (M[1] * M[6] - M[0] * M[7]) * id; inv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}
__device__ __forceinline__ void mul33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// theia::FourPointHomography, minimal case (sfm/pose/four_point_homography.cc:72-102)
__device__ bool four_point_h(const double* corr, double* H) {
  double n1[8], n2[8], T1[9], T2[9];
  normalize_image_points(corr, 4, 4, n1, T1);
  normalize_image_points(corr + 2, 4, 4, n2, T2);
  double AtA[81], W[81], V[81], S[9];
  {
    double A[8 * 9];
    for (int i = 0; i < 4; ++i) {
      const double x = n1[2 * i], y = n1[2 * i + 1], u = n2[2 * i], v = n2[2 * i + 1];
      double* r0 = A + 18 * i; double* r1 = r0 + 9;
      r0[0] = 0; r0[1] = 0; r0[2] = 0; r0[3] = -x; r0[4] = -y; r0[5] = -1.0; r0[6] = x * v; r0[7] = y * v; r0[8] = v;
      r1[0] = x; r1[1] = y; r1[2] = 1.0; r1[3] = 0; r1[4] = 0; r1[5] = 0; r1[6] = -x * u; r1[7] = -y * u; r1[8] = -u;
    }
    for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) { double s = 0.0; for (int k = 0; k < 8; ++k) s += A[k * 9 + r] * A[k * 9 + c]; AtA[r * 9 + c] = s; }
  }
  sl::jacobi_svd<9, false>(AtA, W, nullptr, S, V);  // jacobiSvd(ComputeFullV): only the null vector is used
  double Hn[9], T2i[9], tmp[9];
  for (int k = 0; k < 9; ++k) Hn[k] = V[k * 9 + 8];
  inverse3(T2, T2i);
  mul33(T2i, Hn, tmp);
  mul33(tmp, T1, H);
  return true;
}

// theia::SevenPointFundamentalMatrix (sfm/pose/seven_point_fundamental_matrix.cc:72-152), including the reference's
// coefficient-order quirk: the det(F2)-only term sits at index 0 although polynomial(0) is the highest-degree
// coefficient, and the real parts of all roots are used.
__device__ int seven_point_f(const double* corr, double* F_out) {
  double n1[14], n2[14], T1[9], T2[9];
  normalize_image_points(corr, 7, 4, n1, T1);
  normalize_image_points(corr + 2, 7, 4, n2, T2);
  double ns[18];
  {
    double epi[7 * 9];
    for (int i = 0; i < 7; ++i) {
      const double ax = n1[2 * i], ay = n1[2 * i + 1], bx = n2[2 * i], by = n2[2 * i + 1];
      double* r = epi + 9 * i;
      r[0] = bx * ax; r[1] = by * ax; r[2] = ax; r[3] = bx * ay; r[4] = by * ay; r[5] = ay; r[6] = bx; r[7] = by; r[8] = 1.0;
    }
    sl::FullPivLU<7, 9> lu;
    lu.lu = epi;
    lu.compute();
    if (9 - lu.rank() != 2) return 0;
    sl::kernel_rx9<7>(lu, ns);
  }
  double v1[9], v2[9];
  for (int k = 0; k < 9; ++k) { v1[k] = ns[k * 2] - ns[k * 2 + 1]; v2[k] = ns[k * 2 + 1]; }
#define F1(r, c) v1[(c) * 3 + (r)]
#define F2(r, c) v2[(c) * 3 + (r)]
  double dc[4];
  dc[0] = -(F2(1, 2) * F2(2, 1) - F2(1, 1) * F2(2, 2)) * F2(0, 0) + (F2(0, 2) * F2(2, 1) - F2(0, 1) * F2(2, 2)) * F2(1, 0) -
          (F2(0, 2) * F2(1, 1) - F2(0, 1) * F2(1, 2)) * F2(2, 0);
  dc[1] = -(F2(1, 2) * F2(2, 1) - F2(1, 1) * F2(2, 2)) * F1(0, 0) + (F2(0, 2) * F2(2, 1) - F2(0, 1) * F2(2, 2)) * F1(1, 0) -
          (F2(0, 2) * F2(1, 1) - F2(0, 1) * F2(1, 2)) * F1(2, 0) +
          (F1(2, 2) * F2(1, 1) - F1(2, 1) * F2(1, 2) - F1(1, 2) * F2(2, 1) + F1(1, 1) * F2(2, 2)) * F2(0, 0) -
          (F1(2, 2) * F2(0, 1) - F1(2, 1) * F2(0, 2) - F1(0, 2) * F2(2, 1) + F1(0, 1) * F2(2, 2)) * F2(1, 0) +
          (F1(1, 2) * F2(0, 1) - F1(1, 1) * F2(0, 2) - F1(0, 2) * F2(1, 1) + F1(0, 1) * F2(1, 2)) * F2(2, 0);
  dc[2] = (F1(2, 2) * F2(1, 1) - F1(2, 1) * F2(1, 2) - F1(1, 2) * F2(2, 1) + F1(1, 1) * F2(2, 2)) * F1(0, 0) -
          (F1(2, 2) * F2(0, 1) - F1(2, 1) * F2(0, 2) - F1(0, 2) * F2(2, 1) + F1(0, 1) * F2(2, 2)) * F1(1, 0) +
          (F1(1, 2) * F2(0, 1) - F1(1, 1) * F2(0, 2) - F1(0, 2) * F2(1, 1) + F1(0, 1) * F2(1, 2)) * F1(2, 0) -
          (F1(1, 2) * F1(2, 1) - F1(1, 1) * F1(2, 2)) * F2(0, 0) + (F1(0, 2) * F1(2, 1) - F1(0, 1) * F1(2, 2)) * F2(1, 0) -
          (F1(0, 2) * F1(1, 1) - F1(0, 1) * F1(1, 2)) * F2(2, 0);
  dc[3] = -(F1(1, 2) * F1(2, 1) - F1(1, 1) * F1(2, 2)) * F1(0, 0) + (F1(0, 2) * F1(2, 1) - F1(0, 1) * F1(2, 2)) * F1(1, 0) -
          (F1(0, 2) * F1(1, 1) - F1(0, 1) * F1(1, 2)) * F1(2, 0);
  double re[4], im[4];
  const int nroots = poly_roots(dc, 4, re, im);
  double T2t[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T2t[r * 3 + c] = T2[c * 3 + r];
  for (int s = 0; s < nroots; ++s) {
    double M[9], tmp[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r * 3 + c] = re[s] * F1(r, c) + F2(r, c);
    mul33(T2t, M, tmp);
    mul33(tmp, T1, F_out + 9 * s);
  }
#undef F1
#undef F2
  return nroots;
}


// ---- LO-RANSAC refinement of a relative pose -------------------------------------------------------------------------------
// RelativePoseEstimator::RefineModel (sfm/estimators/estimate_relative_pose.cc:111-138) = BundleAdjustTwoViewsAngular
// (sfm/bundle_adjustment/bundle_adjust_two_views.cc:195-246): rotation (angle-axis) + position on the unit sphere
// (SphereManifold<3>), AngularEpipolarError residuals (angular_epipolar_error.h:50-108) under a TRUNCATED loss of width
// error_thresh, at most 15 trust-region iterations with Ceres' default tolerances. The whole CTA works on one refinement:
// every evaluation is a pass over the model's inliers (one residual, differentiated with a 6-direction dual number, per thread
// and step), J^T J (15) / J^T r (5) / cost are block-reduced in a fixed order and thread 0 takes the trust-region decisions.
__device__ void eigen_matrix_to_angle_axis(const double* R, double aa[3]) {  // Eigen: Quaterniond(R), then AngleAxisd(q)
  double q[4];
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  double angle, axis[3];
  if (n != 0.0) {
    angle = 2.0 * atan2(n, fabs(q[3]));
    if (q[3] < 0.0) n = -n;
    for (int k = 0; k < 3; ++k) axis[k] = q[k] / n;
  } else { angle = 0.0; axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; }
  for (int k = 0; k < 3; ++k) aa[k] = angle * axis[k];
}
__device__ void eigen_angle_axis_to_matrix(const double aa[3], double* R) {  // AngleAxisd(|aa|, aa / |aa|).toRotationMatrix()
  const double angle = sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  const double ax[3] = {aa[0] / angle, aa[1] / angle, aa[2] / angle};
  const double s = sin(angle), c = cos(angle);
  const double sa[3] = {s * ax[0], s * ax[1], s * ax[2]}, c1[3] = {(1.0 - c) * ax[0], (1.0 - c) * ax[1], (1.0 - c) * ax[2]};
  double tmp;
  tmp = c1[0] * ax[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
  tmp = c1[0] * ax[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
  tmp = c1[1] * ax[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
  R[0] = c1[0] * ax[0] + c; R[4] = c1[1] * ax[1] + c; R[8] = c1[2] * ax[2] + c;
}
template <typename T>
__device__ __forceinline__ void ceres_angle_axis_to_matrix(const T* aa, T R[3][3]) {  // ceres/rotation.h AngleAxisToRotationMatrix
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (val(theta2) > DBL_EPSILON) {
    const T theta = d_sqrt(theta2);
    const T wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const T ct = d_cos(theta), st = d_sin(theta);
    R[0][0] = ct + wx * wx * (1.0 - ct);      R[1][0] = wz * st + wx * wy * (1.0 - ct);  R[2][0] = -wy * st + wx * wz * (1.0 - ct);
    R[0][1] = wx * wy * (1.0 - ct) - wz * st; R[1][1] = ct + wy * wy * (1.0 - ct);       R[2][1] = wx * st + wy * wz * (1.0 - ct);
    R[0][2] = wy * st + wx * wz * (1.0 - ct); R[1][2] = -wx * st + wy * wz * (1.0 - ct); R[2][2] = ct + wz * wz * (1.0 - ct);
  } else {
    R[0][0] = T(1.0); R[1][0] = aa[2]; R[2][0] = -aa[1];
    R[0][1] = -aa[2]; R[1][1] = T(1.0); R[2][1] = aa[0];
    R[0][2] = aa[1]; R[1][2] = -aa[0]; R[2][2] = T(1.0);
  }
}
template <typename T>
__device__ __forceinline__ T angular_epipolar_residual(const T* rot, const T* tr, const double* c) {
  const T f1[3] = {T(c[0]), T(c[1]), T(1.0)}, f2[3] = {T(c[2]), T(c[3]), T(1.0)};
  T R[3][3];
  ceres_angle_axis_to_matrix(rot, R);
  T Rf2[3], Rtf2[3], Tf1[3], TRtf2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { Rf2[i] = R[i][0] * f2[0] + R[i][1] * f2[1] + R[i][2] * f2[2]; Rtf2[i] = R[0][i] * f2[0] + R[1][i] * f2[1] + R[2][i] * f2[2]; }
  {
    const T tv = tr[0] * f1[0] + tr[1] * f1[1] + tr[2] * f1[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) Tf1[i] = f1[i] - tr[i] * tv;
    const T tw = tr[0] * Rtf2[0] + tr[1] * Rtf2[1] + tr[2] * Rtf2[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) TRtf2[i] = Rtf2[i] - tr[i] * tw;
  }
  const T a = (f1[0] * Tf1[0] + f1[1] * Tf1[1] + f1[2] * Tf1[2]) + (Rf2[0] * TRtf2[0] + Rf2[1] * TRtf2[1] + Rf2[2] * TRtf2[2]);
  const T cr[3] = {f1[1] * Rtf2[2] - f1[2] * Rtf2[1], f1[2] * Rtf2[0] - f1[0] * Rtf2[2], f1[0] * Rtf2[1] - f1[1] * Rtf2[0]};
  const T b_sqrt = tr[0] * cr[0] + tr[1] * cr[1] + tr[2] * cr[2];
  const T sqrt_term = (a * a) / 4.0 - b_sqrt * b_sqrt;
  if (val(sqrt_term) < 0.0) return T(1000.0);
  return a / 2.0 - d_sqrt(sqrt_term);
}
__device__ void householder3(const double x[3], double v[3], double* beta) {
  const double sigma = x[0] * x[0] + x[1] * x[1];
  v[0] = x[0]; v[1] = x[1]; v[2] = 1.0;
  *beta = 0.0;
  if (sigma <= DBL_EPSILON) { if (x[2] < 0.0) *beta = 2.0; return; }
  const double mu = sqrt(x[2] * x[2] + sigma);
  const double vp = x[2] <= 0.0 ? x[2] - mu : -sigma / (x[2] + mu);
  *beta = 2.0 * vp * vp / (sigma + vp * vp);
  v[0] /= vp; v[1] /= vp;
}
__device__ void sphere3_plus(const double x[3], const double d[2], double out[3]) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1]);
  if (nd == 0.0) { for (int i = 0; i < 3; ++i) out[i] = x[i]; return; }
  double v[3], beta;
  householder3(x, v, &beta);
  const double sbd = sin(nd) / nd;
  const double y[3] = {sbd * d[0], sbd * d[1], cos(nd)};
  const double vty = v[0] * y[0] + v[1] * y[1] + v[2] * y[2];
  const double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int i = 0; i < 3; ++i) out[i] = nx * (y[i] - v[i] * (beta * vty));
}
__device__ void sphere3_plus_jacobian(const double x[3], double J[6]) {
  double v[3], beta;
  householder3(x, v, &beta);
  const double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int i = 0; i < 2; ++i) {
    for (int r = 0; r < 3; ++r) J[r * 2 + i] = -beta * v[i] * v[r];
    J[i * 2 + i] += 1.0;
  }
  for (int k = 0; k < 6; ++k) J[k] *= nx;
}
__device__ bool spd_solve5(const double* A, const double* b, double* x) {
  double L[25], y[5];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = A[i * 5 + j];
      for (int k = 0; k < j; ++k) v -= L[i * 5 + k] * L[j * 5 + k];
      if (i == j) { if (!(v > 0.0)) return false; L[i * 5 + i] = sqrt(v); }
      else L[i * 5 + j] = v / L[j * 5 + j];
    }
  for (int i = 0; i < 5; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[i * 5 + k] * y[k]; y[i] = v / L[i * 5 + i]; }
  for (int i = 4; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 5; ++k) v -= L[k * 5 + i] * x[k]; x[i] = v / L[i * 5 + i]; }
  return true;
}

struct LoShared {
  double rot[3], pos[3], crot[3], cpos[3], scale[5], PJ[6];
  double tot[21], red[RT / 32][21];
  int go, ok;
};
// one pass over the flagged data at (rot, pos): tot = [H(15) | g(5) | cost]
template <bool WANT_J>
__device__ void lo_pass(const double* __restrict__ corr, const uint8_t* __restrict__ flag, int n, double b, const double* rot, const double* pos,
                        LoShared& L) {
  double v[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) v[k] = 0.0;
  for (int i = threadIdx.x; i < n; i += RT) {
    if (!flag[i]) continue;
    const double c[4] = {corr[4 * (size_t)i], corr[4 * (size_t)i + 1], corr[4 * (size_t)i + 2], corr[4 * (size_t)i + 3]};
    if (!WANT_J) {
      const double r = angular_epipolar_residual<double>(rot, pos, c);
      v[20] += 0.5 * fmin(r * r, b);
    } else {
      typedef Dual<6> D6;
      D6 jr[3], jt[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { jr[k] = seed<6>(rot[k], k); jt[k] = seed<6>(pos[k], 3 + k); }
      const D6 res = angular_epipolar_residual<D6>(jr, jt, c);
      const double s = res.a * res.a;
      v[20] += 0.5 * fmin(s, b);
      if (!(s < b)) continue;  // TruncatedLoss: rho' = 0 beyond the width - the Corrector zeroes the row
      double t[5];
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = res.v[k] * L.scale[k];
#pragma unroll
      for (int k = 0; k < 2; ++k) t[3 + k] = (res.v[3] * L.PJ[k] + res.v[4] * L.PJ[2 + k] + res.v[5] * L.PJ[4 + k]) * L.scale[3 + k];
      int q = 0;
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        v[15 + a] += t[a] * res.a;
#pragma unroll
        for (int cc = 0; cc <= a; ++cc) v[q++] += t[a] * t[cc];
      }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = WANT_J ? 0 : 20; k < 21; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) L.red[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 21) {
    double s = 0.0;
    for (int ww = 0; ww < RT / 32; ++ww) s += L.red[ww][threadIdx.x];
    L.tot[threadIdx.x] = s;
  }
  __syncthreads();
}
// Refines m->R / m->p over the flagged correspondences (m->E is left as it is, like the reference). Every thread of the CTA
// calls it; returns final_cost < initial_cost (uniform).
__device__ bool lo_refine_relpose(const double* __restrict__ corr, const uint8_t* __restrict__ flag, int n, double width, Model* m, LoShared& L) {
  const int t = threadIdx.x;
  const double b = width * width;
  if (t == 0) {
    eigen_matrix_to_angle_axis(m->R, L.rot);
    for (int k = 0; k < 3; ++k) L.pos[k] = m->p[k];
    for (int k = 0; k < 5; ++k) L.scale[k] = 1.0;
    sphere3_plus_jacobian(L.pos, L.PJ);
  }
  __syncthreads();
  lo_pass<true>(corr, flag, n, b, L.rot, L.pos, L);
  if (t == 0) { int q = 0; for (int a = 0; a < 5; ++a) { q += a; L.scale[a] = 1.0 / (1.0 + sqrt(L.tot[q])); ++q; } }
  __syncthreads();
  lo_pass<true>(corr, flag, n, b, L.rot, L.pos, L);
  double H[15], g[5], diag[5], x_cost = L.tot[20], min_cost = L.tot[20], initial_cost = L.tot[20], x_norm = 0.0, radius = 1e4, decrease_factor = 2.0, mcc = 0.0;
  bool step_ok = true, reuse_diag = false;
  int iteration = 0, invalid = 0;
  if (t == 0) {
    for (int k = 0; k < 15; ++k) H[k] = L.tot[k];
    for (int k = 0; k < 5; ++k) g[k] = L.tot[15 + k];
    for (int k = 0; k < 3; ++k) x_norm += L.rot[k] * L.rot[k] + L.pos[k] * L.pos[k];
    x_norm = sqrt(x_norm);
  }
  for (;;) {
    if (t == 0) {
      int go = 1;
      for (;;) {
        if (iteration >= 15) { go = 0; break; }
        if (step_ok) { double gmax = 0.0; for (int k = 0; k < 5; ++k) gmax = fmax(gmax, fabs(g[k] / L.scale[k])); if (gmax <= 1e-10) { go = 0; break; } }
        if (radius <= 1e-32) { go = 0; break; }
        ++iteration;
        step_ok = false;
        if (!reuse_diag) { int q = 0; for (int a = 0; a < 5; ++a) { q += a; diag[a] = fmin(fmax(H[q], 1e-6), 1e32); ++q; } }
        reuse_diag = true;
        double M[25], y[5];
        { int q = 0; for (int a = 0; a < 5; ++a) for (int c = 0; c <= a; ++c) { M[a * 5 + c] = H[q]; M[c * 5 + a] = H[q]; ++q; } }
        for (int a = 0; a < 5; ++a) M[a * 5 + a] += diag[a] / radius;
        bool valid = spd_solve5(M, g, y);
        if (valid) {
          double yg = 0.0, yHy = 0.0;
          int q = 0;
          for (int a = 0; a < 5; ++a) { yg += y[a] * g[a]; for (int c = 0; c <= a; ++c) { yHy += (a == c ? 1.0 : 2.0) * y[a] * H[q] * y[c]; ++q; } }
          mcc = yg - 0.5 * yHy;
          valid = isfinite(mcc) && mcc > 0.0;
        }
        if (!valid) {
          if (++invalid >= 5) { go = 0; break; }
          radius /= decrease_factor; decrease_factor *= 2.0;
          continue;
        }
        invalid = 0;
        const double d2[2] = {-y[3] * L.scale[3], -y[4] * L.scale[4]};
        for (int k = 0; k < 3; ++k) L.crot[k] = L.rot[k] + (-y[k] * L.scale[k]);
        sphere3_plus(L.pos, d2, L.cpos);
        break;
      }
      L.go = go;
    }
    __syncthreads();
    if (L.go == 0) break;  // L.go is next written after the barriers inside lo_pass
    lo_pass<false>(corr, flag, n, b, L.crot, L.cpos, L);
    if (t == 0) {
      int go = 1;
      const double cand_cost = L.tot[20];
      double sn = 0.0, cn = 0.0;
      for (int k = 0; k < 3; ++k) {
        sn += (L.crot[k] - L.rot[k]) * (L.crot[k] - L.rot[k]) + (L.cpos[k] - L.pos[k]) * (L.cpos[k] - L.pos[k]);
        cn += L.crot[k] * L.crot[k] + L.cpos[k] * L.cpos[k];
      }
      const double cost_change = x_cost - cand_cost;
      if (sqrt(sn) <= 1e-8 * (x_norm + 1e-8)) go = 0;
      else if (fabs(cost_change) <= 1e-6 * x_cost) go = 0;
      else {
        const double rel = cost_change / mcc;
        if (rel > 1e-3) {
          for (int k = 0; k < 3; ++k) { L.rot[k] = L.crot[k]; L.pos[k] = L.cpos[k]; }
          x_norm = sqrt(cn);
          sphere3_plus_jacobian(L.pos, L.PJ);
          const double u = 2.0 * rel - 1.0;
          radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - u * u * u));
          decrease_factor = 2.0; reuse_diag = false;
          go = 2;
        } else {
          radius /= decrease_factor; decrease_factor *= 2.0;
        }
      }
      L.go = go;
    }
    __syncthreads();
    const int go2 = L.go;
    __syncthreads();  // thread 0 rewrites L.go at the top of the next round: everybody has read it by now
    if (go2 == 0) break;
    if (go2 == 2) {
      lo_pass<true>(corr, flag, n, b, L.rot, L.pos, L);
      if (t == 0) {
        for (int k = 0; k < 15; ++k) H[k] = L.tot[k];
        for (int k = 0; k < 5; ++k) g[k] = L.tot[15 + k];
        x_cost = L.tot[20];
        min_cost = fmin(min_cost, x_cost);
        step_ok = true;
      }
    }
  }
  if (t == 0) {
    for (int k = 0; k < 3; ++k) m->p[k] = L.pos[k];
    eigen_angle_axis_to_matrix(L.rot, m->R);
    L.ok = min_cost < initial_cost ? 1 : 0;
  }
  __syncthreads();
  return L.ok != 0;
}

// ---- estimator policies (solvers/estimator.h): sample size, datum width, model estimation, per-datum error ----
struct RelPoseEst {  // RelativePoseEstimator (sfm/estimators/estimate_relative_pose.cc:65-155)
  static constexpr int S = 5, D = 4, MAXM = 10;
  static constexpr int SOLVE_CTAS = RANSAC_CTAS_PER_SM;  // resident solver CTAs per SM in the round-synchronous schedule
  static constexpr bool HAS_LO = true;
  __device__ static bool refine(const double* corr, const uint8_t* flag, int n, double thresh, Model* m, LoShared& L) {
    return lo_refine_relpose(corr, flag, n, thresh, m, L);
  }
  __device__ static int solve(const double* sample, Model* out) {
    double x1[10], x2[10], Es[90];
    for (int i = 0; i < 5; ++i) { x1[2 * i] = sample[4 * i]; x1[2 * i + 1] = sample[4 * i + 1]; x2[2 * i] = sample[4 * i + 2]; x2[2 * i + 1] = sample[4 * i + 3]; }
    const int ne = five_point(x1, x2, Es);
    int n = 0;
    for (int e = 0; e < ne; ++e) {
      Model m;
      for (int k = 0; k < 9; ++k) m.E[k] = Es[9 * e + k];
      if (best_pose(m.E, sample, 5, m.R, m.p) < 4) continue;
      out[n++] = m;
    }
    return n;
  }
  __device__ __forceinline__ static double error(const double* E, const double* R, const double* p, const double* c) {
    return in_front(c[0], c[1], c[2], c[3], R, p) ? sampson(E, c[0], c[1], c[2], c[3]) : DBL_MAX;
  }
};
struct AbsPoseEst {  // CalibratedAbsolutePoseEstimator, PnPType::KNEIP (sfm/estimators/estimate_calibrated_absolute_pose.cc:63-172)
  static constexpr int S = 3, D = 5, MAXM = 4;  // datum: feature (x, y), world point (X, Y, Z)
  static constexpr int SOLVE_CTAS = RANSAC_CTAS_PER_SM;
  static constexpr bool HAS_LO = false;  // its LO is BundleAdjustView on a one-view reconstruction (:120-153): rejected
  __device__ static bool refine(const double*, const uint8_t*, int, double, Model*, LoShared&) { return false; }
  __device__ static int solve(const double* sample, Model* out) {
    double feat[6], world[9], Rs[36], ts[12];
    for (int i = 0; i < 3; ++i) { feat[2 * i] = sample[5 * i]; feat[2 * i + 1] = sample[5 * i + 1]; for (int k = 0; k < 3; ++k) world[3 * i + k] = sample[5 * i + 2 + k]; }
    const int n = p3p(feat, world, Rs, ts);
    for (int s = 0; s < n; ++s) {
      Model& m = out[s];
      for (int k = 0; k < 9; ++k) { m.E[k] = 0.0; m.R[k] = Rs[9 * s + k]; }
      for (int c = 0; c < 3; ++c) m.p[c] = -(m.R[0 * 3 + c] * ts[3 * s] + m.R[1 * 3 + c] * ts[3 * s + 1] + m.R[2 * 3 + c] * ts[3 * s + 2]);
    }
    return n;
  }
  __device__ __forceinline__ static double error(const double*, const double* R, const double* p, const double* d) {
    const double v0 = d[2] - p[0], v1 = d[3] - p[1], v2 = d[4] - p[2];
    const double px = fma(R[0], v0, fma(R[1], v1, R[2] * v2));
    const double py = fma(R[3], v0, fma(R[4], v1, R[5] * v2));
    const double pz = fma(R[6], v0, fma(R[7], v1, R[8] * v2));
    const double ex = px / pz - d[0], ey = py / pz - d[1];
    return fma(ex, ex, ey * ey);
  }
};
struct HomographyEst {  // HomographyEstimator (sfm/estimators/estimate_homography.cc:62-116); H lives in Model::E
  static constexpr int S = 4, D = 4, MAXM = 1;
  static constexpr int SOLVE_CTAS = 4;  // 128 registers: a C5 view graph (512 pairs) is one wave of solver CTAs instead of two
  static constexpr bool HAS_LO = false;
  __device__ static bool refine(const double*, const uint8_t*, int, double, Model*, LoShared&) { return false; }
  __device__ static int solve(const double* sample, Model* out) {
    for (int k = 0; k < 9; ++k) out->R[k] = 0.0;
    for (int k = 0; k < 3; ++k) out->p[k] = 0.0;
    return four_point_h(sample, out->E) ? 1 : 0;
  }
  __device__ __forceinline__ static double error(const double* H, const double*, const double*, const double* c) {
    const double px = fma(H[0], c[0], fma(H[1], c[1], H[2]));
    const double py = fma(H[3], c[0], fma(H[4], c[1], H[5]));
    const double pz = fma(H[6], c[0], fma(H[7], c[1], H[8]));
    const double ex = c[2] - px / pz, ey = c[3] - py / pz;
    return fma(ex, ex, ey * ey);
  }
};

// ---- std::mt19937 + libstdc++ std::uniform_int_distribution<int> (util/random.cc:46-84) -----------------
struct Mt19937 {
  uint32_t mt[624];
  int idx;
};
__device__ void mt_seed(Mt19937* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
__device__ uint32_t mt_next(Mt19937* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
// uniform_int_distribution<int>(lo, hi): GCC >= 11, 32-bit generator -> Lemire's nearly divisionless method
__device__ int mt_uniform_int(Mt19937* g, int lo, int hi) {
  const uint32_t range = (uint32_t)(hi - lo) + 1u;  // hi - lo < 2^32 - 1 always holds here
  uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return lo + (int)(product >> 32);
}

// ProsacSampler (solvers/prosac_sampler.cc:53-131), RansacType::PROSAC: the k-th sample comes from the top n data points of a
// quality-sorted input, n grown by the schedule of Chum & Matas with T_N = 20000. The reference recomputes the schedule from
// t = 1 on every call; carrying (t_n, t_n_prime, n) performs the same operations in the same order. Its last index `n` may equal
// the number of data points once the schedule covers the whole set (an out-of-bounds read in the reference): clamped.
struct ProsacState {
  double t_n, t_n_prime;
  int n, k, N, m;
};
__device__ void prosac_init(ProsacState* s, int num_datapoints, int min_num_samples) {
  s->N = num_datapoints; s->m = min_num_samples; s->k = 1; s->n = min_num_samples; s->t_n_prime = 1.0;
  double t_n = 20000.0;
  for (int i = 0; i < min_num_samples; ++i) t_n *= (double)(s->n - i) / (double)(num_datapoints - i);
  s->t_n = t_n;
}
__device__ void prosac_sample(ProsacState* s, Mt19937* g, int* subset) {
  if ((double)s->k > s->t_n_prime && s->n < s->N) {
    const double t_n_plus1 = (s->t_n * (s->n + 1.0)) / (s->n + 1.0 - s->m);
    s->t_n_prime += ceil(t_n_plus1 - s->t_n);
    s->t_n = t_n_plus1;
    ++s->n;
  }
  const bool from_top_n = s->t_n_prime < (double)s->k;
  const int count = from_top_n ? s->m : s->m - 1, hi = from_top_n ? s->n - 1 : s->n - 2;
  for (int i = 0; i < count; ++i) {
    int r;
    bool dup;
    do {
      r = mt_uniform_int(g, 0, hi);
      dup = false;
      for (int j = 0; j < i; ++j) dup |= subset[j] == r;
    } while (dup);
    subset[i] = r;
  }
  if (!from_top_n) subset[s->m - 1] = s->n < s->N - 1 ? s->n : s->N - 1;
  ++s->k;
}

// SampleConsensusEstimator::ComputeMaxIterations (sample_consensus_estimator.h:251-297)
__device__ int compute_max_iterations(const ThbRansacParams& P, double min_sample_size, double inlier_ratio,
                                      double log_failure_prob, int total) {
  if (inlier_ratio == 1.0) return P.min_iterations;
  const int ninl = (int)(inlier_ratio * total);
  const double num_samples = P.use_tdd_test ? min_sample_size + 1 : min_sample_size;
  double a = 1.0, b = 1.0;
  for (int i = 0; i < num_samples; ++i) { a *= ninl - i; b *= total - i; }
  const double prob_all_inliers = a / b;
  if (prob_all_inliers < DBL_EPSILON) return P.max_iterations;
  if (prob_all_inliers >= 1.0 - DBL_EPSILON) return P.min_iterations;
  const double num_iterations = log_failure_prob / log(1.0 - prob_all_inliers);
  return (int)fmax((double)P.min_iterations, fmin(num_iterations, (double)P.max_iterations));
}

// Warp-wide score of one model over all data. Returns (cost, #inliers) in every lane; cost = +inf if the model was
// abandoned because its partial cost reached `bail`. If mask != nullptr the inlier flags are written.
template <class Est>
__device__ void score_model(const ThbRansacParams& P, const double* __restrict__ data, int n, const Model& m,
                            double bail, uint8_t* __restrict__ mask, double* cost_out, int* ninl_out, unsigned* scored) {
  // datum i at data[i * D .. i * D + D): the caller's array-of-structs, read through L1
  const int lane = threadIdx.x & 31;
  double cost = 0.0;
  int ninl = 0;
  double E[9], R[9], p[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) { E[k] = m.E[k]; R[k] = m.R[k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = m.p[k];
  const double thresh = P.error_thresh;
  const int steps = (n + 31) / 32;
  for (int s = 0; s < steps; ++s) {
    const int i = s * 32 + lane;
    if (i < n) {
      double d[Est::D];
#pragma unroll
      for (int k = 0; k < Est::D; ++k) d[k] = data[(size_t)i * Est::D + k];
      const double r = Est::error(E, R, p, d);
      const bool inl = r < thresh;
      if (P.use_mle) cost += inl ? r : thresh; else cost += inl ? 0.0 : 1.0;
      ninl += inl ? 1 : 0;
      if (mask) mask[i] = inl ? 1 : 0;
    }
    if ((s & 15) == 15 && bail < DBL_MAX) {
      const double partial = warp_sum(cost);
      if (partial >= bail) { *cost_out = INFINITY; *ninl_out = 0; *scored += min(n, (s + 1) * 32); return; }
    }
  }
  *scored += n;
  cost = warp_sum(cost);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ninl += __shfl_xor_sync(0xffffffffu, ninl, o);
  *cost_out = cost;
  *ninl_out = ninl;
}

// LmedQualityMeasurement::ComputeCost (solvers/lmed_quality_measurement.h:58-118), warp-wide: cost = median of the SQUARED
// residuals (upper median for an even count, mean of the two middle values for an odd one - the reference's nth_element code taken
// literally), inliers = data with residual^2 below (2.5 * 1.4826 * (1 + 5 / (n - m)) * sqrt(median))^2. The order statistic is found
// by an MSB-first radix select on the bit patterns of |r| (monotone for non-negative doubles, and squaring keeps the order, so the
// k-th smallest r^2 is the square of the k-th smallest |r|): 8 passes of 8 bits, each pass RE-EVALUATES the residuals (a Sampson
// error is ~40 flops; the correspondences stay in L1) and histograms the next byte of the keys that match the prefix found so far
// in 256 shared-memory counters per warp. No residual array is stored, so n is unbounded. `hist`: 256 counters of this warp.
template <class Est>
__device__ void score_model_lmed(const double* __restrict__ data, int n, const Model& m, unsigned* __restrict__ hist,
                                 uint8_t* __restrict__ mask, double* cost_out, int* ninl_out, unsigned* scored) {
  const int lane = threadIdx.x & 31;
  double E[9], R[9], p[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) { E[k] = m.E[k]; R[k] = m.R[k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = m.p[k];
  const int steps = (n + 31) / 32;
  auto residual = [&](int i) {
    double d[Est::D];
#pragma unroll
    for (int k = 0; k < Est::D; ++k) d[k] = data[(size_t)i * Est::D + k];
    return Est::error(E, R, p, d);
  };
  unsigned long long prefix = 0;
  unsigned rank = (unsigned)(n / 2);  // 0-based rank of the element nth_element puts at size / 2
  for (int shift = 56; shift >= 0; shift -= 8) {
    const unsigned long long himask = shift == 56 ? 0ull : ~0ull << (shift + 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[lane * 8 + k] = 0;
    __syncwarp();
    for (int s = 0; s < steps; ++s) {
      const int i = s * 32 + lane;
      if (i < n) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(residual(i)));
        if (((key ^ prefix) & himask) == 0) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
      }
    }
    __syncwarp();
    unsigned c[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; sum += c[k]; }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const unsigned excl = incl - sum;
    const bool mine = excl <= rank && rank < incl;
    unsigned bin = 0, nrank = 0;
    if (mine) {
      unsigned acc = excl;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (rank >= acc && rank < acc + c[k]) { bin = lane * 8 + k; nrank = rank - acc; }
        acc += c[k];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    bin = __shfl_sync(0xffffffffu, bin, src);
    rank = __shfl_sync(0xffffffffu, nrank, src);
    prefix |= (unsigned long long)bin << shift;
    __syncwarp();
  }
  const double vk = __longlong_as_double((long long)prefix);
  double median = vk * vk;
  if (n & 1) {  // the element at size / 2 - 1: the largest value below vk, or vk itself when it has duplicates there
    int below = 0;
    double vmax = 0.0;
    for (int s = 0; s < steps; ++s) {
      const int i = s * 32 + lane;
      if (i < n) {
        const double r = fabs(residual(i));
        if (r < vk) { ++below; vmax = fmax(vmax, r); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { below += __shfl_xor_sync(0xffffffffu, below, o); vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o)); }
    const double vlo = below >= n / 2 ? vmax : vk;
    median = 0.5 * (vlo * vlo + median);
    *scored += n;
  }
  const double inlier_threshold = 2.5 * 1.4826 * (1 + 5.0 / (double)(n - Est::S)) * sqrt(median);
  const double squared_inlier_threshold = inlier_threshold * inlier_threshold;
  int ninl = 0;
  for (int s = 0; s < steps; ++s) {
    const int i = s * 32 + lane;
    if (i < n) {
      const double r = residual(i);
      const bool inl = (r * r) < squared_inlier_threshold;
      ninl += inl ? 1 : 0;
      if (mask) mask[i] = inl ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ninl += __shfl_xor_sync(0xffffffffu, ninl, o);
  *scored += 9u * (unsigned)n;
  *cost_out = median;
  *ninl_out = ninl;
}

struct RansacShared {
  Model best;
  int nmodels[BI];
  int model_start[BI + 1];
  int samples[BI][5];  // up to 5 indices per sample
  Mt19937 rng;
  ProsacState prosac;
  double best_cost;
  int max_iterations, it0, finished, num_iterations, have_best, pair;
  unsigned long long stat_samples, stat_models, stat_data;
  long long cyc[4];  // draw, solve, score, scan
  // LO-RANSAC: the scan stops at an improved model that has to be refined and resumes after the refinement
  int scan_b, scan_k, scan_it, lo_pending, num_lo;
  double lo_ratio;
  LoShared lo;
};

// Persistent grid (three CTAs per SM) pulling pairs from an atomic counter: RANSAC iteration counts differ by 100x
// between pairs, and a CTA is latency-bound while its two solver warps run, so several CTAs per SM in different phases
// are what keeps the FP64 pipes busy. (r01: one 256-thread CTA of 255 registers per SM, the candidate models in 54 KB of
// shared memory, one solver warp: ~15 % of the FP64 issue rate.) The candidate models of a batch live in a per-CTA
// global scratch area (L2-resident) so that shared memory only holds the pair's correspondences and the control block.
template <class Est>
__global__ void __launch_bounds__(RT, RANSAC_CTAS_PER_SM) k_ransac(const ThbRansacParams P_in, int num_pairs, const long long* __restrict__ pair_offset,
                                                  const double* __restrict__ corr_all, const uint32_t* __restrict__ seed,
                                                  ThbRelPoseResult* __restrict__ results, uint8_t* __restrict__ mask_all,
                                                  int* __restrict__ idx_ws, Model* model_ws, double* cost_ws,
                                                  int* ninl_ws, int* __restrict__ pair_counter, unsigned long long* __restrict__ stats,
                                                  uint8_t* __restrict__ lo_flags_all, const double* __restrict__ pair_thresh,
                                                  uint32_t* __restrict__ rng_state, int rng_mode, const uint8_t* __restrict__ pair_skip) {
  // pair_thresh: per-pair error_thresh (EstimateTwoViewInfo scales it by the pair's focal lengths); rng_state: 625 words per
  // pair - rng_mode 1 saves the generator after the run, 2 starts from the saved generator instead of seeding it (the
  // reference runs CountHomographyInliers and EstimateTwoViewInfo on ONE generator); pair_skip: pairs not to run
  ThbRansacParams P = P_in;
  constexpr int SS = Est::S, DD = Est::D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RansacShared& S = *reinterpret_cast<RansacShared*>(smem_raw);
  Model* M = model_ws + (size_t)blockIdx.x * BI * MAXM;
  double* g_cost = cost_ws + (size_t)blockIdx.x * BI * MAXM;  // per-model cost / inlier count of the current batch
  int* g_ninl = ninl_ws + (size_t)blockIdx.x * BI * MAXM;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const double log_failure_prob = log(P.failure_probability);
  while (true) {
  __syncthreads();  // the previous pair is completely done (shared state, final scoring)
  if (t == 0) S.pair = atomicAdd(pair_counter, 1);
  __syncthreads();
  const int pair = S.pair;
  if (pair >= num_pairs) break;
  const long long off = pair_offset[pair];
  const int n = (int)(pair_offset[pair + 1] - off);
  ThbRelPoseResult* out = results + pair;
  uint8_t* mask = mask_all ? mask_all + off : nullptr;
  if (pair_thresh) P.error_thresh = pair_thresh[pair];
  if (n < SS || (pair_skip && pair_skip[pair])) {  // RandomSampler::Initialize would CHECK-abort; reported as failure
    if (t == 0) { memset(out, 0, sizeof(*out)); out->num_input_data_points = n; }
    if (mask) for (int i = t; i < n; i += RT) mask[i] = 0;
    continue;
  }
  const double* corr = corr_all + (size_t)off * DD;
  int* sidx = idx_ws + off;  // RandomSampler::sample_indices_ (persistent permutation)
  for (int i = t; i < n; i += RT) sidx[i] = i;
  if (rng_state && rng_mode == 2) {
    const uint32_t* src = rng_state + (size_t)pair * 625;
    for (int i = t; i < 624; i += RT) S.rng.mt[i] = src[i];
    if (t == 0) S.rng.idx = (int)src[624];
  }
  if (t == 0) {
    if (!(rng_state && rng_mode == 2)) mt_seed(&S.rng, seed[pair]);
    S.best_cost = DBL_MAX;
    S.max_iterations = P.max_iterations;
    if (P.min_inlier_ratio > 0) {
      const int mi = compute_max_iterations(P, SS, P.min_inlier_ratio, log_failure_prob, n);
      S.max_iterations = mi < P.max_iterations ? mi : P.max_iterations;
    }
    S.it0 = 0; S.finished = 0; S.num_iterations = 0; S.have_best = 0;
    S.stat_samples = 0; S.stat_models = 0; S.stat_data = 0;
    S.cyc[0] = S.cyc[1] = S.cyc[2] = S.cyc[3] = 0;
    S.num_lo = 0;
    memset(&S.best, 0, sizeof(Model));
    if (P.ransac_type == 1) prosac_init(&S.prosac, n, SS);
  }
  __syncthreads();
  while (true) {
    if (S.it0 >= S.max_iterations) { if (t == 0) { S.num_iterations = S.it0; S.finished = 1; } }
    __syncthreads();
    if (S.finished) break;
    const int it0 = S.it0;
    const int nit = min(BI, S.max_iterations - it0);
    // ---- draw
    long long tick = clock64();
    if (t == 0) {
      if (P.ransac_type == 1) {
        for (int b = 0; b < nit; ++b) prosac_sample(&S.prosac, &S.rng, S.samples[b]);
      } else {
        for (int b = 0; b < nit; ++b)
          for (int i = 0; i < SS; ++i) {
            const int j = mt_uniform_int(&S.rng, i, n - 1);
            const int a = sidx[i], c = sidx[j];
            sidx[i] = c; sidx[j] = a;
            S.samples[b][i] = c;
          }
      }
    }
    __syncthreads();
    if (t == 0) { const long long now = clock64(); S.cyc[0] += now - tick; tick = now; }
    // ---- solve: thread b = iteration it0 + b
    if (t < BI) {
      int nm = 0;
      if (t < nit) {
        double sample[SS * DD];
        for (int i = 0; i < SS; ++i)
          for (int k = 0; k < DD; ++k) sample[DD * i + k] = corr[(size_t)S.samples[t][i] * DD + k];
        nm = Est::solve(sample, M + t * MAXM);  // straight into the CTA's scratch area: no 1.7 KB local copy
      }
      S.nmodels[t] = nm;
    }
    __syncthreads();
    if (t == 0) {
      int acc = 0;
      for (int b = 0; b < BI; ++b) { S.model_start[b] = acc; acc += (b < nit) ? S.nmodels[b] : 0; }
      S.model_start[BI] = acc;
      S.stat_samples += nit; S.stat_models += acc;
      const long long now = clock64(); S.cyc[1] += now - tick; tick = now;
    }
    __syncthreads();
    // ---- score: warp w takes flat models w, w + NW, ...
    const int total_models = S.model_start[BI];
    const double bail = S.best_cost;
    unsigned scored = 0;
    for (int j = w; j < total_models; j += NW) {
      int b = 0;
      while (S.model_start[b + 1] <= j) ++b;
      const int k = j - S.model_start[b];
      double cost; int ninl;
      score_model<Est>(P, corr, n, M[b * MAXM + k], bail, nullptr, &cost, &ninl, &scored);
      if (lane == 0) { g_cost[b * MAXM + k] = cost; g_ninl[b * MAXM + k] = ninl; }
    }
    if (lane == 0 && scored) atomicAdd(&S.stat_data, (unsigned long long)scored);
    __syncthreads();
    // ---- scan in (iteration, model) order
    if (t == 0) {
      { const long long now = clock64(); S.cyc[2] += now - tick; tick = now; }
      S.scan_b = 0; S.scan_k = 0; S.scan_it = it0;
    }
    if constexpr (!Est::HAS_LO) {  // plain replay: one pass of thread 0 over the batch
      if (t == 0) {
        int it = it0;
        for (int b = 0; b < nit; ++b, ++it) {
          if (it >= S.max_iterations) break;
          for (int k = 0; k < S.nmodels[b]; ++k) {
            const double sample_cost = g_cost[b * MAXM + k];
            if (sample_cost < S.best_cost) {
              const double inlier_ratio = (double)g_ninl[b * MAXM + k] / (double)n;
              S.best = M[b * MAXM + k];
              S.best_cost = sample_cost;
              S.have_best = 1;
              if (inlier_ratio < (double)SS / (double)n) continue;
              const int mi = compute_max_iterations(P, SS, inlier_ratio, log_failure_prob, n);
              if (mi < S.max_iterations) S.max_iterations = mi;
            }
          }
        }
        S.it0 = it;
      }
    } else {
    int pending = 1;  // CTA-uniform
    while (pending) {
      if (t == 0) {
        S.lo_pending = 0;
        int it = S.scan_it, b = S.scan_b, k = S.scan_k;
        while (b < nit) {
          if (k == 0 && it >= S.max_iterations) break;
          bool hit = false;
          for (; k < S.nmodels[b]; ++k) {
            const double sample_cost = g_cost[b * MAXM + k];
            if (sample_cost < S.best_cost) {
              const double inlier_ratio = (double)g_ninl[b * MAXM + k] / (double)n;
              S.best = M[b * MAXM + k];
              S.best_cost = sample_cost;
              S.have_best = 1;
              if (inlier_ratio < (double)SS / (double)n) continue;
              if (Est::HAS_LO && P.use_lo && it >= P.lo_start_iterations) {  // sample_consensus_estimator.h:372-380
                S.lo_pending = 1; S.lo_ratio = inlier_ratio; S.scan_b = b; S.scan_k = k + 1; S.scan_it = it;
                hit = true;
                break;
              }
              const int mi = compute_max_iterations(P, SS, inlier_ratio, log_failure_prob, n);
              if (mi < S.max_iterations) S.max_iterations = mi;
            }
          }
          if (hit) break;
          k = 0; ++b; ++it;
        }
        if (!S.lo_pending) S.it0 = it;
      }
      __syncthreads();
      pending = S.lo_pending;
      if (pending) {
        // inliers of the new best model (it was scored with early abandonment off: its cost is below the bail value), then RefineModel
        uint8_t* flags = lo_flags_all + off;
        if (w == 0) { double c_; int n_; unsigned sc_ = 0; score_model<Est>(P, corr, n, S.best, DBL_MAX, flags, &c_, &n_, &sc_); }
        __syncthreads();
        const bool refined = Est::refine(corr, flags, n, P.error_thresh, &S.best, S.lo);
        if (t == 0 && refined) {  // a failed refinement `continue`s: the iteration bound is not updated
          ++S.num_lo;
          const int mi = compute_max_iterations(P, SS, S.lo_ratio, log_failure_prob, n);
          if (mi < S.max_iterations) S.max_iterations = mi;
        }
      }
      __syncthreads();  // S.lo_pending is rewritten by thread 0 at the top of the next round
    }
    }
    if (t == 0) S.cyc[3] += clock64() - tick;
    __syncthreads();
  }
  // ---- final inliers of the best model (sample_consensus_estimator.h:396-414)
  uint8_t* fmask = mask ? mask : (lo_flags_all ? lo_flags_all + off : nullptr);
  double f_cost = 0.0; int f_ninl = 0;
  unsigned f_scored = 0;
  if (w == 0) score_model<Est>(P, corr, n, S.best, DBL_MAX, fmask, &f_cost, &f_ninl, &f_scored);
  if (Est::HAS_LO && P.use_lo) {  // :400-405 - the summary's inliers are those of the model BEFORE this last refinement
    __syncthreads();
    Est::refine(corr, fmask, n, P.error_thresh, &S.best, S.lo);
    if (t == 0) ++S.num_lo;
    __syncthreads();
  }
  if (w == 0 && lane == 0) {
    if (stats) {
      atomicAdd(stats + 0, 1ull); atomicAdd(stats + 1, (unsigned long long)S.num_iterations); atomicAdd(stats + 2, S.stat_samples);
      atomicAdd(stats + 3, S.stat_models + 1); atomicAdd(stats + 4, S.stat_data + f_scored);
      for (int k = 0; k < 4; ++k) atomicAdd(stats + 6 + k, (unsigned long long)S.cyc[k]);
    }
    out->success = 1;
    out->num_inliers = f_ninl;
    out->num_iterations = S.num_iterations;
    out->num_input_data_points = n;
    const double ratio = (double)f_ninl / (double)n;
    out->confidence = 1.0 - pow(1.0 - pow(ratio, (double)SS), (double)S.num_iterations);
    out->best_cost = S.best_cost;
    out->num_lo_iterations = S.num_lo; out->reserved0 = 0;
    for (int k = 0; k < 9; ++k) { out->essential_matrix[k] = S.best.E[k]; out->rotation[k] = S.best.R[k]; }
    for (int k = 0; k < 3; ++k) out->position[k] = S.best.p[k];
  }
  if (rng_state && rng_mode == 1) {
    // The batches draw their samples ahead of the replay: when the loop stops inside a batch the generator has run past the
    // reference's. Rebuild the state the sequential loop leaves behind: num_iterations x SampleSize draws (the draws depend on
    // the ranges only, not on the permutation).
    __syncthreads();
    if (t == 0) {
      mt_seed(&S.rng, seed[pair]);
      for (int it = 0; it < S.num_iterations; ++it)
        for (int i = 0; i < SS; ++i) (void)mt_uniform_int(&S.rng, i, n - 1);
    }
    __syncthreads();
    uint32_t* dst = rng_state + (size_t)pair * 625;
    for (int i = t; i < 624; i += RT) dst[i] = S.rng.mt[i];
    if (t == 0) dst[624] = (uint32_t)S.rng.idx;
  }
  }  // next pair
}

// ---- round-synchronous form of the same loop (no LO) ---------------------------------------------------------------
// The fused kernel above keeps a pair on ONE CTA, so a batch smaller than the grid leaves SMs idle and a single pair runs on four
// warps. Here a ROUND = the same batch of BI iterations for EVERY active pair, one kernel per phase: draw and scan run one pair
// per thread, solve one hypothesis per thread over all pairs, score one warp per (pair, iteration) - a pair's round is spread over
// 33 CTAs. The arithmetic of every hypothesis and of every model score is the code above, unchanged (same Est::solve, same
// score_model, same scan), so results stay bit-identical to the fused kernel and to the oracle.
// Measured (r02, B200, C4-shaped pairs): faster than the fused kernel up to ~1 000 pairs (1 pair 2.5 vs ~6 ms per call), 4 % slower
// at 10 000 pairs (215 vs 208 ms): the solve phase is the same thread-per-hypothesis code and takes 140 ms alone - ~1.5 M cycles
// of dependent FP64 / local-memory latency per hypothesis, throughput saturating at ~65 solves per Mcycle per SM from 4 warps up
// (THB_RS_PROF=1 prints the distribution) - and run as its own kernel it no longer overlaps the score phase of other pairs.
// Also measured and dropped: the solver's 10 x 10 work matrices in shared memory, interleaved per thread (conflict-free for its
// data-dependent indexing; 159 vs 140 ms: the latency is in the dependent arithmetic, not in those loads), 1 / 2 / 4 solver CTAs
// per SM (188-210 ms with a grid-stride loop), 8- and 16-warp score CTAs (67 / 77 vs 64 ms).
struct PairState {
  Mt19937 rng;
  ProsacState prosac;
  Model best;
  double best_cost, thresh;
  long long off;
  int n, pair, max_iterations, it0, num_iterations, have_best, nit, pad;
  unsigned long long stat_samples, stat_models, stat_data;
};

template <class Est>
__global__ void __launch_bounds__(128) k_rs_init(const ThbRansacParams P, int pair0, int count, const long long* __restrict__ pair_offset,
                                                 const uint32_t* __restrict__ seed, ThbRelPoseResult* __restrict__ results,
                                                 uint8_t* __restrict__ mask_all, int* __restrict__ idx_ws, PairState* __restrict__ states,
                                                 int* __restrict__ active, int* __restrict__ counters, const double* __restrict__ pair_thresh,
                                                 const uint32_t* __restrict__ rng_state, int rng_mode, const uint8_t* __restrict__ pair_skip) {
  const int slot = blockIdx.x, pair = pair0 + slot, t = threadIdx.x;
  if (slot >= count) return;
  const long long off = pair_offset[pair];
  const int n = (int)(pair_offset[pair + 1] - off);
  PairState& S = states[slot];
  if (n < Est::S || (pair_skip && pair_skip[pair])) {
    if (t == 0) { memset(results + pair, 0, sizeof(ThbRelPoseResult)); results[pair].num_input_data_points = n; S.n = -1; S.pair = pair; }
    if (mask_all) for (int i = t; i < n; i += blockDim.x) mask_all[off + i] = 0;
    return;
  }
  int* sidx = idx_ws + off;
  for (int i = t; i < n; i += blockDim.x) sidx[i] = i;
  if (rng_state && rng_mode == 2) {
    const uint32_t* src = rng_state + (size_t)pair * 625;
    for (int i = t; i < 624; i += blockDim.x) S.rng.mt[i] = src[i];
    if (t == 0) S.rng.idx = (int)src[624];
  }
  if (t == 0) {
    if (!(rng_state && rng_mode == 2)) mt_seed(&S.rng, seed[pair]);
    ThbRansacParams Pl = P;
    S.thresh = pair_thresh ? pair_thresh[pair] : P.error_thresh;
    S.best_cost = DBL_MAX;
    S.max_iterations = P.max_iterations;
    if (P.min_inlier_ratio > 0) {
      const int mi = compute_max_iterations(Pl, Est::S, P.min_inlier_ratio, log(P.failure_probability), n);
      S.max_iterations = mi < P.max_iterations ? mi : P.max_iterations;
    }
    S.off = off; S.n = n; S.pair = pair; S.it0 = 0; S.num_iterations = 0; S.have_best = 0; S.nit = 0;
    S.stat_samples = 0; S.stat_models = 0; S.stat_data = 0;
    memset(&S.best, 0, sizeof(Model));
    if (P.ransac_type == 1) prosac_init(&S.prosac, n, Est::S);
    atomicMax(&counters[2], n);
    if (S.it0 >= S.max_iterations) S.num_iterations = S.it0;  // finished before the first round
    else active[atomicAdd(&counters[0], 1)] = slot;
  }
}

// draw: RandomSampler for the BI iterations of this round, one pair per thread (the generator is sequential)
template <class Est>
__global__ void __launch_bounds__(64) k_rs_draw(int ransac_type, int na, const int* __restrict__ active, PairState* __restrict__ states,
                                                int* __restrict__ idx_ws, int* __restrict__ samples) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int slot = active[a];
  PairState& S = states[slot];
  const int n = S.n, nit = min(BI, S.max_iterations - S.it0);
  int* sidx = idx_ws + S.off;
  int* out = samples + (size_t)slot * BI * 5;
  if (ransac_type == 1) {
    for (int b = 0; b < nit; ++b) prosac_sample(&S.prosac, &S.rng, out + b * 5);
  } else {
    for (int b = 0; b < nit; ++b)
      for (int i = 0; i < Est::S; ++i) {
        const int j = mt_uniform_int(&S.rng, i, n - 1);
        const int u = sidx[i], c = sidx[j];
        sidx[i] = c; sidx[j] = u;
        out[b * 5 + i] = c;
      }
  }
  S.nit = nit;
  S.stat_samples += nit;
}

// The same draws with the generator staged in shared memory: the one-thread-per-pair loop above is a chain of dependent L2
// accesses (generator index, generator word, two permutation entries: ~1 900 cycles per draw measured on the C5 batch, 0.5 ms per
// round), and the generator is 2.5 KB per pair. A warp takes 16 pairs; their states are interleaved word by word with an odd
// stride (conflict-free for the lanes in lockstep and for the cooperative copy). RandomSampler only (PROSAC keeps k_rs_draw).
constexpr int DRAW_PAIRS = 16, MT_STRIDE = 17;
__device__ __forceinline__ uint32_t mt_next_strided(uint32_t* mt, int& idx) {
  if (idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (mt[i * MT_STRIDE] & 0x80000000u) | (mt[((i + 1) % 624) * MT_STRIDE] & 0x7fffffffu);
      mt[i * MT_STRIDE] = mt[((i + 397) % 624) * MT_STRIDE] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    idx = 0;
  }
  uint32_t y = mt[(idx++) * MT_STRIDE];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__device__ __forceinline__ int mt_uniform_int_strided(uint32_t* mt, int& idx, int lo, int hi) {  // = mt_uniform_int
  const uint32_t range = (uint32_t)(hi - lo) + 1u;
  uint64_t product = (uint64_t)mt_next_strided(mt, idx) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next_strided(mt, idx) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return lo + (int)(product >> 32);
}
// STAGE_IDX: the pairs' index permutations (RandomSampler's persistent shuffle, n ints per pair) are staged as well when 16 of them
// fit next to the generators; the drawing lane then writes its swaps through to global memory (fire and forget) and nothing is
// copied back but the generator states. Otherwise the permutation entries are read from global memory (two L2 trips per draw).
__device__ long long g_draw_prof[4];  // THB_TV_TIMING: cycles of block 0 (copy in, draws, copy out) of the last draw launch
template <class Est, bool STAGE_IDX>
__global__ void __launch_bounds__(256) k_rs_draw_staged(int na, const int* __restrict__ active, PairState* __restrict__ states,
                                                        int* __restrict__ idx_ws, int* __restrict__ samples, int idx_stride) {
  extern __shared__ uint32_t draw_smem[];
  uint32_t* st = draw_smem;                                            // [624][MT_STRIDE]
  int* sx = reinterpret_cast<int*>(draw_smem + 624 * MT_STRIDE);       // [DRAW_PAIRS][idx_stride]
  const int t = threadIdx.x, a0 = blockIdx.x * DRAW_PAIRS;
  const int cnt = min(DRAW_PAIRS, na - a0);
  const long long c0 = clock64();
  for (int p = 0; p < cnt; ++p) {
    const PairState& Sp = states[active[a0 + p]];
    for (int i = t; i < 624; i += blockDim.x) st[i * MT_STRIDE + p] = Sp.rng.mt[i];
    if (STAGE_IDX) {
      const int* src = idx_ws + Sp.off;
      for (int i = t; i < Sp.n; i += blockDim.x) sx[p * idx_stride + i] = src[i];
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (t < cnt) {
    const int slot = active[a0 + t];
    PairState& S = states[slot];
    const int n = S.n, nit = min(BI, S.max_iterations - S.it0);
    int* sidx = idx_ws + S.off;
    int* sl = sx + t * idx_stride;
    int* out = samples + (size_t)slot * BI * 5;
    uint32_t* mt = st + t;
    int idx = S.rng.idx;
    for (int b = 0; b < nit; ++b)
      for (int i = 0; i < Est::S; ++i) {
        const int j = mt_uniform_int_strided(mt, idx, i, n - 1);
        int u, c;
        if (STAGE_IDX) { u = sl[i]; c = sl[j]; sl[i] = c; sl[j] = u; }
        else { u = sidx[i]; c = sidx[j]; }
        sidx[i] = c; sidx[j] = u;
        out[b * 5 + i] = c;
      }
    S.rng.idx = idx;
    S.nit = nit;
    S.stat_samples += nit;
  }
  __syncthreads();
  const long long c2 = clock64();
  for (int p = 0; p < cnt; ++p) {
    uint32_t* dst = states[active[a0 + p]].rng.mt;
    for (int i = t; i < 624; i += blockDim.x) dst[i] = st[i * MT_STRIDE + p];
  }
  if (blockIdx.x == 0 && t == 0) { g_draw_prof[0] = c1 - c0; g_draw_prof[1] = c2 - c1; g_draw_prof[2] = clock64() - c2; }
}

// solve: thread (a, b) = hypothesis b of active pair a
template <class Est>
__global__ void __launch_bounds__(RT, Est::SOLVE_CTAS) k_rs_solve(int na, const int* __restrict__ active, const PairState* __restrict__ states,
                                                                      const double* __restrict__ corr_all, const int* __restrict__ samples,
                                                                      Model* __restrict__ model_ws, int* __restrict__ nmodels, long long* __restrict__ prof) {
  constexpr int SS = Est::S, DD = Est::D;
  const int b = threadIdx.x;
  for (int a = blockIdx.x; a < na; a += gridDim.x) {
    const long long t0 = prof ? clock64() : 0;  // the grid bounds the solver threads per SM (their work matrices live in L1)
    const int slot = active[a];
    const PairState& S = states[slot];
    int nm = 0;
    if (b < S.nit) {
      const double* corr = corr_all + (size_t)S.off * DD;
      const int* smp = samples + ((size_t)slot * BI + b) * 5;
      double sample[SS * DD];
      for (int i = 0; i < SS; ++i)
        for (int k = 0; k < DD; ++k) sample[DD * i + k] = corr[(size_t)smp[i] * DD + k];
      nm = Est::solve(sample, model_ws + ((size_t)slot * BI + b) * MAXM);
    }
    nmodels[(size_t)slot * BI + b] = nm;
    if (prof) prof[(size_t)a * BI + b] = clock64() - t0;
  }
}

// score: warp (a, b) scores the models of hypothesis b of active pair a; consecutive warps share the pair's correspondences
template <class Est, int WARPS, int CTAS, bool LMED = false>
__global__ void __launch_bounds__(32 * WARPS, CTAS) k_rs_score(const ThbRansacParams P, const int* __restrict__ active, PairState* __restrict__ states,
                                                     const double* __restrict__ corr_all, const Model* __restrict__ model_ws,
                                                     const int* __restrict__ nmodels, double* __restrict__ cost_ws, int* __restrict__ ninl_ws) {
  const int slot = active[blockIdx.x / (BI / WARPS)], w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = (blockIdx.x % (BI / WARPS)) * WARPS + w;
  PairState& S = states[slot];
  const int nm = nmodels[(size_t)slot * BI + b];
  if (nm == 0) return;
  ThbRansacParams Pl = P;
  Pl.error_thresh = S.thresh;
  const double* corr = corr_all + (size_t)S.off * Est::D;
  const double bail = S.best_cost;
  unsigned scored = 0;
  for (int k = 0; k < nm; ++k) {
    const size_t m = ((size_t)slot * BI + b) * MAXM + k;
    double cost; int ninl;
    if constexpr (LMED) {
      __shared__ unsigned hist[WARPS * 256];
      score_model_lmed<Est>(corr, S.n, model_ws[m], hist + w * 256, nullptr, &cost, &ninl, &scored);
    } else {
      score_model<Est>(Pl, corr, S.n, model_ws[m], bail, nullptr, &cost, &ninl, &scored);
    }
    if (lane == 0) { cost_ws[m] = cost; ninl_ws[m] = ninl; }
  }
  if (lane == 0) { atomicAdd(&S.stat_data, (unsigned long long)scored); atomicAdd(&S.stat_models, (unsigned long long)nm); }
}

// scan: the sequential replay of the round in (iteration, model) order, one pair per thread; pairs that go on are appended to
// the next active list
template <class Est>
__global__ void __launch_bounds__(64) k_rs_scan(const ThbRansacParams P, int na, const int* __restrict__ active, PairState* __restrict__ states,
                                                const Model* __restrict__ model_ws, const int* __restrict__ nmodels,
                                                const double* __restrict__ cost_ws, const int* __restrict__ ninl_ws,
                                                int* __restrict__ next_active, int* __restrict__ next_count) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= na) return;
  const int slot = active[a];
  PairState& S = states[slot];
  ThbRansacParams Pl = P;
  Pl.error_thresh = S.thresh;
  const double log_failure_prob = log(P.failure_probability);
  const int n = S.n, nit = S.nit;
  int it = S.it0, max_it = S.max_iterations;
  double best_cost = S.best_cost;
  for (int b = 0; b < nit; ++b, ++it) {
    if (it >= max_it) break;
    const int nm = nmodels[(size_t)slot * BI + b];
    for (int k = 0; k < nm; ++k) {
      const size_t m = ((size_t)slot * BI + b) * MAXM + k;
      const double sample_cost = cost_ws[m];
      if (sample_cost < best_cost) {
        const double inlier_ratio = (double)ninl_ws[m] / (double)n;
        S.best = model_ws[m];
        best_cost = sample_cost;
        S.have_best = 1;
        if (inlier_ratio < (double)Est::S / (double)n) continue;
        const int mi = compute_max_iterations(Pl, Est::S, inlier_ratio, log_failure_prob, n);
        if (mi < max_it) max_it = mi;
      }
    }
  }
  S.best_cost = best_cost; S.max_iterations = max_it; S.it0 = it;
  if (it >= max_it) S.num_iterations = it;
  else next_active[atomicAdd(next_count, 1)] = slot;
}

// final inliers of the best model and the result record (sample_consensus_estimator.h:396-414), one warp per pair
template <class Est, bool LMED = false>
__global__ void __launch_bounds__(128) k_rs_final(const ThbRansacParams P, int count, PairState* __restrict__ states, const double* __restrict__ corr_all,
                                                  const uint32_t* __restrict__ seed, ThbRelPoseResult* __restrict__ results,
                                                  uint8_t* __restrict__ mask_all, unsigned long long* __restrict__ stats,
                                                  uint32_t* __restrict__ rng_state, int rng_mode) {
  const int slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (slot >= count) return;
  PairState& S = states[slot];
  if (S.n < 0) return;  // rejected by k_rs_init
  ThbRansacParams Pl = P;
  Pl.error_thresh = S.thresh;
  const double* corr = corr_all + (size_t)S.off * Est::D;
  double f_cost = 0.0; int f_ninl = 0; unsigned f_scored = 0;
  if constexpr (LMED) {
    __shared__ unsigned hist[4 * 256];
    score_model_lmed<Est>(corr, S.n, S.best, hist + (threadIdx.x >> 5) * 256, mask_all ? mask_all + S.off : nullptr, &f_cost, &f_ninl, &f_scored);
  } else {
    score_model<Est>(Pl, corr, S.n, S.best, DBL_MAX, mask_all ? mask_all + S.off : nullptr, &f_cost, &f_ninl, &f_scored);
  }
  if (lane == 0) {
    ThbRelPoseResult* out = results + S.pair;
    if (stats) {
      atomicAdd(stats + 0, 1ull); atomicAdd(stats + 1, (unsigned long long)S.num_iterations); atomicAdd(stats + 2, S.stat_samples);
      atomicAdd(stats + 3, S.stat_models + 1); atomicAdd(stats + 4, S.stat_data + f_scored);
    }
    out->success = 1;
    out->num_inliers = f_ninl;
    out->num_iterations = S.num_iterations;
    out->num_input_data_points = S.n;
    const double ratio = (double)f_ninl / (double)S.n;
    out->confidence = 1.0 - pow(1.0 - pow(ratio, (double)Est::S), (double)S.num_iterations);
    out->best_cost = S.best_cost;
    out->num_lo_iterations = 0; out->reserved0 = 0;
    for (int k = 0; k < 9; ++k) { out->essential_matrix[k] = S.best.E[k]; out->rotation[k] = S.best.R[k]; }
    for (int k = 0; k < 3; ++k) out->position[k] = S.best.p[k];
    if (rng_state && rng_mode == 1) {  // the generator as the sequential loop leaves it: num_iterations x SampleSize draws from the seed
      mt_seed(&S.rng, seed[S.pair]);
      for (int it = 0; it < S.num_iterations; ++it)
        for (int i = 0; i < Est::S; ++i) (void)mt_uniform_int(&S.rng, i, S.n - 1);
      uint32_t* dst = rng_state + (size_t)S.pair * 625;
      for (int i = 0; i < 624; ++i) dst[i] = S.rng.mt[i];
      dst[624] = (uint32_t)S.rng.idx;
    }
  }
}

__global__ void k_five_point(const double* __restrict__ x1, const double* __restrict__ x2, int count, double* __restrict__ E_out,
                             int* __restrict__ nsol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double a[10], b[10], Es[90];
  for (int k = 0; k < 10; ++k) { a[k] = x1[10 * (size_t)i + k]; b[k] = x2[10 * (size_t)i + k]; }
  for (int k = 0; k < 90; ++k) Es[k] = 0.0;
  const int n = five_point(a, b, Es);
  nsol[i] = n;
  for (int k = 0; k < 90; ++k) E_out[90 * (size_t)i + k] = Es[k];
}

// one CTA per pair: 32 flags -> one word by warp ballot
__global__ void k_pack_masks(const uint8_t* __restrict__ mask, const long long* __restrict__ pair_offset,
                             const long long* __restrict__ word_offset, uint32_t* __restrict__ words) {
  const int p = blockIdx.x;
  const long long off = pair_offset[p];
  const int n = (int)(pair_offset[p + 1] - off);
  uint32_t* out = words + word_offset[p];
  const int lane = threadIdx.x & 31;
  for (int base = (threadIdx.x >> 5) * 32; base < n; base += blockDim.x) {
    const int i = base + lane;
    const unsigned bits = __ballot_sync(0xffffffffu, i < n && mask[off + i] != 0);
    if (lane == 0) out[base >> 5] = bits;
  }
}

// FP64 issue-rate probe: 8 independent DFMA chains per thread
__global__ void k_dfma_peak(int iters, double m, double* sink) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double c = 1e-9;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) *sink = s;
}

struct Bufs {  // scoped, stream-ordered device allocations (no cudaMalloc / cudaFree per batch after the first)
  std::vector<void*> p;
  cudaStream_t st = nullptr;
  ~Bufs() { for (void* q : p) cudaFreeAsync(q, st); }
  template <typename T> T* get(size_t n) {
    void* q = nullptr;
    if (cudaMallocAsync(&q, (n ? n : 1) * sizeof(T), st) != cudaSuccess) return nullptr;
    p.push_back(q);
    return (T*)q;
  }
};

thread_local ThbRansacStats g_last_stats = {};

int check_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) THB_FAIL(THB_E_NO_DEVICE, "no CUDA device visible; libtheia_b200 has no CPU path");
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  THB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) THB_FAIL(THB_E_NO_DEVICE, "device is not sm_100 (B200); kernels are built for sm_100a only");
  static std::once_flag pool_once[64];  // keep freed blocks in the default pool (same policy as the BA path), per device
  if (dev >= 0 && dev < 64)
    std::call_once(pool_once[dev], [dev] {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    });
  return THB_OK;
}


// The device part of a batch: scratch from the stream's pool, one persistent launch. Everything is a device pointer; no
// synchronisation (the caller's stream order is the only dependency), so pipelines chain it with other kernels.
// Round-synchronous driver (k_rs_*): chunks of pairs, per round one launch per phase over the chunk's active pairs. The host
// reads the number of active pairs after every round (4 bytes), so unlike the fused kernel this path synchronises the stream.
template <class Est>
int launch_ransac_rounds(cudaStream_t st, Bufs& B, const ThbRansacParams& p, int np, const long long* d_off, const double* d_corr,
                         const uint32_t* d_seed, long long total, ThbRelPoseResult* d_res, uint8_t* d_mask, const double* d_thresh,
                         uint32_t* d_rng, int rng_mode, const uint8_t* d_skip) {
  constexpr int kChunk = 4096;  // (16384 measured: 10 000 pairs in one chunk 212.5 vs 212.1 ms in three - no gain, 4x the scratch) bounds the per-hypothesis scratch; equal chunks: the last rounds of a chunk run few pairs, a short last chunk would be all tail
  const int nchunks = (np + kChunk - 1) / kChunk;
  const int C = (np + nchunks - 1) / nchunks;
  int* d_idx = B.get<int>((size_t)total);
  PairState* d_states = B.get<PairState>((size_t)C);
  int* d_samples = B.get<int>((size_t)C * BI * 5);
  Model* d_models = B.get<Model>((size_t)C * BI * MAXM);
  double* d_cost = B.get<double>((size_t)C * BI * MAXM);
  int* d_ninl = B.get<int>((size_t)C * BI * MAXM);
  int* d_nm = B.get<int>((size_t)C * BI);
  int* d_active = B.get<int>((size_t)2 * C);
  int* d_count = B.get<int>(3);  // active pairs of the current / next round, largest n of the chunk
  unsigned long long* d_stats = B.get<unsigned long long>(10);
  constexpr size_t kDrawStateBytes = sizeof(uint32_t) * 624 * MT_STRIDE;
  if (!d_idx || !d_states || !d_samples || !d_models || !d_cost || !d_ninl || !d_nm || !d_active || !d_count || !d_stats) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemsetAsync(d_stats, 0, sizeof(unsigned long long) * 10, st));
  cudaEvent_t ev[5];
  for (auto& e : ev) THB_CUDA_CHECK(cudaEventCreate(&e));
  double phase_ms[4] = {0, 0, 0, 0};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  long long* d_prof = getenv("THB_RS_PROF") ? B.get<long long>((size_t)C * BI) : nullptr;
  int rc = THB_OK, num_rounds = 0;
  for (int pair0 = 0; pair0 < np && rc == THB_OK; pair0 += C) {
    const int count = std::min(C, np - pair0);
    cudaMemsetAsync(d_count, 0, sizeof(int) * 3, st);
    k_rs_init<Est><<<count, 128, 0, st>>>(p, pair0, count, d_off, d_seed, d_res, d_mask, d_idx, d_states, d_active, d_count, d_thresh, d_rng, rng_mode, d_skip);
    int na = 0, cur = 0, h_cnt[3] = {0, 0, 0};
    if (cudaMemcpyAsync(h_cnt, d_count, sizeof(int) * 3, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = THB_E_CUDA; break; }
    na = h_cnt[0];
    // the sampler's permutations of 16 pairs next to their generators in shared memory, if the largest pair of the chunk allows it
    const int idx_stride = h_cnt[2] | 1;
    const size_t draw_smem = kDrawStateBytes + sizeof(int) * (size_t)DRAW_PAIRS * idx_stride;
    const bool stage_idx = draw_smem <= 200 * 1024;
    if (stage_idx && cudaFuncSetAttribute(k_rs_draw_staged<Est, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)draw_smem) != cudaSuccess) { rc = THB_E_CUDA; break; }
    while (na > 0) {
      int* act = d_active + (size_t)cur * C;
      int* nxt = d_active + (size_t)(1 - cur) * C;
      cudaEventRecord(ev[0], st);
      if (p.ransac_type == 1) k_rs_draw<Est><<<(na + 63) / 64, 64, 0, st>>>(p.ransac_type, na, act, d_states, d_idx, d_samples);
      else if (stage_idx) k_rs_draw_staged<Est, true><<<(na + DRAW_PAIRS - 1) / DRAW_PAIRS, 256, draw_smem, st>>>(na, act, d_states, d_idx, d_samples, idx_stride);
      else k_rs_draw_staged<Est, false><<<(na + DRAW_PAIRS - 1) / DRAW_PAIRS, 32, kDrawStateBytes, st>>>(na, act, d_states, d_idx, d_samples, 0);
      cudaEventRecord(ev[1], st);
      k_rs_solve<Est><<<na, RT, 0, st>>>(na, act, d_states, d_corr, d_samples, d_models, d_nm, d_prof);
      if (d_prof) {  // THB_RS_PROF: distribution of the per-hypothesis solve time (SM cycles)
        std::vector<long long> h((size_t)na * BI);
        cudaMemcpyAsync(h.data(), d_prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        std::vector<long long> v(h), cta_max;
        for (int a2 = 0; a2 < na; ++a2) cta_max.push_back(*std::max_element(h.begin() + (size_t)a2 * BI, h.begin() + (size_t)(a2 + 1) * BI));
        std::sort(v.begin(), v.end()); std::sort(cta_max.begin(), cta_max.end());
        double mean = 0; for (long long x : v) mean += (double)x; mean /= (double)v.size();
        double cmean = 0; for (long long x : cta_max) cmean += (double)x; cmean /= (double)cta_max.size();
        fprintf(stderr, "RSPROF na=%d threads: mean %.0f p50 %lld p90 %lld p99 %lld max %lld | cta max: mean %.0f p50 %lld max %lld\n", na, mean, v[v.size() / 2],
                v[v.size() * 9 / 10], v[v.size() * 99 / 100], v.back(), cmean, cta_max[cta_max.size() / 2], cta_max.back());
      }
      cudaEventRecord(ev[2], st);
      if (p.ransac_type == 2) k_rs_score<Est, 4, 6, true><<<na * (BI / 4), 128, 0, st>>>(p, act, d_states, d_corr, d_models, d_nm, d_cost, d_ninl);
      else k_rs_score<Est, 4, 6><<<na * (BI / 4), 128, 0, st>>>(p, act, d_states, d_corr, d_models, d_nm, d_cost, d_ninl);
      cudaEventRecord(ev[3], st);
      cudaMemsetAsync(d_count + (1 - cur), 0, sizeof(int), st);
      k_rs_scan<Est><<<(na + 63) / 64, 64, 0, st>>>(p, na, act, d_states, d_models, d_nm, d_cost, d_ninl, nxt, d_count + (1 - cur));
      cudaEventRecord(ev[4], st);
      if (cudaMemcpyAsync(&na, d_count + (1 - cur), sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = THB_E_CUDA; break; }
      for (int k = 0; k < 4; ++k) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) == cudaSuccess) phase_ms[k] += ms; }
      cur = 1 - cur;
      ++num_rounds;
    }
    if (rc != THB_OK) break;
    if (p.ransac_type == 2) k_rs_final<Est, true><<<(count + 3) / 4, 128, 0, st>>>(p, count, d_states, d_corr, d_seed, d_res, d_mask, d_stats, d_rng, rng_mode);
    else k_rs_final<Est><<<(count + 3) / 4, 128, 0, st>>>(p, count, d_states, d_corr, d_seed, d_res, d_mask, d_stats, d_rng, rng_mode);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc != THB_OK) THB_FAIL(rc, "round-synchronous RANSAC: CUDA error");
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(&g_last_stats, d_stats, sizeof(ThbRansacStats), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (getenv("THB_TV_TIMING")) {
    long long dp[4] = {0, 0, 0, 0};
    cudaMemcpyFromSymbol(dp, g_draw_prof, sizeof(dp));
    fprintf(stderr, "  last draw launch, block 0: copy in %lld, draws %lld, copy out %lld cycles\n", dp[0], dp[1], dp[2]);
  }
  if (getenv("THB_TV_TIMING"))
    fprintf(stderr, "  rounds<S=%d>: %d pairs, %d rounds, draw %.2f solve %.2f score %.2f scan %.2f ms\n", Est::S, np, num_rounds, phase_ms[0], phase_ms[1], phase_ms[2], phase_ms[3]);
  // phase shares: device time of the phase kernels in nanoseconds (the fused kernel reports CTA cycles in the same fields)
  g_last_stats.cycles_draw = (uint64_t)(phase_ms[0] * 1e6); g_last_stats.cycles_solve = (uint64_t)(phase_ms[1] * 1e6);
  g_last_stats.cycles_score = (uint64_t)(phase_ms[2] * 1e6); g_last_stats.cycles_scan = (uint64_t)(phase_ms[3] * 1e6);
  return THB_OK;
}

template <class Est>
int launch_ransac(cudaStream_t st, Bufs& B, const ThbRansacParams& p, int np, const long long* d_off, const double* d_corr,
                  const uint32_t* d_seed, long long total, ThbRelPoseResult* d_res, uint8_t* d_mask, const double* d_thresh,
                  uint32_t* d_rng, int rng_mode, const uint8_t* d_skip) {
  {  // Small batches cannot fill the GPU with one CTA per pair: the round-synchronous kernels spread a pair over many CTAs
     // (measured at the end of r02, C4-shaped pairs, rounds vs fused: 1 pair 2.5 vs ~6 ms, 16 pairs 4.4 vs 7.1 ms, 148 pairs 7.3 vs 9.4 ms,
     // 1 250 pairs 28.1 vs 32.9 ms, 2 500 pairs 52.5 vs 57.1 ms, 5 000 pairs 107.2 vs 108.9 ms; 10 000 pairs 212.1 vs 208.8 ms - there the
     // fused kernel's mix of phases per SM wins). THB_RANSAC_MODE=fused|rounds forces one.
    const char* m = getenv("THB_RANSAC_MODE");  // read per call: tests switch it
    const int mode = !m ? 0 : std::string(m) == "fused" ? 1 : std::string(m) == "rounds" ? 2 : 0;
    const bool rounds = mode == 2 || (mode == 0 && np <= 6000) || p.ransac_type == 2;  // LMED scoring lives in k_rs_score only
    if (!(p.use_lo && Est::HAS_LO) && rounds)
      return launch_ransac_rounds<Est>(st, B, p, np, d_off, d_corr, d_seed, total, d_res, d_mask, d_thresh, d_rng, rng_mode, d_skip);
  }
  int* d_idx = B.get<int>((size_t)total);
  if (!d_idx) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  // Shared memory holds the control block only. Staging the pair's correspondences there was measured 20 % slower on C4
  // (41.7k vs 50.5k pairs/s): three staged copies take 207 KB of the SM's 256 KB L1/shared array and the five-point
  // solver's per-thread work matrices (local memory) then miss L1; a pair's 64 KB stays L1-resident between models anyway.
  const size_t want = ((sizeof(RansacShared) + 31) / 32) * 32;
  THB_CUDA_CHECK(cudaFuncSetAttribute(k_ransac<Est>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int grid = std::min(np, RANSAC_CTAS_PER_SM * sms);
  Model* d_models = B.get<Model>((size_t)grid * BI * MAXM);
  double* d_cost = B.get<double>((size_t)grid * BI * MAXM);
  int* d_ninl = B.get<int>((size_t)grid * BI * MAXM);
  int* d_counter = B.get<int>(1);
  unsigned long long* d_stats = B.get<unsigned long long>(10);
  uint8_t* d_lo_flags = (p.use_lo && Est::HAS_LO) ? B.get<uint8_t>((size_t)total) : nullptr;
  if (!d_models || !d_cost || !d_ninl || !d_counter || !d_stats || (p.use_lo && Est::HAS_LO && !d_lo_flags)) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemsetAsync(d_counter, 0, sizeof(int), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_stats, 0, sizeof(unsigned long long) * 10, st));
  k_ransac<Est><<<grid, RT, want, st>>>(p, np, d_off, d_corr, d_seed, d_res, d_mask, d_idx, d_models, d_cost, d_ninl, d_counter, d_stats, d_lo_flags,
                                        d_thresh, d_rng, rng_mode, d_skip);
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(&g_last_stats, d_stats, sizeof(ThbRansacStats), cudaMemcpyDeviceToHost, st));
  return THB_OK;
}

template <class Est>
int run_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* inlier_mask, void* cuda_stream) {
  if (!b || !p || !results) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  // the reference CHECK-aborts on these (sample_consensus_estimator.h:217-223)
  if (!(p->error_thresh > 0) || !(p->failure_probability > 0 && p->failure_probability < 1) || p->min_inlier_ratio < 0 ||
      p->min_inlier_ratio > 1 || p->max_iterations < p->min_iterations) THB_FAIL(THB_E_INVALID_ARGUMENT, "invalid RansacParameters");
  if (p->use_lo && !Est::HAS_LO) {
    if (Est::S == 3) THB_FAIL(THB_E_UNSUPPORTED, "use_lo for the absolute-pose estimator (BundleAdjustView on a one-view reconstruction) is not implemented");
    // HomographyEstimator has no RefineModel: the reference's LO branch is a `continue` that only skips the iteration-bound update
    THB_FAIL(THB_E_UNSUPPORTED, "use_lo is only implemented for the relative-pose estimator");
  }
  if (p->ransac_type < 0 || p->ransac_type > 2) THB_FAIL(THB_E_UNSUPPORTED, "RansacType::RANSAC, PROSAC and LMED are implemented (EXHAUSTIVE is not)");
  if (p->ransac_type == 2 && p->use_lo) THB_FAIL(THB_E_UNSUPPORTED, "use_lo with RansacType::LMED is not implemented");
  if (b->num_pairs < 0 || (b->memory_space != THB_MEM_HOST && b->memory_space != THB_MEM_DEVICE)) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad batch");
  if (b->num_pairs == 0) return THB_OK;
  if (!b->pair_offset || !b->seed) THB_FAIL(THB_E_INVALID_ARGUMENT, "null batch array");
  int rc = check_device();
  if (rc != THB_OK) return rc;
  constexpr int DD = Est::D;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int np = b->num_pairs;
  const bool host = b->memory_space == THB_MEM_HOST;
  std::vector<long long> h_off(np + 1);
  if (host) std::memcpy(h_off.data(), b->pair_offset, sizeof(long long) * (np + 1));
  else THB_CUDA_CHECK(cudaMemcpy(h_off.data(), b->pair_offset, sizeof(long long) * (np + 1), cudaMemcpyDeviceToHost));
  int max_n = 0;
  for (int i = 0; i < np; ++i) {
    const long long n = h_off[i + 1] - h_off[i];
    if (n < 0 || n > 2147483647LL || h_off[0] != 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "pair_offset must start at 0 and be non-decreasing");
    if ((int)n > max_n) max_n = (int)n;
  }
  const long long total = h_off[np];
  if (total > 0 && !b->corr) THB_FAIL(THB_E_INVALID_ARGUMENT, "null corr");
  Bufs B;
  B.st = st;
  const long long* d_off; const double* d_corr; const uint32_t* d_seed; ThbRelPoseResult* d_res; uint8_t* d_mask = nullptr;
  if (host) {
    long long* o = B.get<long long>(np + 1); double* c = B.get<double>((size_t)total * DD); uint32_t* s = B.get<uint32_t>(np);
    d_res = B.get<ThbRelPoseResult>(np);
    if (inlier_mask) d_mask = B.get<uint8_t>((size_t)total);
    if (!o || !c || !s || !d_res || (inlier_mask && !d_mask)) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
    THB_CUDA_CHECK(cudaMemcpyAsync(o, b->pair_offset, sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(c, b->corr, sizeof(double) * DD * total, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(s, b->seed, sizeof(uint32_t) * np, cudaMemcpyHostToDevice, st));
    d_off = o; d_corr = c; d_seed = s;
  } else {
    d_off = (const long long*)b->pair_offset; d_corr = b->corr; d_seed = b->seed; d_res = results; d_mask = inlier_mask;
  }
  rc = launch_ransac<Est>(st, B, *p, np, d_off, d_corr, d_seed, total, d_res, d_mask, nullptr, nullptr, 0, nullptr);
  if (rc != THB_OK) return rc;
  if (host) {
    THB_CUDA_CHECK(cudaMemcpyAsync(results, d_res, sizeof(ThbRelPoseResult) * np, cudaMemcpyDeviceToHost, st));
    if (inlier_mask) THB_CUDA_CHECK(cudaMemcpyAsync(inlier_mask, d_mask, (size_t)total, cudaMemcpyDeviceToHost, st));
  }
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

}  // namespace
}  // namespace thb
#include "two_view.cuh"
namespace thb {
namespace {

// EstimateTwoViewInfo / VerifyMatches for a batch (include/theia_b200.h). verify = false: stages normalise -> RANSAC -> info.
int run_two_view(const ThbPairBatch* b, const ThbViewIntrinsics* i1, const ThbViewIntrinsics* i2, const ThbTwoViewOptions* O,
                 ThbTwoViewInfo* info, uint8_t* out_mask, void* cuda_stream, bool verify) {
  if (!b || !i1 || !i2 || !O || !info) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  if (O->ransac_type != 0) THB_FAIL(THB_E_UNSUPPORTED, "only RansacType::RANSAC is implemented");
  if (!(O->max_sampson_error_pixels > 0) || !(O->expected_ransac_confidence > 0 && O->expected_ransac_confidence < 1) ||
      O->max_ransac_iterations < O->min_ransac_iterations) THB_FAIL(THB_E_INVALID_ARGUMENT, "invalid two-view options");
  if (b->num_pairs < 0 || (b->memory_space != THB_MEM_HOST && b->memory_space != THB_MEM_DEVICE)) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad batch");
  if (b->num_pairs == 0) return THB_OK;
  if (!b->pair_offset || !b->seed) THB_FAIL(THB_E_INVALID_ARGUMENT, "null batch array");
  int rc = check_device();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int np = b->num_pairs;
  const bool host = b->memory_space == THB_MEM_HOST;
  std::vector<long long> h_off(np + 1);
  if (host) std::memcpy(h_off.data(), b->pair_offset, sizeof(long long) * (np + 1));
  else THB_CUDA_CHECK(cudaMemcpy(h_off.data(), b->pair_offset, sizeof(long long) * (np + 1), cudaMemcpyDeviceToHost));
  for (int i = 0; i < np; ++i)
    if (h_off[i + 1] < h_off[i] || h_off[i + 1] - h_off[i] > 2147483647LL || h_off[0] != 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "pair_offset must start at 0 and be non-decreasing");
  const long long total = h_off[np];
  if (total > 0 && !b->corr) THB_FAIL(THB_E_INVALID_ARGUMENT, "null corr");
  Bufs B;
  B.st = st;
  const long long* d_off; const double* d_px; const uint32_t* d_seed; const ThbViewIntrinsics *d_i1, *d_i2; ThbTwoViewInfo* d_info; uint8_t* d_out = nullptr;
  if (host) {
    long long* o = B.get<long long>(np + 1); double* c = B.get<double>((size_t)total * 4); uint32_t* s = B.get<uint32_t>(np);
    ThbViewIntrinsics* a1 = B.get<ThbViewIntrinsics>(np); ThbViewIntrinsics* a2 = B.get<ThbViewIntrinsics>(np);
    d_info = B.get<ThbTwoViewInfo>(np);
    d_out = B.get<uint8_t>((size_t)total);
    if (!o || !c || !s || !a1 || !a2 || !d_info || !d_out) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
    THB_CUDA_CHECK(cudaMemcpyAsync(o, b->pair_offset, sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(c, b->corr, sizeof(double) * 4 * total, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(s, b->seed, sizeof(uint32_t) * np, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(a1, i1, sizeof(ThbViewIntrinsics) * np, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(a2, i2, sizeof(ThbViewIntrinsics) * np, cudaMemcpyHostToDevice, st));
    d_off = o; d_px = c; d_seed = s; d_i1 = a1; d_i2 = a2;
  } else {
    d_off = (const long long*)b->pair_offset; d_px = b->corr; d_seed = b->seed; d_i1 = i1; d_i2 = i2; d_info = info;
    d_out = out_mask ? out_mask : B.get<uint8_t>((size_t)total);
    if (!d_out) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  }
  double* d_thresh = B.get<double>(np); uint8_t* d_skip = B.get<uint8_t>(np); int* d_uncal = B.get<int>(1);
  double* d_norm = B.get<double>((size_t)total * 4);
  ThbRelPoseResult* d_res = B.get<ThbRelPoseResult>(np);
  uint8_t* d_inl = B.get<uint8_t>((size_t)total);
  if (!d_thresh || !d_skip || !d_uncal || !d_norm || !d_res || !d_inl) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemsetAsync(d_uncal, 0, sizeof(int), st));
  k_tv_prepare<<<(np + 127) / 128, 128, 0, st>>>(np, d_off, d_i1, d_i2, *O, verify ? 1 : 0, d_thresh, d_skip, d_uncal);
  int h_uncal = 0;
  THB_CUDA_CHECK(cudaMemcpyAsync(&h_uncal, d_uncal, sizeof(int), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_uncal) THB_FAIL(THB_E_UNSUPPORTED, "a pair has a view without a focal-length prior (or a camera model off the hot path): the uncalibrated branch of EstimateTwoViewInfo is not implemented");
  // THB_TV_TIMING=1: device time of the stages (homography RANSAC, relative-pose RANSAC, triangulation + two-view BA) on stderr
  const bool timing = getenv("THB_TV_TIMING") != nullptr;
  const int use_lo_for_log = O->use_lo;
  cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (timing) { for (auto& e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], st); }
  ThbRansacParams rp;
  thb_ransac_default_params(&rp);
  rp.failure_probability = 1.0 - O->expected_ransac_confidence; rp.min_iterations = O->min_ransac_iterations; rp.max_iterations = O->max_ransac_iterations;
  rp.use_mle = O->use_mle; rp.error_thresh = 1.0;
  uint32_t* d_rng = nullptr;
  ThbRelPoseResult* d_hom = nullptr;
  if (verify) {  // CountHomographyInliers (:331-366): pixel matches, threshold from the cameras before SetupCameras (image size 0)
    d_rng = B.get<uint32_t>((size_t)np * 625);
    d_hom = B.get<ThbRelPoseResult>(np);
    if (!d_rng || !d_hom) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
    ThbRansacParams hp = rp;
    hp.error_thresh = O->max_sampson_error_pixels * O->max_sampson_error_pixels;
    rc = launch_ransac<HomographyEst>(st, B, hp, np, d_off, d_px, d_seed, total, d_hom, nullptr, nullptr, d_rng, 1, d_skip);
    if (rc != THB_OK) return rc;
  }
  if (timing) cudaEventRecord(tev[1], st);
  k_tv_normalize<<<np, 128, 0, st>>>(d_off, d_px, d_i1, d_i2, d_norm);
  rp.use_lo = O->use_lo; rp.lo_start_iterations = O->lo_start_iterations;
  rc = launch_ransac<RelPoseEst>(st, B, rp, np, d_off, d_norm, d_seed, total, d_res, verify ? d_inl : d_out, d_thresh, d_rng, verify ? 2 : 0, d_skip);
  if (rc != THB_OK) return rc;
  if (timing) cudaEventRecord(tev[2], st);
  if (!verify) {
    k_tv_info<<<(np + 127) / 128, 128, 0, st>>>(np, d_res, d_i1, d_i2, d_skip, d_info);
  } else {
    double* d_p0 = B.get<double>((size_t)total * 4); double* d_p1 = B.get<double>((size_t)total * 4); double* d_ps = B.get<double>((size_t)total * 4);
    uint8_t* d_tri = B.get<uint8_t>((size_t)total); uint8_t* d_zero = B.get<uint8_t>((size_t)total + 16); int* d_two = B.get<int>(2);
    if (!d_p0 || !d_p1 || !d_ps || !d_tri || !d_zero || !d_two) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
    THB_CUDA_CHECK(cudaMemsetAsync(d_zero, 0, (size_t)total + 16, st));
    const int two[2] = {0, 1};
    THB_CUDA_CHECK(cudaMemcpyAsync(d_two, two, sizeof(two), cudaMemcpyHostToDevice, st));
    k_tv_verify<<<np, TV_THREADS, 0, st>>>(d_off, d_px, d_i1, d_i2, *O, d_res, d_inl, d_hom, d_skip, d_p0, d_p1, d_ps, d_tri, d_zero, d_two, d_info, d_out);
  }
  THB_CUDA_CHECK(cudaGetLastError());
  if (timing) {
    cudaEventRecord(tev[3], st); cudaEventSynchronize(tev[3]);
    float a = 0.f, b2 = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b2, tev[1], tev[2]); cudaEventElapsedTime(&c, tev[2], tev[3]);
    fprintf(stderr, "two-view batch of %d pairs: homography RANSAC %.2f ms, normalise + relative-pose RANSAC (LO %d) %.2f ms, info / triangulation + two-view BA %.2f ms\n",
            np, a, b2, use_lo_for_log, c);
    for (auto& e : tev) cudaEventDestroy(e);
  }
  if (host) {
    THB_CUDA_CHECK(cudaMemcpyAsync(info, d_info, sizeof(ThbTwoViewInfo) * np, cudaMemcpyDeviceToHost, st));
    if (out_mask) THB_CUDA_CHECK(cudaMemcpyAsync(out_mask, d_out, (size_t)total, cudaMemcpyDeviceToHost, st));
  }
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

// one thread per independent minimal problem: which = 0 P3P, 1 four-point homography, 2 seven-point F
template <int IN_A, int IN_B, int OUT_A, int OUT_B>
__global__ void k_minimal_solver(const double* __restrict__ a, const double* __restrict__ b, int count, double* __restrict__ oa,
                                 double* __restrict__ ob, int* __restrict__ nsol, int which) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double ia[IN_A], ib[IN_B > 0 ? IN_B : 1], ra[OUT_A], rb[OUT_B > 0 ? OUT_B : 1];
  for (int k = 0; k < IN_A; ++k) ia[k] = a[(size_t)i * IN_A + k];
  for (int k = 0; k < IN_B; ++k) ib[k] = b[(size_t)i * IN_B + k];
  for (int k = 0; k < OUT_A; ++k) ra[k] = 0.0;
  for (int k = 0; k < OUT_B; ++k) rb[k] = 0.0;
  int n = 0;
  if (which == 0) n = p3p(ia, ib, ra, rb);
  else if (which == 1) n = four_point_h(ia, ra) ? 1 : 0;
  else n = seven_point_f(ia, ra);
  nsol[i] = n;
  for (int k = 0; k < OUT_A; ++k) oa[(size_t)i * OUT_A + k] = ra[k];
  for (int k = 0; k < OUT_B; ++k) ob[(size_t)i * OUT_B + k] = rb[k];
}

template <int IN_A, int IN_B, int OUT_A, int OUT_B>
int run_solver(const double* a, const double* b, int count, double* oa, double* ob, int* nsol, void* cuda_stream, int which) {
  if (count == 0) return THB_OK;
  int rc = check_device();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Bufs B;
  B.st = st;
  double* da = B.get<double>((size_t)count * IN_A); double* db = B.get<double>((size_t)count * (IN_B > 0 ? IN_B : 1));
  double* doa = B.get<double>((size_t)count * OUT_A); double* dob = B.get<double>((size_t)count * (OUT_B > 0 ? OUT_B : 1));
  int* dn = B.get<int>(count);
  if (!da || !db || !doa || !dob || !dn) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemcpyAsync(da, a, sizeof(double) * IN_A * count, cudaMemcpyHostToDevice, st));
  if (IN_B > 0) THB_CUDA_CHECK(cudaMemcpyAsync(db, b, sizeof(double) * IN_B * count, cudaMemcpyHostToDevice, st));
  k_minimal_solver<IN_A, IN_B, OUT_A, OUT_B><<<(count + 31) / 32, 32, 0, st>>>(da, db, count, doa, dob, dn, which);
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(oa, doa, sizeof(double) * OUT_A * count, cudaMemcpyDeviceToHost, st));
  if (OUT_B > 0) THB_CUDA_CHECK(cudaMemcpyAsync(ob, dob, sizeof(double) * OUT_B * count, cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(nsol, dn, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

}  // namespace
}  // namespace thb

using namespace thb;

extern "C" {

void thb_ransac_default_params(ThbRansacParams* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  // RansacParameters() (solvers/sample_consensus_estimator.h:59-68)
  p->error_thresh = -1.0; p->failure_probability = 0.01; p->min_inlier_ratio = 0.0;
  p->min_iterations = 100; p->max_iterations = 2147483647; p->use_mle = 0; p->use_lo = 0; p->lo_start_iterations = 50;
  p->ransac_type = 0;
}

int thb_ransac_relpose_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* inlier_mask,
                             void* cuda_stream) {
  return run_batch<RelPoseEst>(b, p, results, inlier_mask, cuda_stream);
}
int thb_ransac_abspose_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* inlier_mask,
                             void* cuda_stream) {
  return run_batch<AbsPoseEst>(b, p, results, inlier_mask, cuda_stream);
}
int thb_ransac_homography_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* inlier_mask,
                                void* cuda_stream) {
  return run_batch<HomographyEst>(b, p, results, inlier_mask, cuda_stream);
}

void thb_two_view_default_options(ThbTwoViewOptions* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  // EstimateTwoViewInfoOptions (estimate_twoview_info.h:51-80), TwoViewMatchGeometricVerification::Options (.h:53-92)
  o->max_sampson_error_pixels = 6.0; o->expected_ransac_confidence = 0.9999; o->min_ransac_iterations = 10; o->max_ransac_iterations = 1000;
  o->use_mle = 1; o->use_lo = 0; o->lo_start_iterations = 10; o->ransac_type = 0;
  o->min_num_inlier_matches = 30; o->bundle_adjustment = 1; o->triangulation_max_reprojection_error = 15.0;
  o->min_triangulation_angle_degrees = 4.0; o->final_max_reprojection_error = 5.0;
}
int thb_estimate_two_view_info_batch(const ThbPairBatch* b, const ThbViewIntrinsics* i1, const ThbViewIntrinsics* i2, const ThbTwoViewOptions* o,
                                     ThbTwoViewInfo* info, uint8_t* inlier_mask, void* cuda_stream) {
  return run_two_view(b, i1, i2, o, info, inlier_mask, cuda_stream, false);
}
int thb_verify_two_view_matches_batch(const ThbPairBatch* b, const ThbViewIntrinsics* i1, const ThbViewIntrinsics* i2, const ThbTwoViewOptions* o,
                                      ThbTwoViewInfo* info, uint8_t* verified_mask, void* cuda_stream) {
  return run_two_view(b, i1, i2, o, info, verified_mask, cuda_stream, true);
}

int thb_ransac_last_stats(ThbRansacStats* stats) {
  if (!stats) THB_FAIL(THB_E_INVALID_ARGUMENT, "null stats");
  *stats = g_last_stats;
  return THB_OK;
}

int thb_pack_inlier_masks(const uint8_t* mask, const int64_t* pair_offset, const int64_t* word_offset, int32_t num_pairs,
                          int32_t memory_space, uint32_t* words, void* cuda_stream) {
  if (num_pairs < 0 || (memory_space != THB_MEM_HOST && memory_space != THB_MEM_DEVICE)) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  if (num_pairs == 0) return THB_OK;
  if (!mask || !pair_offset || !word_offset || !words) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  int rc = check_device();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (memory_space == THB_MEM_DEVICE) {
    k_pack_masks<<<num_pairs, 128, 0, st>>>(mask, (const long long*)pair_offset, (const long long*)word_offset, words);
    THB_CUDA_CHECK(cudaGetLastError());
    return THB_OK;
  }
  const long long total = pair_offset[num_pairs], nwords = word_offset[num_pairs];
  for (int i = 0; i < num_pairs; ++i)
    if (pair_offset[i + 1] < pair_offset[i] || word_offset[i + 1] - word_offset[i] < (pair_offset[i + 1] - pair_offset[i] + 31) / 32)
      THB_FAIL(THB_E_INVALID_ARGUMENT, "offsets must be non-decreasing with ceil(n/32) words per pair");
  Bufs B;
  B.st = st;
  uint8_t* dm = B.get<uint8_t>((size_t)total); long long* dpo = B.get<long long>(num_pairs + 1); long long* dwo = B.get<long long>(num_pairs + 1);
  uint32_t* dw = B.get<uint32_t>((size_t)nwords);
  if (!dm || !dpo || !dwo || !dw) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemcpyAsync(dm, mask, (size_t)total, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(dpo, pair_offset, sizeof(long long) * (num_pairs + 1), cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(dwo, word_offset, sizeof(long long) * (num_pairs + 1), cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(uint32_t) * (size_t)nwords, st));
  k_pack_masks<<<num_pairs, 128, 0, st>>>(dm, dpo, dwo, dw);
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(words, dw, sizeof(uint32_t) * (size_t)nwords, cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

int thb_fp64_peak_tflops(int32_t repeats, double* tflops, void* cuda_stream) {
  if (!tflops || repeats <= 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  int rc = check_device();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  Bufs B;
  B.st = st;
  double* sink = B.get<double>(1);
  if (!sink) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  cudaEvent_t a, b;
  THB_CUDA_CHECK(cudaEventCreate(&a)); THB_CUDA_CHECK(cudaEventCreate(&b));
  const int iters = 1 << 14, grid = sms * 4, threads = 256;
  double best = 0.0;
  for (int r = 0; r < repeats + 1; ++r) {  // first launch is the warm-up
    cudaEventRecord(a, st);
    k_dfma_peak<<<grid, threads, 0, st>>>(iters, 1.0000001, sink);
    cudaEventRecord(b, st);
    if (cudaEventSynchronize(b) != cudaSuccess) { cudaEventDestroy(a); cudaEventDestroy(b); THB_FAIL(THB_E_CUDA, "fp64 peak kernel failed"); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * 8.0 * (double)iters * threads * grid;
    if (r > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  *tflops = best;
  return THB_OK;
}

int thb_p3p(const double* features, const double* world_points, int32_t count, double* R_out, double* t_out, int32_t* num_solutions,
            void* cuda_stream) {
  if (!features || !world_points || !R_out || !t_out || !num_solutions || count < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  return run_solver<6, 9, 36, 12>(features, world_points, count, R_out, t_out, num_solutions, cuda_stream, 0);
}
int thb_four_point_homography(const double* corr, int32_t count, double* H_out, int32_t* ok, void* cuda_stream) {
  if (!corr || !H_out || !ok || count < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  return run_solver<16, 0, 9, 0>(corr, nullptr, count, H_out, nullptr, ok, cuda_stream, 1);
}
int thb_seven_point_fundamental_matrix(const double* corr, int32_t count, double* F_out, int32_t* num_solutions, void* cuda_stream) {
  if (!corr || !F_out || !num_solutions || count < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  return run_solver<28, 0, 27, 0>(corr, nullptr, count, F_out, nullptr, num_solutions, cuda_stream, 2);
}

int thb_five_point_relative_pose(const double* x1, const double* x2, int32_t count, double* E_out, int32_t* num_solutions,
                                 void* cuda_stream) {
  if (!x1 || !x2 || !E_out || !num_solutions || count < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  if (count == 0) return THB_OK;
  int rc = check_device();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Bufs B;
  B.st = st;
  double* a = B.get<double>((size_t)count * 10); double* b = B.get<double>((size_t)count * 10);
  double* e = B.get<double>((size_t)count * 90); int* n = B.get<int>(count);
  if (!a || !b || !e || !n) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemcpyAsync(a, x1, sizeof(double) * 10 * count, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(b, x2, sizeof(double) * 10 * count, cudaMemcpyHostToDevice, st));
  k_five_point<<<(count + 31) / 32, 32, 0, st>>>(a, b, count, e, n);
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(E_out, e, sizeof(double) * 90 * count, cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(num_solutions, n, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

}  // extern "C"
