// ORACLE — test infrastructure only. Nothing under pytheiasfm_b200/ may include, link or
// call this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker and the CPU baseline.
//
// CPU restatement of the reference's RANSAC relative-pose path (paths under /root/reference/src/theia):
//   EstimateRelativePose / RelativePoseEstimator        sfm/estimators/estimate_relative_pose.cc:65-172
//   SampleConsensusEstimator::Estimate, ComputeMaxIterations   solvers/sample_consensus_estimator.h:251-415
//   RandomSampler / RandomNumberGenerator::RandInt      solvers/random_sampler.cc:53-72, util/random.cc:46-84
//   MLEQualityMeasurement / InlierSupport               solvers/mle_quality_measurement.h:58-72, inlier_support.h:54-64
//   FivePointRelativePose                               sfm/pose/five_point_relative_pose.cc:65-293
//   DecomposeEssentialMatrix / GetBestPoseFromEssentialMatrix  sfm/pose/essential_matrix_utils.cc:57-80,109-149
//   IsTriangulatedPointInFrontOfCameras                 sfm/triangulation/triangulation.cc:216-232
//   SquaredSampsonDistance                              sfm/pose/util.cc:56-69
// The RNG is libstdc++'s std::mt19937 + std::uniform_int_distribution<int>, used directly (the reference
// uses the same types, util/random.cc:46-84). Eigen decompositions: eigen_restated.h (PARITY UNPINNED).
// Compiled with -ffp-contract=off so that +,-,*,/,sqrt round exactly as on the device (-fmad=false).

#include <omp.h>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "../include/theia_b200.h"
#include "eigen_restated.h"
#include "jet.h"
#include "camera_models.h"

namespace oracle {
namespace {

// Monomial bookkeeping for polynomials in (x, y, z), GrevLex order as in five_point_relative_pose.cc:65-140.
struct Mono { int x, y, z; };
const Mono kM1[4] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
const Mono kM2[10] = {{2, 0, 0}, {1, 1, 0}, {0, 2, 0}, {1, 0, 1}, {0, 1, 1}, {0, 0, 2}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
const Mono kM3[20] = {{3, 0, 0}, {2, 1, 0}, {1, 2, 0}, {0, 3, 0}, {2, 0, 1}, {1, 1, 1}, {0, 2, 1}, {1, 0, 2}, {0, 1, 2}, {0, 0, 3},
                      {2, 0, 0}, {1, 1, 0}, {0, 2, 0}, {1, 0, 1}, {0, 1, 1}, {0, 0, 2}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
struct MonoTables {
  int t11[4][4];   // index in kM2 of kM1[a] * kM1[b]
  int t21[10][4];  // index in kM3 of kM2[a] * kM1[b]
  MonoTables() {
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        for (int k = 0; k < 10; ++k)
          if (kM2[k].x == kM1[a].x + kM1[b].x && kM2[k].y == kM1[a].y + kM1[b].y && kM2[k].z == kM1[a].z + kM1[b].z) t11[a][b] = k;
    for (int a = 0; a < 10; ++a)
      for (int b = 0; b < 4; ++b)
        for (int k = 0; k < 20; ++k)
          if (kM3[k].x == kM2[a].x + kM1[b].x && kM3[k].y == kM2[a].y + kM1[b].y && kM3[k].z == kM2[a].z + kM1[b].z) t21[a][b] = k;
  }
};
const MonoTables kT;

inline void Mul11(const double* a, const double* b, double* out) {  // MultiplyDegOnePoly
  for (int k = 0; k < 10; ++k) out[k] = 0.0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[kT.t11[i][j]] += a[i] * b[j];
}
inline void Mul21(const double* a, const double* b, double* out) {  // MultiplyDegTwoDegOnePoly
  for (int k = 0; k < 20; ++k) out[k] = 0.0;
  for (int i = 0; i < 10; ++i) for (int j = 0; j < 4; ++j) out[kT.t21[i][j]] += a[i] * b[j];
}

// five_point_relative_pose.cc:212-293, minimal (5-point) branch. E_out: up to 10 row-major 3x3.
int FivePoint(const double* x1, const double* x2, double* E_out) {
  double epi[5 * 9];
  for (int i = 0; i < 5; ++i) {
    const double ax = x1[2 * i], ay = x1[2 * i + 1], bx = x2[2 * i], by = x2[2 * i + 1];
    double* r = epi + 9 * i;
    r[0] = bx * ax; r[1] = by * ax; r[2] = ax; r[3] = bx * ay; r[4] = by * ay; r[5] = ay; r[6] = bx; r[7] = by; r[8] = 1.0;
  }
  FullPivLU<5, 9> lu;
  lu.compute(epi);
  if (lu.dimensionOfKernel() != 4) return 0;
  double ns[9 * 4];
  lu.kernel(ns);
  // null_space_matrix[i][j] = row (i + 3 j) of the null space: the polynomial (in x, y, z, 1) of E(i, j)
  const double* E[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) E[i][j] = ns + 4 * (i + 3 * j);
  double C[10 * 20];
  {  // GetTraceConstraint: E E^T E - 1/2 trace(E E^T) E
    double eet[3][3][10], tr[10], tmp[10], t20[20];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < 10; ++k) eet[i][j][k] = 0.0;
        for (int m = 0; m < 3; ++m) { Mul11(E[i][m], E[j][m], tmp); for (int k = 0; k < 10; ++k) eet[i][j][k] += tmp[k]; }
        for (int k = 0; k < 10; ++k) eet[i][j][k] *= 2.0;
      }
    for (int k = 0; k < 10; ++k) tr[k] = eet[0][0][k] + eet[1][1][k] + eet[2][2][k];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double* row = C + 20 * (3 * i + j);
        for (int k = 0; k < 20; ++k) row[k] = 0.0;
        for (int m = 0; m < 3; ++m) { Mul21(eet[i][m], E[m][j], t20); for (int k = 0; k < 20; ++k) row[k] += t20[k]; }
        Mul21(tr, E[i][j], t20);
        for (int k = 0; k < 20; ++k) row[k] -= 0.5 * t20[k];
      }
  }
  {  // GetDeterminantConstraint
    double a[10], b[10], d[10], t20[20];
    double* row = C + 20 * 9;
    for (int k = 0; k < 20; ++k) row[k] = 0.0;
    const int idx[3][2][2][2] = {{{{0, 1}, {1, 2}}, {{0, 2}, {1, 1}}}, {{{0, 2}, {1, 0}}, {{0, 0}, {1, 2}}}, {{{0, 0}, {1, 1}}, {{0, 1}, {1, 0}}}};
    for (int c = 0; c < 3; ++c) {
      Mul11(E[idx[c][0][0][0]][idx[c][0][0][1]], E[idx[c][0][1][0]][idx[c][0][1][1]], a);
      Mul11(E[idx[c][1][0][0]][idx[c][1][0][1]], E[idx[c][1][1][0]][idx[c][1][1][1]], b);
      for (int k = 0; k < 10; ++k) d[k] = a[k] - b[k];
      Mul21(d, E[2][c], t20);
      for (int k = 0; k < 20; ++k) row[k] += t20[k];
    }
  }
  // eliminate: solve C[:, :10] X = C[:, 10:]
  double L[100], Rm[100], X[100];
  for (int r = 0; r < 10; ++r) for (int c = 0; c < 10; ++c) { L[r * 10 + c] = C[r * 20 + c]; Rm[r * 10 + c] = C[r * 20 + 10 + c]; }
  FullPivLU<10, 10> clu;
  clu.compute(L);
  clu.solve<10>(Rm, X);
  double act[100];
  for (int i = 0; i < 100; ++i) act[i] = 0.0;
  const int src[6] = {0, 1, 2, 4, 5, 7};
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 10; ++c) act[r * 10 + c] = X[src[r] * 10 + c];
  act[6 * 10 + 0] = -1.0; act[7 * 10 + 1] = -1.0; act[8 * 10 + 3] = -1.0; act[9 * 10 + 6] = -1.0;
  EigenSolverReal<10> es;
  es.compute(act);
  int n = 0;
  for (int i = 0; i < 10; ++i) {
    if (es.eig_im[i] != 0.0) continue;
    double e9[9];
    for (int r = 0; r < 9; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += ns[r * 4 + k] * es.vec[(6 + k) * 10 + i];
      e9[r] = s;
    }
    // Map<Matrix<double,9,1>>(ematrix.data()): column-major fill -> E(r, c) = e9[c*3 + r]
    double* Eo = E_out + 9 * n;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Eo[r * 3 + c] = e9[c * 3 + r];
    ++n;
  }
  return n;
}

inline double Det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// triangulation.cc:216-232. Explicit fma in a fixed order (shared with the device code) so that the
// cheirality test and the Sampson distance are bit-reproducible across CPU and GPU.
inline bool InFront(const double* c, const double* R, const double* pos) {
  const double x1 = c[0], y1 = c[1], x2 = c[2], y2 = c[3];
  const double d2x = std::fma(R[0], x2, std::fma(R[3], y2, R[6]));  // R^T [x2 y2 1]
  const double d2y = std::fma(R[1], x2, std::fma(R[4], y2, R[7]));
  const double d2z = std::fma(R[2], x2, std::fma(R[5], y2, R[8]));
  const double dir1_sq = std::fma(x1, x1, std::fma(y1, y1, 1.0));
  const double dir2_sq = std::fma(d2x, d2x, std::fma(d2y, d2y, d2z * d2z));
  const double dir1_dir2 = std::fma(x1, d2x, std::fma(y1, d2y, d2z));
  const double dir1_pos = std::fma(x1, pos[0], std::fma(y1, pos[1], pos[2]));
  const double dir2_pos = std::fma(d2x, pos[0], std::fma(d2y, pos[1], d2z * pos[2]));
  return std::fma(dir2_sq, dir1_pos, -(dir1_dir2 * dir2_pos)) > 0.0 && std::fma(dir1_dir2, dir1_pos, -(dir1_sq * dir2_pos)) > 0.0;
}

// pose/util.cc:56-69
inline double Sampson(const double* F, const double* c) {
  const double x0 = c[0], x1 = c[1], y0 = c[2], y1 = c[3];
  const double ex0 = std::fma(F[0], x0, std::fma(F[1], x1, F[2]));
  const double ex1 = std::fma(F[3], x0, std::fma(F[4], x1, F[5]));
  const double ex2 = std::fma(F[6], x0, std::fma(F[7], x1, F[8]));
  const double num = std::fma(y0, ex0, std::fma(y1, ex1, ex2));
  const double dy0 = std::fma(y0, F[0], std::fma(y1, F[3], F[6]));
  const double dy1 = std::fma(y0, F[1], std::fma(y1, F[4], F[7]));
  const double den = std::fma(dy0, dy0, std::fma(dy1, dy1, std::fma(ex0, ex0, ex1 * ex1)));
  return num * num / den;
}

// essential_matrix_utils.cc:57-80 and 109-149 over the sample's correspondences.
int BestPose(const double* E, const double* corr, int n, double* Rbest, double* pbest) {
  double U[9], S[3], V[9];
  JacobiSVD3(E, U, S, V);
  if (Det3(U) < 0) for (int r = 0; r < 3; ++r) U[r * 3 + 2] *= -1.0;
  if (Det3(V) < 0) for (int r = 0; r < 3; ++r) V[r * 3 + 2] *= -1.0;
  // d = [0 1 0; -1 0 0; 0 0 1]; R1 = U d V^T, R2 = U d^T V^T
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
  const double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  for (int k = 0; k < 3; ++k) t[k] /= tn;
  int best = -1, best_count = -1;
  double Rc[4][9], pc[4][3];
  for (int i = 0; i < 4; ++i) {
    const double* Ri = R[i / 2];
    const double sgn = (i % 2 == 0) ? 1.0 : -1.0;
    for (int k = 0; k < 9; ++k) Rc[i][k] = Ri[k];
    for (int k = 0; k < 3; ++k) pc[i][k] = -(Ri[0 * 3 + k] * (sgn * t[0]) + Ri[1 * 3 + k] * (sgn * t[1]) + Ri[2 * 3 + k] * (sgn * t[2]));
    int count = 0;
    for (int m = 0; m < n; ++m) count += InFront(corr + 4 * m, Rc[i], pc[i]) ? 1 : 0;
    if (count > best_count) { best_count = count; best = i; }  // std::max_element: first maximum
  }
  for (int k = 0; k < 9; ++k) Rbest[k] = Rc[best][k];
  for (int k = 0; k < 3; ++k) pbest[k] = pc[best][k];
  return best_count;
}

// ---------------------------------------------------------------------------------------------------------
// math/find_polynomial_roots_companion_matrix.cc:89-233 (balance :89-137, companion :139-151) and
// math/polynomial.cc:176-231 (degree 1 and 2). coeffs: highest degree first. Returns the number of roots.
template <int D>
int CompanionRoots(const double* monic_tail /* p1..pD of the monic polynomial */, double* re, double* im) {
  double C[D * D], Off[D * D];
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
      for (int k = 0; k < D; ++k) { row_norm += std::fabs(Off[i * D + k]); col_norm += std::fabs(Off[k * D + i]); }
      int exponent = 0;
      std::frexp(row_norm / col_norm, &exponent);
      exponent /= 2;
      if (exponent != 0) {
        const double scaled_col = std::ldexp(col_norm, exponent), scaled_row = std::ldexp(row_norm, -exponent);
        if (scaled_col + scaled_row < gamma * (col_norm + row_norm)) {
          changed = true;
          const double fr = std::ldexp(1.0, -exponent), fc = std::ldexp(1.0, exponent);
          for (int k = 0; k < D; ++k) Off[i * D + k] *= fr;
          for (int k = 0; k < D; ++k) Off[k * D + i] *= fc;
        }
      }
    }
  } while (changed);
  for (int i = 0; i < D; ++i) Off[i * D + i] = C[i * D + i];
  EigenSolverReal<D> es;
  es.compute(Off, false);
  if (!es.ok) return 0;
  for (int i = 0; i < D; ++i) { re[i] = es.eig_re[i]; im[i] = es.eig_im[i]; }
  return D;
}
// FindPolynomialRoots for a polynomial of size n+1 (degree <= 4 here). Real parts in re, imaginary in im.
int PolyRoots(const double* poly_in, int size, double* re, double* im) {
  int lead = 0;
  while (lead < size - 1 && poly_in[lead] == 0.0) ++lead;  // RemoveLeadingZeros
  const double* p = poly_in + lead;
  const int degree = size - lead - 1;
  if (degree == 0) return 0;
  if (degree == 1) { re[0] = -p[1] / p[0]; im[0] = 0.0; return 1; }
  if (degree == 2) {
    const double a = p[0], b = p[1], c = p[2];
    const double D = b * b - 4 * a * c, sqrt_D = std::sqrt(std::fabs(D));
    im[0] = im[1] = 0.0;
    if (D >= 0) {
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
  if (degree == 3) return CompanionRoots<3>(tail, re, im);
  return CompanionRoots<4>(tail, re, im);
}

inline void Cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double Dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void Normalize3(double* a) { const double n = std::sqrt(Dot3(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; }

// sfm/pose/perspective_three_point.cc:182-291 (Kneip), with SolvePlaneRotation :58-130 and Backsubstitute :135-178.
// feat: 3 x (x, y); world: 3 x (X, Y, Z). Outputs up to 4 (R row-major, t). Returns the number of solutions.
int P3P(const double* feat, const double* world, double* Rs, double* ts) {
  double f[3][3], w[3][3];
  for (int i = 0; i < 3; ++i) {
    f[i][0] = feat[2 * i]; f[i][1] = feat[2 * i + 1]; f[i][2] = 1.0;
    Normalize3(f[i]);
    for (int k = 0; k < 3; ++k) w[i][k] = world[3 * i + k];
  }
  double w10[3], w20[3], cr[3];
  for (int k = 0; k < 3; ++k) { w10[k] = w[1][k] - w[0][k]; w20[k] = w[2][k] - w[0][k]; }
  Cross(w10, w20, cr);
  if (Dot3(cr, cr) < 1e-6) return 0;
  double Tc[3][3];  // intermediate camera frame (rows)
  auto camera_frame = [&]() {
    for (int k = 0; k < 3; ++k) Tc[0][k] = f[0][k];
    Cross(f[0], f[1], Tc[2]); Normalize3(Tc[2]);
    Cross(Tc[2], Tc[0], Tc[1]);
  };
  camera_frame();
  double ip[3];
  for (int r = 0; r < 3; ++r) ip[r] = Dot3(Tc[r], f[2]);
  if (ip[2] > 0) {
    for (int k = 0; k < 3; ++k) { std::swap(f[0][k], f[1][k]); }
    camera_frame();
    for (int r = 0; r < 3; ++r) ip[r] = Dot3(Tc[r], f[2]);
    for (int k = 0; k < 3; ++k) std::swap(w[0][k], w[1][k]);
    for (int k = 0; k < 3; ++k) { w10[k] = w[1][k] - w[0][k]; w20[k] = w[2][k] - w[0][k]; }
  }
  double Nw[3][3];  // intermediate world frame (rows)
  for (int k = 0; k < 3; ++k) Nw[0][k] = w10[k];
  Normalize3(Nw[0]);
  Cross(Nw[0], w20, Nw[2]); Normalize3(Nw[2]);
  Cross(Nw[2], Nw[0], Nw[1]);
  double iw[3];
  for (int r = 0; r < 3; ++r) iw[r] = Dot3(Nw[r], w20);
  const double d_12 = std::sqrt(Dot3(w10, w10));
  // SolvePlaneRotation
  const double f_1 = ip[0] / ip[2], f_2 = ip[1] / ip[2], p_1 = iw[0], p_2 = iw[1];
  const double cos_beta = Dot3(f[0], f[1]);
  double b = 1.0 / (1.0 - cos_beta * cos_beta) - 1.0;
  b = cos_beta < 0 ? -std::sqrt(b) : std::sqrt(b);
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
  const int nroots = PolyRoots(co, 5, re, im);
  for (int s = 0; s < nroots; ++s) {
    const double cos_theta = re[s];
    const double cot_alpha = (-f_1 * p_1 / f_2 - cos_theta * p_2 + d_12 * b) / (-f_1 * cos_theta * p_2 / f_2 + p_1 - d_12);
    // Backsubstitute
    const double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    const double sin_alpha = std::sqrt(1.0 / (cot_alpha * cot_alpha + 1.0));
    double cos_alpha = std::sqrt(1.0 - sin_alpha * sin_alpha);
    if (cot_alpha < 0) cos_alpha = -cos_alpha;
    const double k = sin_alpha * b + cos_alpha;
    const double c_nu[3] = {d_12 * cos_alpha * k, cos_theta * d_12 * sin_alpha * k, sin_theta * d_12 * sin_alpha * k};
    double trans[3];
    for (int c = 0; c < 3; ++c) trans[c] = w[0][c] + (Nw[0][c] * c_nu[0] + Nw[1][c] * c_nu[1] + Nw[2][c] * c_nu[2]);  // N^T c_nu
    const double Q[3][3] = {{-cos_alpha, -sin_alpha * cos_theta, -sin_alpha * sin_theta},
                            {sin_alpha, -cos_alpha * cos_theta, -cos_alpha * sin_theta},
                            {0, -sin_theta, cos_theta}};
    // rotation = (N^T Q^T Tc)^T = Tc^T Q N
    double QN[3][3], R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) QN[r][c] = Q[r][0] * Nw[0][c] + Q[r][1] * Nw[1][c] + Q[r][2] * Nw[2][c];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = Tc[0][r] * QN[0][c] + Tc[1][r] * QN[1][c] + Tc[2][r] * QN[2][c];
    for (int k2 = 0; k2 < 9; ++k2) Rs[9 * s + k2] = R[k2];
    for (int r = 0; r < 3; ++r) ts[3 * s + r] = -(R[r * 3] * trans[0] + R[r * 3 + 1] * trans[1] + R[r * 3 + 2] * trans[2]);
  }
  return nroots;
}

// pose/util.cc:81-111
void NormalizeImagePoints(const double* pts, int n, int stride, double* out, double* T) {
  double cx = 0.0, cy = 0.0;
  for (int i = 0; i < n; ++i) { cx += pts[i * stride]; cy += pts[i * stride + 1]; }
  cx /= n; cy /= n;
  double sq = 0.0;
  for (int i = 0; i < n; ++i) { const double dx = pts[i * stride] - cx, dy = pts[i * stride + 1] - cy; sq += dx * dx; sq += dy * dy; }
  const double rms = std::sqrt(sq / n);
  const double nf = std::sqrt(2.0) / rms;
  T[0] = nf; T[1] = 0; T[2] = -1.0 * nf * cx; T[3] = 0; T[4] = nf; T[5] = -1.0 * nf * cy; T[6] = 0; T[7] = 0; T[8] = 1;
  for (int i = 0; i < n; ++i) {
    const double x = pts[i * stride], y = pts[i * stride + 1];
    const double hx = T[0] * x + T[1] * y + T[2], hy = T[3] * x + T[4] * y + T[5], hw = T[6] * x + T[7] * y + T[8];
    out[2 * i] = hx / hw; out[2 * i + 1] = hy / hw;
  }
}
inline void Inverse3(const double* M, double* inv) {
  const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  const double det = M[0] * c00 + M[1] * c01 + M[2] * c02, id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = (M[2] * M[7] - M[1] * M[8]) * id; inv[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  inv[3] = c01 * id; inv[4] = (M[0] * M[8] - M[2] * M[6]) * id; inv[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  inv[6] = c02 * id; inv[7] = (M[1] * M[6] - M[0] * M[7]) * id; inv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}
inline void Mul33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// sfm/pose/four_point_homography.cc:72-102, minimal case (4 correspondences x1,y1,x2,y2). H row-major.
bool FourPointH(const double* corr, double* H) {
  double n1[8], n2[8], T1[9], T2[9];
  NormalizeImagePoints(corr, 4, 4, n1, T1);
  NormalizeImagePoints(corr + 2, 4, 4, n2, T2);
  double A[8 * 9];
  for (int i = 0; i < 4; ++i) {
    const double x = n1[2 * i], y = n1[2 * i + 1], u = n2[2 * i], v = n2[2 * i + 1];
    double* r0 = A + 18 * i; double* r1 = r0 + 9;
    r0[0] = 0; r0[1] = 0; r0[2] = 0; r0[3] = -x; r0[4] = -y; r0[5] = -1.0; r0[6] = x * v; r0[7] = y * v; r0[8] = v;
    r1[0] = x; r1[1] = y; r1[2] = 1.0; r1[3] = 0; r1[4] = 0; r1[5] = 0; r1[6] = -x * u; r1[7] = -y * u; r1[8] = -u;
  }
  double AtA[81];
  for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) { double s = 0.0; for (int k = 0; k < 8; ++k) s += A[k * 9 + r] * A[k * 9 + c]; AtA[r * 9 + c] = s; }
  double U[81], S[9], V[81];
  JacobiSVDN<9>(AtA, U, S, V);
  double Hn[9];
  for (int k = 0; k < 9; ++k) Hn[k] = V[k * 9 + 8];  // Map<Matrix3d>(null).transpose(): H(r,c) = null[3r + c]
  double T2i[9], tmp[9];
  Inverse3(T2, T2i);
  Mul33(T2i, Hn, tmp);
  Mul33(tmp, T1, H);
  return true;
}

// sfm/pose/seven_point_fundamental_matrix.cc:72-152 — including its coefficient-order quirk: the
// det(F2)-only term is stored at index 0 although polynomial(0) is the highest-degree coefficient, and the real
// parts of ALL roots are used. img: 7 x (x1,y1,x2,y2). F_out: up to 3 row-major matrices.
int SevenPointF(const double* corr, double* F_out) {
  double n1[14], n2[14], T1[9], T2[9];
  NormalizeImagePoints(corr, 7, 4, n1, T1);
  NormalizeImagePoints(corr + 2, 7, 4, n2, T2);
  double epi[7 * 9];
  for (int i = 0; i < 7; ++i) {
    const double ax = n1[2 * i], ay = n1[2 * i + 1], bx = n2[2 * i], by = n2[2 * i + 1];
    double* r = epi + 9 * i;
    r[0] = bx * ax; r[1] = by * ax; r[2] = ax; r[3] = bx * ay; r[4] = by * ay; r[5] = ay; r[6] = bx; r[7] = by; r[8] = 1.0;
  }
  FullPivLU<7, 9> lu;
  lu.compute(epi);
  if (lu.dimensionOfKernel() != 2) return 0;
  double ns[18];
  lu.kernel(ns);
  double v1[9], v2[9];
  for (int k = 0; k < 9; ++k) { v1[k] = ns[k * 2] - ns[k * 2 + 1]; v2[k] = ns[k * 2 + 1]; }
  // Map<const Matrix3d>(vec): column-major, M(r, c) = vec[c*3 + r]
  auto F1 = [&](int r, int c) { return v1[c * 3 + r]; };
  auto F2 = [&](int r, int c) { return v2[c * 3 + r]; };
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
  const int nroots = PolyRoots(dc, 4, re, im);
  double T2t[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T2t[r * 3 + c] = T2[c * 3 + r];
  for (int s = 0; s < nroots; ++s) {
    double M[9], tmp[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r * 3 + c] = re[s] * F1(r, c) + F2(r, c);
    Mul33(T2t, M, tmp);
    Mul33(tmp, T1, F_out + 9 * s);
  }
  return nroots;
}

// ---------------------------------------------------------------------------------------------------------
// Estimator policies (solvers/estimator.h): sample size, model estimation, per-datum error.
struct Model { double E[9], R[9], p[3]; };  // 21 doubles: the ThbRelPoseResult payload

// ---- LO-RANSAC refinement of a relative pose: RelativePoseEstimator::RefineModel (estimate_relative_pose.cc:111-138) ->
// BundleAdjustTwoViewsAngular (sfm/bundle_adjustment/bundle_adjust_two_views.cc:195-246) over AngularEpipolarError
// (angular_epipolar_error.h:50-108): rotation (angle-axis, 3) and position (unit 3-vector, SphereManifold<3>), TRUNCATED loss
// of width error_thresh, 15 iterations, Ceres' default tolerances and trust-region radii (the local SetSolverOptions of
// bundle_adjust_two_views.cc:60-72 overrides the linear solver with DENSE_SCHUR and nothing else). Eigen conversions
// (AngleAxisd <-> Matrix3d via Quaterniond) and the Ceres minimiser are restated: PARITY UNPINNED like the rest.
inline void EigenRotationMatrixToAngleAxis(const double* R /*row-major*/, double aa[3]) {
  // Eigen::Quaternion = Matrix3 (QuaternionBase::operator=, quat_product / rotation matrix branch), then AngleAxis = Quaternion
  double q[4];  // x, y, z, w
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  double angle, axis[3];
  if (n != 0.0) {
    angle = 2.0 * std::atan2(n, std::fabs(q[3]));
    if (q[3] < 0.0) n = -n;
    for (int k = 0; k < 3; ++k) axis[k] = q[k] / n;
  } else { angle = 0.0; axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; }
  for (int k = 0; k < 3; ++k) aa[k] = angle * axis[k];
}
inline void EigenAngleAxisToRotationMatrix(const double aa[3], double* R /*row-major*/) {
  // rot_vec_out.angle() = |aa|; axis = aa / angle; toRotationMatrix() (Eigen/src/Geometry/AngleAxis.h)
  const double angle = std::sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  const double ax[3] = {aa[0] / angle, aa[1] / angle, aa[2] / angle};
  const double s = std::sin(angle), c = std::cos(angle);
  const double sa[3] = {s * ax[0], s * ax[1], s * ax[2]}, c1[3] = {(1.0 - c) * ax[0], (1.0 - c) * ax[1], (1.0 - c) * ax[2]};
  double tmp;
  tmp = c1[0] * ax[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
  tmp = c1[0] * ax[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
  tmp = c1[1] * ax[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
  R[0] = c1[0] * ax[0] + c; R[4] = c1[1] * ax[1] + c; R[8] = c1[2] * ax[2] + c;
}

template <typename T>
inline void CeresAngleAxisToRotationMatrix(const T* aa, T R[3][3]) {  // ceres/rotation.h (external)
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > DBL_EPSILON) {
    const T theta = sqrt_(theta2);
    const T wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const T ct = cos_(theta), st = sin_(theta);
    R[0][0] = ct + wx * wx * (1.0 - ct);      R[1][0] = wz * st + wx * wy * (1.0 - ct);  R[2][0] = -wy * st + wx * wz * (1.0 - ct);
    R[0][1] = wx * wy * (1.0 - ct) - wz * st; R[1][1] = ct + wy * wy * (1.0 - ct);       R[2][1] = wx * st + wy * wz * (1.0 - ct);
    R[0][2] = wy * st + wx * wz * (1.0 - ct); R[1][2] = -wx * st + wy * wz * (1.0 - ct); R[2][2] = ct + wz * wz * (1.0 - ct);
  } else {
    R[0][0] = T(1.0); R[1][0] = aa[2]; R[2][0] = -aa[1];
    R[0][1] = -aa[2]; R[1][1] = T(1.0); R[2][1] = aa[0];
    R[0][2] = aa[1]; R[1][2] = -aa[0]; R[2][2] = T(1.0);
  }
}

// AngularEpipolarError::operator() (angular_epipolar_error.h:61-96), as written (including R f2 . T R^T f2)
template <typename T>
inline T AngularEpipolarResidual(const T* rot, const T* tr, const double* c /*x1 y1 x2 y2*/) {
  const T f1[3] = {T(c[0]), T(c[1]), T(1.0)}, f2[3] = {T(c[2]), T(c[3]), T(1.0)};
  T R[3][3];
  CeresAngleAxisToRotationMatrix(rot, R);
  T Rf2[3], Rtf2[3];
  for (int i = 0; i < 3; ++i) { Rf2[i] = R[i][0] * f2[0] + R[i][1] * f2[1] + R[i][2] * f2[2]; Rtf2[i] = R[0][i] * f2[0] + R[1][i] * f2[1] + R[2][i] * f2[2]; }
  auto apply_term = [&](const T* v, T* out) {  // (I - t t^T) v
    const T tv = tr[0] * v[0] + tr[1] * v[1] + tr[2] * v[2];
    for (int i = 0; i < 3; ++i) out[i] = v[i] - tr[i] * tv;
  };
  T Tf1[3], TRtf2[3];
  apply_term(f1, Tf1); apply_term(Rtf2, TRtf2);
  const T a = (f1[0] * Tf1[0] + f1[1] * Tf1[1] + f1[2] * Tf1[2]) + (Rf2[0] * TRtf2[0] + Rf2[1] * TRtf2[1] + Rf2[2] * TRtf2[2]);
  const T cr[3] = {f1[1] * Rtf2[2] - f1[2] * Rtf2[1], f1[2] * Rtf2[0] - f1[0] * Rtf2[2], f1[0] * Rtf2[1] - f1[1] * Rtf2[0]};
  const T b_sqrt = tr[0] * cr[0] + tr[1] * cr[1] + tr[2] * cr[2];
  const T sqrt_term = (a * a) / 4.0 - b_sqrt * b_sqrt;
  if (sqrt_term < 0.0) return T(1000.0);
  return a / 2.0 - sqrt_(sqrt_term);
}

// ceres SphereManifold<3> on the position (Householder as for the 4-vector points of ba_oracle.cc)
inline void Householder3(const double x[3], double v[3], double* beta) {
  const double sigma = x[0] * x[0] + x[1] * x[1];
  v[0] = x[0]; v[1] = x[1]; v[2] = 1.0;
  *beta = 0.0;
  if (sigma <= DBL_EPSILON) { if (x[2] < 0.0) *beta = 2.0; return; }
  const double mu = std::sqrt(x[2] * x[2] + sigma);
  const double vp = x[2] <= 0.0 ? x[2] - mu : -sigma / (x[2] + mu);
  *beta = 2.0 * vp * vp / (sigma + vp * vp);
  v[0] /= vp; v[1] /= vp;
}
inline void Sphere3Plus(const double x[3], const double d[2], double out[3]) {
  const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1]);
  if (nd == 0.0) { for (int i = 0; i < 3; ++i) out[i] = x[i]; return; }
  double v[3], beta;
  Householder3(x, v, &beta);
  const double sbd = std::sin(nd) / nd;
  const double y[3] = {sbd * d[0], sbd * d[1], std::cos(nd)};
  const double vty = v[0] * y[0] + v[1] * y[1] + v[2] * y[2];
  const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int i = 0; i < 3; ++i) out[i] = nx * (y[i] - v[i] * (beta * vty));
}
inline void Sphere3PlusJacobian(const double x[3], double J[6] /*3x2 row-major*/) {
  double v[3], beta;
  Householder3(x, v, &beta);
  const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int i = 0; i < 2; ++i) {
    for (int r = 0; r < 3; ++r) J[r * 2 + i] = -beta * v[i] * v[r];
    J[i * 2 + i] += 1.0;
  }
  for (int k = 0; k < 6; ++k) J[k] *= nx;
}

inline bool SmallSpdSolve(int n, const double* A, const double* b, double* x) {  // Cholesky, n <= 8
  double L[64], y[8];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = A[i * n + j];
      for (int k = 0; k < j; ++k) v -= L[i * n + k] * L[j * n + k];
      if (i == j) { if (!(v > 0.0)) return false; L[i * n + i] = std::sqrt(v); }
      else L[i * n + j] = v / L[j * n + j];
    }
  for (int i = 0; i < n; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[i * n + k] * y[k]; y[i] = v / L[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < n; ++k) v -= L[k * n + i] * x[k]; x[i] = v / L[i * n + i]; }
  return true;
}

// TrustRegionMinimizer on (rotation, position). Returns IsSolutionUsable-style success; costs as ceres::Solver::Summary.
inline bool AngularBundleAdjust(const double* corr, int n, double width, int max_iterations, double rot[3], double pos[3],
                                double* initial_cost, double* final_cost) {
  typedef Jet<6> J6;
  const double b = width * width;  // TruncatedLoss(a): rho(s) = min(s, a^2) (loss_functions.cc:40-44)
  struct Eval { double cost; double H[15]; double g[5]; };
  auto evaluate = [&](const double* r3, const double* p3, bool want_j, const double* scale, Eval* e) {
    e->cost = 0.0;
    if (want_j) { for (double& v : e->H) v = 0.0; for (double& v : e->g) v = 0.0; }
    double PJ[6];
    if (want_j) Sphere3PlusJacobian(p3, PJ);
    for (int i = 0; i < n; ++i) {
      if (!want_j) {
        const double r = AngularEpipolarResidual<double>(r3, p3, corr + 4 * (size_t)i);
        e->cost += 0.5 * std::min(r * r, b);
        continue;
      }
      J6 jr[3], jt[3];
      for (int k = 0; k < 3; ++k) { jr[k] = J6(r3[k], k); jt[k] = J6(p3[k], 3 + k); }
      const J6 res = AngularEpipolarResidual<J6>(jr, jt, corr + 4 * (size_t)i);
      const double s = res.a * res.a;
      e->cost += 0.5 * std::min(s, b);
      const double w = s < b ? 1.0 : 0.0;  // Corrector with rho'' = 0: residual and Jacobian times sqrt(rho')
      if (w == 0.0) continue;
      double t[5];
      for (int k = 0; k < 3; ++k) t[k] = res.v[k];
      for (int k = 0; k < 2; ++k) t[3 + k] = res.v[3] * PJ[0 * 2 + k] + res.v[4] * PJ[1 * 2 + k] + res.v[5] * PJ[2 * 2 + k];
      for (int k = 0; k < 5; ++k) t[k] *= scale[k];
      int q = 0;
      for (int a = 0; a < 5; ++a) { e->g[a] += t[a] * res.a; for (int c = 0; c <= a; ++c) e->H[q++] += t[a] * t[c]; }
    }
  };
  double scale[5] = {1, 1, 1, 1, 1};
  Eval E;
  evaluate(rot, pos, true, scale, &E);
  { int q = 0; for (int a = 0; a < 5; ++a) { q += a; scale[a] = 1.0 / (1.0 + std::sqrt(E.H[q])); ++q; } }
  evaluate(rot, pos, true, scale, &E);
  double x_cost = E.cost, min_cost = E.cost;
  *initial_cost = E.cost;
  double x_norm = std::sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]);
  double radius = 1e4, decrease_factor = 2.0, diag[5];
  bool step_ok = true, reuse_diag = false, failure = false;
  int iteration = 0, invalid = 0;
  for (;;) {
    if (iteration >= max_iterations) break;
    if (step_ok) { double gmax = 0.0; for (int k = 0; k < 5; ++k) gmax = std::max(gmax, std::fabs(E.g[k] / scale[k])); if (gmax <= 1e-10) break; }
    if (radius <= 1e-32) break;
    ++iteration;
    step_ok = false;
    if (!reuse_diag) { int q = 0; for (int a = 0; a < 5; ++a) { q += a; diag[a] = std::min(std::max(E.H[q], 1e-6), 1e32); ++q; } }
    reuse_diag = true;
    double M[25], y[5];
    { int q = 0; for (int a = 0; a < 5; ++a) for (int c = 0; c <= a; ++c) { M[a * 5 + c] = E.H[q]; M[c * 5 + a] = E.H[q]; ++q; } }
    for (int a = 0; a < 5; ++a) M[a * 5 + a] += diag[a] / radius;
    bool valid = SmallSpdSolve(5, M, E.g, y);
    double mcc = 0.0;
    if (valid) {
      double yg = 0.0, yHy = 0.0;
      int q = 0;
      for (int a = 0; a < 5; ++a) { yg += y[a] * E.g[a]; for (int c = 0; c <= a; ++c) { yHy += (a == c ? 1.0 : 2.0) * y[a] * E.H[q] * y[c]; ++q; } }
      mcc = yg - 0.5 * yHy;
      valid = std::isfinite(mcc) && mcc > 0.0;
    }
    if (!valid) {
      if (++invalid >= 5) { failure = true; break; }
      radius /= decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    double cr[3], cp[3], d2[2] = {-y[3] * scale[3], -y[4] * scale[4]};
    for (int k = 0; k < 3; ++k) cr[k] = rot[k] + (-y[k] * scale[k]);
    Sphere3Plus(pos, d2, cp);
    Eval C;
    evaluate(cr, cp, false, scale, &C);
    double sn = 0.0, cn = 0.0;
    for (int k = 0; k < 3; ++k) { sn += (cr[k] - rot[k]) * (cr[k] - rot[k]) + (cp[k] - pos[k]) * (cp[k] - pos[k]); cn += cr[k] * cr[k] + cp[k] * cp[k]; }
    if (std::sqrt(sn) <= 1e-8 * (x_norm + 1e-8)) break;
    const double cost_change = x_cost - C.cost;
    if (std::fabs(cost_change) <= 1e-6 * x_cost) break;
    const double rel = cost_change / mcc;
    if (rel > 1e-3) {
      for (int k = 0; k < 3; ++k) { rot[k] = cr[k]; pos[k] = cp[k]; }
      x_norm = std::sqrt(cn);
      evaluate(rot, pos, true, scale, &E);
      x_cost = E.cost;
      min_cost = std::min(min_cost, x_cost);
      step_ok = true;
      radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3)));
      decrease_factor = 2.0; reuse_diag = false;
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
    }
  }
  *final_cost = min_cost;
  return !failure;
}

// RelativePoseEstimator::RefineModel: E is NOT recomputed from the refined pose (estimate_relative_pose.cc:111-138)
inline bool RefineRelativePose(const double* corr, int n, double error_thresh, Model* m) {
  double rot[3], pos[3] = {m->p[0], m->p[1], m->p[2]};
  EigenRotationMatrixToAngleAxis(m->R, rot);
  double c0 = 0.0, c1 = 0.0;
  AngularBundleAdjust(corr, n, error_thresh, 15, rot, pos, &c0, &c1);
  for (int k = 0; k < 3; ++k) m->p[k] = pos[k];
  EigenAngleAxisToRotationMatrix(rot, m->R);
  return c1 < c0;
}

struct RelPoseEst {  // RelativePoseEstimator, estimate_relative_pose.cc:65-155
  static constexpr int S = 5, D = 4, MAXM = 10;
  static constexpr bool HAS_LO = true;
  static bool Refine(const double* inl, int n, double thresh, Model* m) { return RefineRelativePose(inl, n, thresh, m); }
  static int Solve(const double* sample, Model* out) {
    double x1[10], x2[10], Es[90];
    for (int i = 0; i < 5; ++i) { x1[2 * i] = sample[4 * i]; x1[2 * i + 1] = sample[4 * i + 1]; x2[2 * i] = sample[4 * i + 2]; x2[2 * i + 1] = sample[4 * i + 3]; }
    const int ne = FivePoint(x1, x2, Es);
    int n = 0;
    for (int e = 0; e < ne; ++e) {
      Model m;
      std::memcpy(m.E, Es + 9 * e, sizeof(m.E));
      if (BestPose(m.E, sample, 5, m.R, m.p) < 4) continue;
      out[n++] = m;
    }
    return n;
  }
  static double Error(const Model& m, const double* c) { return InFront(c, m.R, m.p) ? Sampson(m.E, c) : DBL_MAX; }
};

struct AbsPoseEst {  // CalibratedAbsolutePoseEstimator with PnPType::KNEIP, estimate_calibrated_absolute_pose.cc:63-172
  static constexpr int S = 3, D = 5, MAXM = 4;
  static constexpr bool HAS_LO = false;  // LO = BundleAdjustView on a one-view reconstruction (:120-153): not restated
  static bool Refine(const double*, int, double, Model*) { return false; }   // datum: feature (x, y), world point (X, Y, Z)
  static int Solve(const double* sample, Model* out) {
    double feat[6], world[9], Rs[36], ts[12];
    for (int i = 0; i < 3; ++i) { feat[2 * i] = sample[5 * i]; feat[2 * i + 1] = sample[5 * i + 1]; for (int k = 0; k < 3; ++k) world[3 * i + k] = sample[5 * i + 2 + k]; }
    const int n = P3P(feat, world, Rs, ts);
    for (int s = 0; s < n; ++s) {
      Model& m = out[s];
      std::memset(&m, 0, sizeof(m));
      for (int k = 0; k < 9; ++k) m.R[k] = Rs[9 * s + k];
      for (int c = 0; c < 3; ++c) m.p[c] = -(m.R[0 * 3 + c] * ts[3 * s] + m.R[1 * 3 + c] * ts[3 * s + 1] + m.R[2 * 3 + c] * ts[3 * s + 2]);  // -R^T t
    }
    return n;
  }
  static double Error(const Model& m, const double* d) {  // squared reprojection error, :158-167
    const double v[3] = {d[2] - m.p[0], d[3] - m.p[1], d[4] - m.p[2]};
    const double px = std::fma(m.R[0], v[0], std::fma(m.R[1], v[1], m.R[2] * v[2]));
    const double py = std::fma(m.R[3], v[0], std::fma(m.R[4], v[1], m.R[5] * v[2]));
    const double pz = std::fma(m.R[6], v[0], std::fma(m.R[7], v[1], m.R[8] * v[2]));
    const double ex = px / pz - d[0], ey = py / pz - d[1];
    return std::fma(ex, ex, ey * ey);
  }
};

struct HomographyEst {  // HomographyEstimator, estimate_homography.cc:62-116; the model lives in Model::E
  static constexpr int S = 4, D = 4, MAXM = 1;
  static constexpr bool HAS_LO = false;  // Estimator::RefineModel default returns false (solvers/estimator.h)
  static bool Refine(const double*, int, double, Model*) { return false; }
  static int Solve(const double* sample, Model* out) {
    std::memset(out, 0, sizeof(Model));
    return FourPointH(sample, out->E) ? 1 : 0;
  }
  static double Error(const Model& m, const double* c) {  // one-way transfer error
    const double* H = m.E;
    const double px = std::fma(H[0], c[0], std::fma(H[1], c[1], H[2]));
    const double py = std::fma(H[3], c[0], std::fma(H[4], c[1], H[5]));
    const double pz = std::fma(H[6], c[0], std::fma(H[7], c[1], H[8]));
    const double ex = c[2] - px / pz, ey = c[3] - py / pz;
    return std::fma(ex, ex, ey * ey);
  }
};

// sample_consensus_estimator.h:251-297
int ComputeMaxIterations(const ThbRansacParams& P, double min_sample_size, double inlier_ratio, double log_failure_prob, int total) {
  if (inlier_ratio == 1.0) return P.min_iterations;
  const int ninl = static_cast<int>(inlier_ratio * total);
  const double num_samples = P.use_tdd_test ? min_sample_size + 1 : min_sample_size;  // :272-279
  double a = 1.0, b = 1.0;
  for (int i = 0; i < num_samples; ++i) { a *= ninl - i; b *= total - i; }
  const double prob_all_inliers = a / b;
  if (prob_all_inliers < DBL_EPSILON) return P.max_iterations;
  if (prob_all_inliers >= 1.0 - DBL_EPSILON) return P.min_iterations;
  const double num_iterations = log_failure_prob / std::log(1.0 - prob_all_inliers);
  return std::max(static_cast<double>(P.min_iterations), std::min(num_iterations, static_cast<double>(P.max_iterations)));
}

// LmedQualityMeasurement::ComputeCost (solvers/lmed_quality_measurement.h:58-118): the cost is the median of the SQUARED
// residuals (the estimators' residuals are squared distances already; the reference squares them again), taken with
// std::nth_element exactly as there - the upper median for an even count, the mean of the two middle values for an odd one -
// and the inliers are the data below OpenCV's robust threshold 2.5 * 1.4826 * (1 + 5 / (n - m)) * sqrt(median).
inline double LmedCost(const std::vector<double>& residuals, int min_sample_size, std::vector<int>* inliers) {
  std::vector<double> sq(residuals.size());
  for (size_t i = 0; i < residuals.size(); ++i) sq[i] = residuals[i] * residuals[i];
  std::nth_element(sq.begin(), sq.begin() + sq.size() / 2, sq.end());
  double median = sq[sq.size() / 2];
  if ((sq.size() % 2) != 0) {
    std::nth_element(sq.begin(), sq.begin() + (sq.size() / 2) - 1, sq.end());
    median = 0.5 * (sq[(sq.size() / 2) - 1] + median);
  }
  const double inlier_threshold = 2.5 * 1.4826 * (1 + 5.0 / (residuals.size() - min_sample_size)) * std::sqrt(median);
  const double squared_inlier_threshold = inlier_threshold * inlier_threshold;
  for (size_t i = 0; i < residuals.size(); ++i)
    if ((residuals[i] * residuals[i]) < squared_inlier_threshold) inliers->push_back(static_cast<int>(i));
  return median;
}

template <class Est>
double Score(const ThbRansacParams& P, const double* data, int n, const Model& m, std::vector<int>* inliers) {
  inliers->clear();
  if (P.ransac_type == 2) {  // RansacType::LMED (solvers/lmed.h:65-72)
    std::vector<double> residuals(n);
    for (int i = 0; i < n; ++i) residuals[i] = Est::Error(m, data + Est::D * (size_t)i);
    return LmedCost(residuals, Est::S, inliers);
  }
  double cost = 0.0;
  for (int i = 0; i < n; ++i) {
    const double r = Est::Error(m, data + Est::D * (size_t)i);
    if (P.use_mle) {
      if (r < P.error_thresh) { cost += r; inliers->push_back(i); } else cost += P.error_thresh;
    } else if (r < P.error_thresh) {
      inliers->push_back(i);
    }
  }
  return P.use_mle ? cost : static_cast<double>(n - static_cast<int>(inliers->size()));
}

// SampleConsensusEstimator::Estimate (sample_consensus_estimator.h:299-415) with RandomSampler (random_sampler.cc:53-72)
// ProsacSampler (solvers/prosac_sampler.cc:53-131): the k-th sample draws from the top n data points (data sorted by quality),
// n grown by the schedule of Chum & Matas (Eq. 3-5) with T_N = 20000. The reference recomputes the schedule from t = 1 on every
// call; carrying (t_n, t_n_prime, n) from call to call performs the same operations in the same order. Its last index `n` can
// equal the number of data points once the schedule has reached the whole set (an out-of-bounds read there): clamped here.
struct ProsacSampler {
  double t_n = 0.0, t_n_prime = 1.0;
  int n = 0, k = 1, N = 0, m = 0;
  void Initialize(int num_datapoints, int min_num_samples) {
    N = num_datapoints; m = min_num_samples; k = 1; n = m; t_n_prime = 1.0;
    t_n = 20000.0;
    for (int i = 0; i < m; ++i) t_n *= static_cast<double>(n - i) / (N - i);
  }
  void Sample(std::mt19937& gen, int* subset) {
    const int t = k;  // the schedule step of this call (the reference's loop body for t = k)
    if (t > t_n_prime && n < N) {
      const double t_n_plus1 = (t_n * (n + 1.0)) / (n + 1.0 - m);
      t_n_prime += std::ceil(t_n_plus1 - t_n);
      t_n = t_n_plus1;
      ++n;
    }
    auto draw_unique = [&](int count, int hi) {
      for (int i = 0; i < count; ++i) {
        int r;
        bool dup;
        do {
          std::uniform_int_distribution<int> dist(0, hi);
          r = dist(gen);
          dup = false;
          for (int j = 0; j < i; ++j) dup |= subset[j] == r;
        } while (dup);
        subset[i] = r;
      }
    };
    if (t_n_prime < k) draw_unique(m, n - 1);
    else { draw_unique(m - 1, n - 2); subset[m - 1] = std::min(n, N - 1); }
    ++k;
  }
};

template <class Est>
void EstimatePairWith(std::mt19937& gen, const ThbRansacParams& P, const double* data, int n, ThbRelPoseResult* out, uint8_t* mask);
template <class Est>
void EstimatePair(const ThbRansacParams& P, const double* data, int n, uint32_t seed, ThbRelPoseResult* out, uint8_t* mask) {
  std::mt19937 gen(seed);  // RandomNumberGenerator rng(seed) reseeds the process-wide generator (util/random.cc:54-62)
  EstimatePairWith<Est>(gen, P, data, n, out, mask);
}
template <class Est>
void EstimatePairWith(std::mt19937& gen, const ThbRansacParams& P, const double* data, int n, ThbRelPoseResult* out, uint8_t* mask) {
  constexpr int S = Est::S;
  std::memset(out, 0, sizeof(*out));
  out->num_input_data_points = n;
  if (mask) std::memset(mask, 0, n);
  if (n < S) return;  // RandomSampler::Initialize CHECK_GE -> reported as failure instead of aborting
  std::vector<int> sample_indices(n);
  std::iota(sample_indices.begin(), sample_indices.end(), 0);
  const double log_failure_prob = std::log(P.failure_probability);
  double best_cost = DBL_MAX;
  int max_iterations = P.max_iterations;
  if (P.min_inlier_ratio > 0) max_iterations = std::min(ComputeMaxIterations(P, S, P.min_inlier_ratio, log_failure_prob, n), P.max_iterations);
  Model best;
  std::memset(&best, 0, sizeof(best));
  std::vector<int> inl;
  std::vector<double> inl_data;
  int num_lo = 0;
  auto gather = [&](const std::vector<int>& idx) {  // GetInlierDatum
    inl_data.resize(idx.size() * Est::D);
    for (size_t q = 0; q < idx.size(); ++q) for (int k = 0; k < Est::D; ++k) inl_data[q * Est::D + k] = data[Est::D * (size_t)idx[q] + k];
  };
  ProsacSampler prosac;
  if (P.ransac_type == 1) prosac.Initialize(n, S);
  int it;
  for (it = 0; it < max_iterations; ++it) {
    double sample[S * Est::D];
    if (P.ransac_type == 1) {  // RansacType::PROSAC (create_and_initialize_ransac_variant.h)
      int subset[S];
      prosac.Sample(gen, subset);
      for (int i = 0; i < S; ++i) for (int k = 0; k < Est::D; ++k) sample[Est::D * i + k] = data[Est::D * (size_t)subset[i] + k];
    } else {
      for (int i = 0; i < S; ++i) {
        std::uniform_int_distribution<int> dist(i, n - 1);
        std::swap(sample_indices[i], sample_indices[dist(gen)]);
        for (int k = 0; k < Est::D; ++k) sample[Est::D * i + k] = data[Est::D * (size_t)sample_indices[i] + k];
      }
    }
    Model models[Est::MAXM];
    const int nm = Est::Solve(sample, models);
    for (int e = 0; e < nm; ++e) {
      const Model& m = models[e];
      const double cost = Score<Est>(P, data, n, m, &inl);
      const double inlier_ratio = static_cast<double>(inl.size()) / static_cast<double>(n);
      if (cost < best_cost) {
        best = m; best_cost = cost;
        if (inlier_ratio < static_cast<double>(S) / static_cast<double>(n)) continue;
        if (it >= P.lo_start_iterations && P.use_lo) {  // sample_consensus_estimator.h:372-380
          gather(inl);
          if (!Est::Refine(inl_data.data(), static_cast<int>(inl.size()), P.error_thresh, &best)) continue;
          ++num_lo;
        }
        max_iterations = std::min(ComputeMaxIterations(P, S, inlier_ratio, log_failure_prob, n), max_iterations);
      }
    }
  }
  Score<Est>(P, data, n, best, &inl);
  if (P.use_lo) {  // :400-405: the summary's inliers are those of the model BEFORE this last refinement
    gather(inl);
    Est::Refine(inl_data.data(), static_cast<int>(inl.size()), P.error_thresh, &best);
    ++num_lo;
  }
  out->success = 1;
  out->num_iterations = it;
  out->num_inliers = static_cast<int>(inl.size());
  const double ratio = static_cast<double>(inl.size()) / n;
  out->confidence = 1.0 - std::pow(1.0 - std::pow(ratio, static_cast<double>(S)), out->num_iterations);
  out->best_cost = best_cost;
  out->num_lo_iterations = num_lo;
  std::memcpy(out->essential_matrix, best.E, sizeof(best.E));
  std::memcpy(out->rotation, best.R, sizeof(best.R));
  std::memcpy(out->position, best.p, sizeof(best.p));
  if (mask) for (int i : inl) mask[i] = 1;
}


extern "C" int oracle_triangulate_midpoint_batch(const double* org, const double* dir, const int64_t* off, int32_t num_tracks, double* out, uint8_t* ok);
// ---- EstimateTwoViewInfo (calibrated) and TwoViewMatchGeometricVerification::VerifyMatches for one pair ---------------------
// estimate_twoview_info.cc:67-192, two_view_match_geometric_verification.cc:114-366. The two RANSACs of VerifyMatches run on
// ONE generator (homography_params.rng = etvi_options.rng, and every RandomNumberGenerator wraps the same thread_local
// std::mt19937): CountHomographyInliers consumes it first, EstimateRelativePose continues where it stopped.
extern "C" int oracle_ba_solve(const ThbBaProblem* p, const ThbBaOptions* o, ThbBaSummary* s);
extern "C" void oracle_ba_default_options(ThbBaOptions* o);

inline double ScaledThreshold(double t, int w, int h) {  // reconstruction_estimator_utils.cc:97-110
  if (w == 0 && h == 0) return t;
  return t * static_cast<double>(std::max(w, h)) / 1024.0;
}
inline bool AcceptableReprojection(int model, const double* K, const double ext[6], const double X[4], const double* feat, double sq_max) {
  const double adj[3] = {X[0] - X[3] * ext[0], X[1] - X[3] * ext[1], X[2] - X[3] * ext[2]};
  double pc[3], pix[2];
  AngleAxisRotatePoint(ext + 3, adj, pc);
  if (pc[2] / X[3] < 0.0) return false;  // Camera::ProjectPoint < 0
  if (!ProjectByModel<double>(model, K, pc, pix)) return false;
  const double ex = feat[0] - pix[0], ey = feat[1] - pix[1];
  return ex * ex + ey * ey < sq_max;
}

void TwoViewPair(const ThbTwoViewOptions& O, const ThbViewIntrinsics& I1, const ThbViewIntrinsics& I2, const double* px, int n, uint32_t seed,
                 bool verify, ThbTwoViewInfo* info, uint8_t* out_mask) {
  std::memset(info, 0, sizeof(*info));
  if (out_mask) std::memset(out_mask, 0, n);
  if (verify && n < O.min_num_inlier_matches) return;
  std::mt19937 gen(seed);
  ThbRansacParams rp;
  std::memset(&rp, 0, sizeof(rp));
  rp.failure_probability = 1.0 - O.expected_ransac_confidence; rp.min_inlier_ratio = 0.0; rp.min_iterations = O.min_ransac_iterations;
  rp.max_iterations = O.max_ransac_iterations; rp.use_mle = O.use_mle; rp.lo_start_iterations = 50;
  if (verify) {  // CountHomographyInliers: cameras not set up yet -> image size 0 -> unscaled thresholds
    ThbRansacParams hp = rp;
    hp.error_thresh = O.max_sampson_error_pixels * O.max_sampson_error_pixels;
    ThbRelPoseResult hres;
    EstimatePairWith<HomographyEst>(gen, hp, px, n, &hres, nullptr);
    info->num_homography_inliers = hres.num_inliers;
  }
  // NormalizeFeatures
  std::vector<double> norm(4 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double q1[3], q2[3];
    PixelToCamera(I1.model, I1.params, px + 4 * (size_t)i, q1);
    PixelToCamera(I2.model, I2.params, px + 4 * (size_t)i + 2, q2);
    norm[4 * i] = q1[0] / q1[2]; norm[4 * i + 1] = q1[1] / q1[2]; norm[4 * i + 2] = q2[0] / q2[2]; norm[4 * i + 3] = q2[1] / q2[2];
  }
  rp.use_lo = O.use_lo; rp.lo_start_iterations = O.lo_start_iterations;
  rp.error_thresh = ScaledThreshold(O.max_sampson_error_pixels, I1.image_width, I1.image_height) *
                    ScaledThreshold(O.max_sampson_error_pixels, I2.image_width, I2.image_height) / (I1.params[0] * I2.params[0]);
  ThbRelPoseResult res;
  std::vector<uint8_t> inl(n, 0);
  EstimatePairWith<RelPoseEst>(gen, rp, norm.data(), n, &res, inl.data());
  if (!res.success) return;
  EigenRotationMatrixToAngleAxis(res.rotation, info->rotation_2);
  for (int k = 0; k < 3; ++k) info->position_2[k] = res.position[k];
  info->focal_length_1 = I1.params[0]; info->focal_length_2 = I2.params[0];
  info->num_verified_matches = res.num_inliers;
  info->visibility_score = 0;  // inlier list still empty when the score is computed (estimate_twoview_info.cc:186-189)
  info->num_ransac_iterations = res.num_iterations;
  if (!verify) { info->success = 1; if (out_mask) std::memcpy(out_mask, inl.data(), n); return; }
  if (res.num_inliers < O.min_num_inlier_matches) return;
  if (!(O.bundle_adjustment && res.num_inliers > O.min_num_inlier_matches)) {
    if (out_mask) std::memcpy(out_mask, inl.data(), n);
    info->success = res.num_inliers > O.min_num_inlier_matches;
    return;
  }
  // SetupCameras + TriangulatePoints
  double ext1[6] = {0, 0, 0, 0, 0, 0}, ext2[6];
  for (int k = 0; k < 3; ++k) { ext2[k] = info->position_2[k]; ext2[3 + k] = info->rotation_2[k]; }
  const double cos_min = std::cos(O.min_triangulation_angle_degrees * 3.14159265358979323846 / 180.0);
  const double sq_tri = O.triangulation_max_reprojection_error * O.triangulation_max_reprojection_error;
  std::vector<int> idx;
  std::vector<double> pts;
  const double neg_aa[3] = {-ext2[3], -ext2[4], -ext2[5]};
  for (int i = 0; i < n; ++i) {
    if (!inl[i]) continue;
    double q1[3], q2[3], d2[3];
    PixelToCamera(I1.model, I1.params, px + 4 * (size_t)i, q1);
    PixelToCamera(I2.model, I2.params, px + 4 * (size_t)i + 2, q2);
    AngleAxisRotatePoint(neg_aa, q2, d2);  // R^T q2
    const double n1 = std::sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2]), n2 = std::sqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]);
    const double d1[3] = {q1[0] / n1, q1[1] / n1, q1[2] / n1};
    for (double& v : d2) v /= n2;
    if (!(d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2] < cos_min)) continue;
    const double org[6] = {0, 0, 0, ext2[0], ext2[1], ext2[2]}, dir[6] = {d1[0], d1[1], d1[2], d2[0], d2[1], d2[2]};
    const int64_t roff[2] = {0, 2};
    double X[4]; uint8_t ok = 0;
    oracle_triangulate_midpoint_batch(org, dir, roff, 1, X, &ok);
    if (!ok) continue;
    if (!AcceptableReprojection(I1.model, I1.params, ext1, X, px + 4 * (size_t)i, sq_tri) ||
        !AcceptableReprojection(I2.model, I2.params, ext2, X, px + 4 * (size_t)i + 2, sq_tri)) continue;
    idx.push_back(i);
    pts.insert(pts.end(), X, X + 4);
  }
  info->num_triangulated = static_cast<int>(idx.size());
  if (static_cast<int>(idx.size()) < O.min_num_inlier_matches) return;
  // BundleAdjustTwoViews: camera 1 constant, camera 2 free, 4-vector points free, intrinsics constant (calibrated)
  const int m = static_cast<int>(idx.size());
  std::vector<double> cam(12), intr(2 * THB_INTR_STRIDE), xy(4 * (size_t)m);
  std::vector<int32_t> ocam(2 * (size_t)m), opt(2 * (size_t)m), group = {0, 1}, model = {I1.model, I2.model};
  std::vector<uint8_t> cconst = {THB_CAM_CONST_ALL, 0};
  for (int k = 0; k < 6; ++k) { cam[k] = ext1[k]; cam[6 + k] = ext2[k]; }
  for (int k = 0; k < THB_INTR_STRIDE; ++k) { intr[k] = I1.params[k]; intr[THB_INTR_STRIDE + k] = I2.params[k]; }
  for (int q = 0; q < m; ++q) {
    ocam[2 * q] = 0; ocam[2 * q + 1] = 1; opt[2 * q] = q; opt[2 * q + 1] = q;
    for (int k = 0; k < 4; ++k) xy[4 * q + k] = px[4 * (size_t)idx[q] + k];
  }
  ThbBaProblem P;
  std::memset(&P, 0, sizeof(P));
  P.num_cameras = 2; P.num_groups = 2; P.num_points = m; P.num_observations = 2 * m; P.memory_space = THB_MEM_HOST;
  P.cam_ext = cam.data(); P.cam_const = cconst.data(); P.cam_group = group.data(); P.intr = intr.data(); P.intr_model = model.data();
  P.intr_const = nullptr; P.pts = pts.data(); P.pt_const = nullptr; P.obs_cam = ocam.data(); P.obs_pt = opt.data(); P.obs_xy = xy.data();
  ThbBaOptions bo;
  oracle_ba_default_options(&bo);
  bo.use_homogeneous_point_parametrization = 0; bo.use_inner_iterations = 0; bo.max_trust_region_radius = 1e16; bo.max_num_iterations = 100;
  ThbBaSummary sum;
  const int rc = oracle_ba_solve(&P, &bo, &sum);
  info->ba_iterations = sum.num_iterations; info->ba_initial_cost = sum.initial_cost; info->ba_final_cost = sum.final_cost;
  if (rc != THB_OK || !sum.success) return;
  const double sq_fin = O.final_max_reprojection_error * O.final_max_reprojection_error;
  int kept = 0;
  for (int q = 0; q < m; ++q) {
    const bool ok = AcceptableReprojection(I1.model, I1.params, &cam[0], &pts[4 * (size_t)q], px + 4 * (size_t)idx[q], sq_fin) &&
                    AcceptableReprojection(I2.model, I2.params, &cam[6], &pts[4 * (size_t)q], px + 4 * (size_t)idx[q] + 2, sq_fin);
    if (ok) { ++kept; if (out_mask) out_mask[idx[q]] = 1; }
  }
  const double pn = std::sqrt(cam[6] * cam[6] + cam[7] * cam[7] + cam[8] * cam[8]);
  for (int k = 0; k < 3; ++k) { info->rotation_2[k] = cam[9 + k]; info->position_2[k] = cam[6 + k] / pn; }
  info->num_verified_matches = kept;
  info->success = kept > O.min_num_inlier_matches;
}

template <class Est>
int RunBatch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* mask, int32_t threads) {
  if (!b || !p || !results) return THB_E_INVALID_ARGUMENT;
  if (!(p->error_thresh > 0) || !(p->failure_probability > 0 && p->failure_probability < 1) || p->min_inlier_ratio < 0 ||
      p->min_inlier_ratio > 1 || p->max_iterations < p->min_iterations) return THB_E_INVALID_ARGUMENT;
  if ((p->use_lo && !Est::HAS_LO) || p->ransac_type < 0 || p->ransac_type > 2 || (p->ransac_type == 2 && p->use_lo)) return THB_E_UNSUPPORTED;
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
  for (int i = 0; i < b->num_pairs; ++i) {
    const int64_t o = b->pair_offset[i];
    const int n = static_cast<int>(b->pair_offset[i + 1] - o);
    EstimatePair<Est>(*p, b->corr + Est::D * o, n, b->seed[i], results + i, mask ? mask + o : nullptr);
  }
  return THB_OK;
}

}  // namespace
}  // namespace oracle

extern "C" {

void oracle_ransac_default_params(ThbRansacParams* p) {
  std::memset(p, 0, sizeof(*p));
  p->error_thresh = -1.0; p->failure_probability = 0.01; p->min_inlier_ratio = 0.0;
  p->min_iterations = 100; p->max_iterations = 2147483647; p->use_mle = 0; p->use_lo = 0; p->lo_start_iterations = 50;
  p->ransac_type = 0;
}

int oracle_five_point(const double* x1, const double* x2, int32_t count, double* E_out, int32_t* num_solutions) {
  for (int i = 0; i < count; ++i) {
    for (int k = 0; k < 90; ++k) E_out[(size_t)i * 90 + k] = 0.0;
    num_solutions[i] = oracle::FivePoint(x1 + 10 * (size_t)i, x2 + 10 * (size_t)i, E_out + 90 * (size_t)i);
  }
  return THB_OK;
}

int oracle_two_view_batch(const ThbPairBatch* b, const ThbViewIntrinsics* i1, const ThbViewIntrinsics* i2, const ThbTwoViewOptions* o,
                          ThbTwoViewInfo* info, uint8_t* mask, int32_t verify) {
  if (!b || !i1 || !i2 || !o || !info) return THB_E_INVALID_ARGUMENT;
  for (int i = 0; i < b->num_pairs; ++i) if (!i1[i].focal_length_is_set || !i2[i].focal_length_is_set) return THB_E_UNSUPPORTED;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b->num_pairs; ++i) {
    const int64_t off = b->pair_offset[i];
    oracle::TwoViewPair(*o, i1[i], i2[i], b->corr + 4 * off, static_cast<int>(b->pair_offset[i + 1] - off), b->seed[i], verify != 0, info + i,
                        mask ? mask + off : nullptr);
  }
  return THB_OK;
}
int oracle_ransac_relpose_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* mask, int32_t threads) {
  return oracle::RunBatch<oracle::RelPoseEst>(b, p, results, mask, threads);
}
int oracle_ransac_abspose_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* mask, int32_t threads) {
  return oracle::RunBatch<oracle::AbsPoseEst>(b, p, results, mask, threads);
}
int oracle_ransac_homography_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* mask, int32_t threads) {
  return oracle::RunBatch<oracle::HomographyEst>(b, p, results, mask, threads);
}
int oracle_p3p(const double* feat, const double* world, int32_t count, double* R_out, double* t_out, int32_t* num_solutions) {
  for (int i = 0; i < count; ++i) {
    for (int k = 0; k < 36; ++k) R_out[(size_t)i * 36 + k] = 0.0;
    for (int k = 0; k < 12; ++k) t_out[(size_t)i * 12 + k] = 0.0;
    num_solutions[i] = oracle::P3P(feat + 6 * (size_t)i, world + 9 * (size_t)i, R_out + 36 * (size_t)i, t_out + 12 * (size_t)i);
  }
  return THB_OK;
}
// theia::TriangulateMidpoint (sfm/triangulation/triangulation.cc:130-157): A = sum (I4 - d d^T), b = sum (I4 - d d^T) (o, 1),
// Eigen::LLT<Matrix4d> solve (unblocked lower Cholesky, forward / backward substitution).
int oracle_triangulate_midpoint_batch(const double* org, const double* dir, const int64_t* off, int32_t num_tracks, double* out,
                                      uint8_t* ok) {
  for (int t = 0; t < num_tracks; ++t) {
    double A[4][4] = {}, b[4] = {};
    for (int64_t q = off[t]; q < off[t + 1]; ++q) {
      const double d[4] = {dir[3 * q], dir[3 * q + 1], dir[3 * q + 2], 0.0};
      const double o[4] = {org[3 * q], org[3 * q + 1], org[3 * q + 2], 1.0};
      for (int r = 0; r < 4; ++r) {
        double s = 0.0;
        for (int c = 0; c < 4; ++c) { const double T = (r == c ? 1.0 : 0.0) - d[r] * d[c]; A[r][c] += T; s += T * o[c]; }
        b[r] += s;
      }
    }
    bool good = off[t + 1] - off[t] >= 2;
    double L[4][4] = {};
    for (int k = 0; k < 4; ++k) {
      double x = A[k][k];
      for (int j = 0; j < k; ++j) x -= L[k][j] * L[k][j];
      if (!(x > 0.0)) good = false;
      x = std::sqrt(x);
      L[k][k] = x;
      for (int r = k + 1; r < 4; ++r) {
        double v = A[r][k];
        for (int j = 0; j < k; ++j) v -= L[r][j] * L[k][j];
        L[r][k] = v / x;
      }
    }
    double y[4], z[4];
    for (int r = 0; r < 4; ++r) { double v = b[r]; for (int j = 0; j < r; ++j) v -= L[r][j] * y[j]; y[r] = v / L[r][r]; }
    for (int r = 3; r >= 0; --r) { double v = y[r]; for (int j = r + 1; j < 4; ++j) v -= L[j][r] * z[j]; z[r] = v / L[r][r]; }
    for (int r = 0; r < 4; ++r) out[4 * (size_t)t + r] = good ? z[r] : 0.0;
    if (ok) ok[t] = good ? 1 : 0;
  }
  return THB_OK;
}
int oracle_four_point_homography(const double* corr, int32_t count, double* H_out, int32_t* ok) {
  for (int i = 0; i < count; ++i) ok[i] = oracle::FourPointH(corr + 16 * (size_t)i, H_out + 9 * (size_t)i) ? 1 : 0;
  return THB_OK;
}
int oracle_seven_point_fundamental(const double* corr, int32_t count, double* F_out, int32_t* num_solutions) {
  for (int i = 0; i < count; ++i) {
    for (int k = 0; k < 27; ++k) F_out[(size_t)i * 27 + k] = 0.0;
    num_solutions[i] = oracle::SevenPointF(corr + 28 * (size_t)i, F_out + 27 * (size_t)i);
  }
  return THB_OK;
}
int oracle_poly_roots(const double* poly, int32_t size, double* re, double* im) { return oracle::PolyRoots(poly, size, re, im); }

// helpers exposed for the unit tests of the restated decompositions
void oracle_jacobi_svd3(const double* A, double* U, double* S, double* V) { oracle::JacobiSVD3(A, U, S, V); }
int oracle_eigen10(const double* A, double* re, double* im, double* vec) {
  oracle::EigenSolverReal<10> es;
  es.compute(A);
  for (int i = 0; i < 10; ++i) { re[i] = es.eig_re[i]; im[i] = es.eig_im[i]; }
  for (int i = 0; i < 100; ++i) vec[i] = es.vec[i];
  return es.ok ? 0 : 1;
}
int oracle_fullpivlu_kernel_5x9(const double* A, double* ker) {
  oracle::FullPivLU<5, 9> lu;
  lu.compute(A);
  const int dk = lu.dimensionOfKernel();
  if (dk == 4) lu.kernel(ker);
  return dk;
}
void oracle_fullpivlu_solve10(const double* A, const double* B, double* X) {
  oracle::FullPivLU<10, 10> lu;
  lu.compute(A);
  lu.solve<10>(B, X);
}
double oracle_sampson(const double* F, const double* c) { return oracle::Sampson(F, c); }
int oracle_best_pose(const double* E, const double* corr, int n, double* R, double* p) { return oracle::BestPose(E, corr, n, R, p); }

}  // extern "C"
