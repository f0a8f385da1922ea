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

struct Model { double E[9], R[9], p[3]; };

// sample_consensus_estimator.h:251-297
int ComputeMaxIterations(const ThbRansacParams& P, double min_sample_size, double inlier_ratio, double log_failure_prob, int total) {
  if (inlier_ratio == 1.0) return P.min_iterations;
  const int ninl = static_cast<int>(inlier_ratio * total);
  const double num_samples = min_sample_size;
  double a = 1.0, b = 1.0;
  for (int i = 0; i < num_samples; ++i) { a *= ninl - i; b *= total - i; }
  const double prob_all_inliers = a / b;
  if (prob_all_inliers < DBL_EPSILON) return P.max_iterations;
  if (prob_all_inliers >= 1.0 - DBL_EPSILON) return P.min_iterations;
  const double num_iterations = log_failure_prob / std::log(1.0 - prob_all_inliers);
  return std::max(static_cast<double>(P.min_iterations), std::min(num_iterations, static_cast<double>(P.max_iterations)));
}

double Score(const ThbRansacParams& P, const double* corr, int n, const Model& m, std::vector<int>* inliers) {
  inliers->clear();
  double cost = 0.0;
  for (int i = 0; i < n; ++i) {
    const double r = InFront(corr + 4 * i, m.R, m.p) ? Sampson(m.E, corr + 4 * i) : DBL_MAX;
    if (P.use_mle) {
      if (r < P.error_thresh) { cost += r; inliers->push_back(i); } else cost += P.error_thresh;
    } else if (r < P.error_thresh) {
      inliers->push_back(i);
    }
  }
  return P.use_mle ? cost : static_cast<double>(n - static_cast<int>(inliers->size()));
}

void EstimatePair(const ThbRansacParams& P, const double* corr, int n, uint32_t seed, ThbRelPoseResult* out, uint8_t* mask) {
  std::memset(out, 0, sizeof(*out));
  out->num_input_data_points = n;
  if (mask) std::memset(mask, 0, n);
  if (n < 5) return;  // RandomSampler::Initialize CHECK_GE -> reported as failure instead of aborting
  std::mt19937 gen(seed);
  std::vector<int> sample_indices(n);
  std::iota(sample_indices.begin(), sample_indices.end(), 0);
  const double log_failure_prob = std::log(P.failure_probability);
  double best_cost = DBL_MAX;
  int max_iterations = P.max_iterations;
  if (P.min_inlier_ratio > 0) max_iterations = std::min(ComputeMaxIterations(P, 5, P.min_inlier_ratio, log_failure_prob, n), P.max_iterations);
  Model best;
  std::memset(&best, 0, sizeof(best));
  std::vector<int> inl;
  int it;
  for (it = 0; it < max_iterations; ++it) {
    int idx[5];
    for (int i = 0; i < 5; ++i) {
      std::uniform_int_distribution<int> dist(i, n - 1);
      std::swap(sample_indices[i], sample_indices[dist(gen)]);
      idx[i] = sample_indices[i];
    }
    double x1[10], x2[10], sc[20];
    for (int i = 0; i < 5; ++i) {
      const double* c = corr + 4 * (size_t)idx[i];
      x1[2 * i] = c[0]; x1[2 * i + 1] = c[1]; x2[2 * i] = c[2]; x2[2 * i + 1] = c[3];
      for (int k = 0; k < 4; ++k) sc[4 * i + k] = c[k];
    }
    double Es[90];
    const int ne = FivePoint(x1, x2, Es);
    for (int e = 0; e < ne; ++e) {
      Model m;
      std::memcpy(m.E, Es + 9 * e, sizeof(m.E));
      if (BestPose(m.E, sc, 5, m.R, m.p) < 4) continue;
      const double cost = Score(P, corr, n, m, &inl);
      const double inlier_ratio = static_cast<double>(inl.size()) / static_cast<double>(n);
      if (cost < best_cost) {
        best = m; best_cost = cost;
        if (inlier_ratio < 5.0 / static_cast<double>(n)) continue;
        max_iterations = std::min(ComputeMaxIterations(P, 5, inlier_ratio, log_failure_prob, n), max_iterations);
      }
    }
  }
  Score(P, corr, n, best, &inl);
  out->success = 1;
  out->num_iterations = it;
  out->num_inliers = static_cast<int>(inl.size());
  const double ratio = static_cast<double>(inl.size()) / n;
  out->confidence = 1.0 - std::pow(1.0 - std::pow(ratio, 5.0), out->num_iterations);
  out->best_cost = best_cost;
  std::memcpy(out->essential_matrix, best.E, sizeof(best.E));
  std::memcpy(out->rotation, best.R, sizeof(best.R));
  std::memcpy(out->position, best.p, sizeof(best.p));
  if (mask) for (int i : inl) mask[i] = 1;
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

int oracle_ransac_relpose_batch(const ThbPairBatch* b, const ThbRansacParams* p, ThbRelPoseResult* results, uint8_t* mask, int32_t threads) {
  if (!b || !p || !results) return THB_E_INVALID_ARGUMENT;
  if (!(p->error_thresh > 0) || !(p->failure_probability > 0 && p->failure_probability < 1) || p->min_inlier_ratio < 0 ||
      p->min_inlier_ratio > 1 || p->max_iterations < p->min_iterations) return THB_E_INVALID_ARGUMENT;
  if (p->use_lo || p->ransac_type != 0) return THB_E_UNSUPPORTED;
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
  for (int i = 0; i < b->num_pairs; ++i) {
    const int64_t o = b->pair_offset[i];
    const int n = static_cast<int>(b->pair_offset[i + 1] - o);
    oracle::EstimatePair(*p, b->corr + 4 * o, n, b->seed[i], results + i, mask ? mask + o : nullptr);
  }
  return THB_OK;
}

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
