// Hot path 2: RANSAC two-view relative-pose verification, one CTA per image pair, whole RANSAC on device.
//
// Stands behind theia::EstimateRelativePose (sfm/estimators/estimate_relative_pose.cc:159-172) =
// SampleConsensusEstimator<RelativePoseEstimator>::Estimate (solvers/sample_consensus_estimator.h:299-415)
// with RandomSampler (solvers/random_sampler.cc:53-72) on std::mt19937 (util/random.cc:46-84).
//
// The reference loop is sequential (persistent sampler permutation, strict-< best update in model order,
// adaptive iteration bound). It is replayed exactly: iterations are processed in batches of 32 —
//   draw   : one thread advances a bit-exact mt19937 + libstdc++ uniform_int_distribution (Lemire) and the
//            partial Fisher-Yates permutation, 5 indices per iteration
//   solve  : one THREAD per hypothesis runs the five-point solver (FP64 SIMT is issue-bound, so 32 hypotheses
//            per warp cost the same as one) and the essential-matrix decomposition + cheirality vote
//   score  : one WARP per model scores every correspondence (cheirality-gated Sampson) from shared memory,
//            lane-strided, warp-shuffle reduction; a model whose partial cost already exceeds the best cost
//            known at batch start is abandoned (it can never win the strict-< test: exactness is preserved)
//   scan   : one thread walks the (iteration, model) costs in order, updates the best model and the adaptive
//            bound, and discards everything past the terminating iteration
// This file is compiled with -fmad=false (see small_linalg.cuh); the scoring formulas use explicit fma() in a
// fixed order so that they are both fast and bit-reproducible against the CPU oracle.
#include <cfloat>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "small_linalg.cuh"

namespace thb {
namespace {

constexpr int RT = 256;   // threads per CTA
constexpr int NW = RT / 32;
constexpr int BI = 32;    // iterations per batch
constexpr int MAXM = 10;  // five-point solutions per sample

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
    sl::kernel_5x9(lu, ns);
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
  double L[100], Rm[100], X[100];
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

// SampleConsensusEstimator::ComputeMaxIterations (sample_consensus_estimator.h:251-297)
__device__ int compute_max_iterations(const ThbRansacParams& P, double min_sample_size, double inlier_ratio,
                                      double log_failure_prob, int total) {
  if (inlier_ratio == 1.0) return P.min_iterations;
  const int ninl = (int)(inlier_ratio * total);
  double a = 1.0, b = 1.0;
  for (int i = 0; i < min_sample_size; ++i) { a *= ninl - i; b *= total - i; }
  const double prob_all_inliers = a / b;
  if (prob_all_inliers < DBL_EPSILON) return P.max_iterations;
  if (prob_all_inliers >= 1.0 - DBL_EPSILON) return P.min_iterations;
  const double num_iterations = log_failure_prob / log(1.0 - prob_all_inliers);
  return (int)fmax((double)P.min_iterations, fmin(num_iterations, (double)P.max_iterations));
}

// Warp-wide score of one model over all correspondences. Returns (cost, #inliers) in every lane; cost = +inf if
// the model was abandoned because its partial cost reached `bail`. If mask != nullptr the inlier flags are written.
__device__ void score_model(const ThbRansacParams& P, const double4* __restrict__ corr, int n, const Model& m, double bail,
                            uint8_t* __restrict__ mask, double* cost_out, int* ninl_out) {
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
      const double4 c = corr[i];
      const double r = in_front(c.x, c.y, c.z, c.w, R, p) ? sampson(E, c.x, c.y, c.z, c.w) : DBL_MAX;
      const bool inl = r < thresh;
      if (P.use_mle) cost += inl ? r : thresh; else cost += inl ? 0.0 : 1.0;
      ninl += inl ? 1 : 0;
      if (mask) mask[i] = inl ? 1 : 0;
    }
    if ((s & 15) == 15 && bail < DBL_MAX) {
      const double partial = warp_sum(cost);
      if (partial >= bail) { *cost_out = INFINITY; *ninl_out = 0; return; }
    }
  }
  cost = warp_sum(cost);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ninl += __shfl_xor_sync(0xffffffffu, ninl, o);
  *cost_out = cost;
  *ninl_out = ninl;
}

struct RansacShared {
  Model models[BI * MAXM];
  Model best;
  double cost[BI * MAXM];
  int ninl[BI * MAXM];
  int nmodels[BI];
  int model_start[BI + 1];
  int samples[BI][5];
  Mt19937 rng;
  double best_cost;
  int max_iterations, it0, finished, num_iterations, have_best;
};

__global__ void __launch_bounds__(RT) k_ransac_relpose(ThbRansacParams P, int num_pairs, const long long* __restrict__ pair_offset,
                                                       const double* __restrict__ corr_all, const uint32_t* __restrict__ seed,
                                                       ThbRelPoseResult* __restrict__ results, uint8_t* __restrict__ mask_all,
                                                       int* __restrict__ idx_ws, int smem_corr_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RansacShared& S = *reinterpret_cast<RansacShared*>(smem_raw);
  double4* s_corr = reinterpret_cast<double4*>(smem_raw + ((sizeof(RansacShared) + 31) / 32) * 32);
  const int pair = blockIdx.x;
  if (pair >= num_pairs) return;
  const long long off = pair_offset[pair];
  const int n = (int)(pair_offset[pair + 1] - off);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  ThbRelPoseResult* out = results + pair;
  uint8_t* mask = mask_all ? mask_all + off : nullptr;
  if (n < 5) {  // RandomSampler::Initialize would CHECK-abort; reported as failure
    if (t == 0) { memset(out, 0, sizeof(*out)); out->num_input_data_points = n; }
    if (mask) for (int i = t; i < n; i += RT) mask[i] = 0;
    return;
  }
  const double4* g_corr = reinterpret_cast<const double4*>(corr_all) + off;
  const bool in_smem = n <= smem_corr_cap;
  const double4* corr = g_corr;
  if (in_smem) {
    for (int i = t; i < n; i += RT) s_corr[i] = g_corr[i];
    corr = s_corr;
  }
  int* sidx = idx_ws + off;  // RandomSampler::sample_indices_ (persistent permutation)
  for (int i = t; i < n; i += RT) sidx[i] = i;
  const double log_failure_prob = log(P.failure_probability);
  if (t == 0) {
    mt_seed(&S.rng, seed[pair]);
    S.best_cost = DBL_MAX;
    S.max_iterations = P.max_iterations;
    if (P.min_inlier_ratio > 0) {
      const int mi = compute_max_iterations(P, 5, P.min_inlier_ratio, log_failure_prob, n);
      S.max_iterations = mi < P.max_iterations ? mi : P.max_iterations;
    }
    S.it0 = 0; S.finished = 0; S.num_iterations = 0; S.have_best = 0;
    memset(&S.best, 0, sizeof(Model));
  }
  __syncthreads();
  while (true) {
    if (S.it0 >= S.max_iterations) { if (t == 0) { S.num_iterations = S.it0; S.finished = 1; } }
    __syncthreads();
    if (S.finished) break;
    const int it0 = S.it0;
    const int nit = min(BI, S.max_iterations - it0);
    // ---- draw
    if (t == 0) {
      for (int b = 0; b < nit; ++b)
        for (int i = 0; i < 5; ++i) {
          const int j = mt_uniform_int(&S.rng, i, n - 1);
          const int a = sidx[i], c = sidx[j];
          sidx[i] = c; sidx[j] = a;
          S.samples[b][i] = c;
        }
    }
    __syncthreads();
    // ---- solve: thread b of warp 0 = iteration it0 + b
    if (w == 0) {
      int nm = 0;
      if (lane < nit) {
        double x1[10], x2[10], sc[20], Es[90];
        for (int i = 0; i < 5; ++i) {
          const double4 c = corr[S.samples[lane][i]];
          x1[2 * i] = c.x; x1[2 * i + 1] = c.y; x2[2 * i] = c.z; x2[2 * i + 1] = c.w;
          sc[4 * i] = c.x; sc[4 * i + 1] = c.y; sc[4 * i + 2] = c.z; sc[4 * i + 3] = c.w;
        }
        const int ne = five_point(x1, x2, Es);
        for (int e = 0; e < ne; ++e) {
          Model m;
          for (int k = 0; k < 9; ++k) m.E[k] = Es[9 * e + k];
          if (best_pose(m.E, sc, 5, m.R, m.p) < 4) continue;
          S.models[lane * MAXM + nm] = m;
          ++nm;
        }
      }
      if (lane < BI) S.nmodels[lane] = nm;
      __syncwarp();
      if (lane == 0) {
        int acc = 0;
        for (int b = 0; b < BI; ++b) { S.model_start[b] = acc; acc += (b < nit) ? S.nmodels[b] : 0; }
        S.model_start[BI] = acc;
      }
    }
    __syncthreads();
    // ---- score: warp w takes flat models w, w + NW, ...
    const int total_models = S.model_start[BI];
    const double bail = S.best_cost;
    for (int j = w; j < total_models; j += NW) {
      int b = 0;
      while (S.model_start[b + 1] <= j) ++b;
      const int k = j - S.model_start[b];
      double cost; int ninl;
      score_model(P, corr, n, S.models[b * MAXM + k], bail, nullptr, &cost, &ninl);
      if (lane == 0) { S.cost[b * MAXM + k] = cost; S.ninl[b * MAXM + k] = ninl; }
    }
    __syncthreads();
    // ---- scan in (iteration, model) order
    if (t == 0) {
      int it = it0;
      for (int b = 0; b < nit; ++b, ++it) {
        if (it >= S.max_iterations) break;
        for (int k = 0; k < S.nmodels[b]; ++k) {
          const double sample_cost = S.cost[b * MAXM + k];
          if (sample_cost < S.best_cost) {
            const double inlier_ratio = (double)S.ninl[b * MAXM + k] / (double)n;
            S.best = S.models[b * MAXM + k];
            S.best_cost = sample_cost;
            S.have_best = 1;
            if (inlier_ratio < 5.0 / (double)n) continue;
            const int mi = compute_max_iterations(P, 5, inlier_ratio, log_failure_prob, n);
            if (mi < S.max_iterations) S.max_iterations = mi;
          }
        }
      }
      S.it0 = it;
    }
    __syncthreads();
  }
  // ---- final inliers of the best model (sample_consensus_estimator.h:396-414)
  if (w == 0) {
    double cost; int ninl;
    score_model(P, corr, n, S.best, DBL_MAX, mask, &cost, &ninl);
    if (lane == 0) {
      out->success = 1;
      out->num_inliers = ninl;
      out->num_iterations = S.num_iterations;
      out->num_input_data_points = n;
      const double ratio = (double)ninl / (double)n;
      out->confidence = 1.0 - pow(1.0 - pow(ratio, 5.0), (double)S.num_iterations);
      out->best_cost = S.best_cost;
      for (int k = 0; k < 9; ++k) { out->essential_matrix[k] = S.best.E[k]; out->rotation[k] = S.best.R[k]; }
      for (int k = 0; k < 3; ++k) out->position[k] = S.best.p[k];
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

struct Bufs {
  std::vector<void*> p;
  ~Bufs() { for (void* q : p) cudaFree(q); }
  template <typename T> T* get(size_t n) {
    void* q = nullptr;
    if (cudaMalloc(&q, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    p.push_back(q);
    return (T*)q;
  }
};

int check_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) THB_FAIL(THB_E_NO_DEVICE, "no CUDA device visible; libtheia_b200 has no CPU path");
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp pr;
  THB_CUDA_CHECK(cudaGetDeviceProperties(&pr, dev));
  if (pr.major != 10) THB_FAIL(THB_E_NO_DEVICE, "device is not sm_100 (B200); kernels are built for sm_100a only");
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
  if (!b || !p || !results) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  // the reference CHECK-aborts on these (sample_consensus_estimator.h:217-223)
  if (!(p->error_thresh > 0) || !(p->failure_probability > 0 && p->failure_probability < 1) || p->min_inlier_ratio < 0 ||
      p->min_inlier_ratio > 1 || p->max_iterations < p->min_iterations) THB_FAIL(THB_E_INVALID_ARGUMENT, "invalid RansacParameters");
  if (p->use_lo) THB_FAIL(THB_E_UNSUPPORTED, "use_lo (LO-RANSAC refinement by two-view BA) is not implemented");
  if (p->ransac_type != 0) THB_FAIL(THB_E_UNSUPPORTED, "only RansacType::RANSAC is implemented");
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
  int max_n = 0;
  for (int i = 0; i < np; ++i) {
    const long long n = h_off[i + 1] - h_off[i];
    if (n < 0 || n > 2147483647LL || h_off[0] != 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "pair_offset must start at 0 and be non-decreasing");
    if ((int)n > max_n) max_n = (int)n;
  }
  const long long total = h_off[np];
  if (total > 0 && !b->corr) THB_FAIL(THB_E_INVALID_ARGUMENT, "null corr");
  Bufs B;
  const long long* d_off; const double* d_corr; const uint32_t* d_seed; ThbRelPoseResult* d_res; uint8_t* d_mask = nullptr;
  if (host) {
    long long* o = B.get<long long>(np + 1); double* c = B.get<double>((size_t)total * 4); uint32_t* s = B.get<uint32_t>(np);
    d_res = B.get<ThbRelPoseResult>(np);
    if (inlier_mask) d_mask = B.get<uint8_t>((size_t)total);
    if (!o || !c || !s || !d_res || (inlier_mask && !d_mask)) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
    THB_CUDA_CHECK(cudaMemcpyAsync(o, b->pair_offset, sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(c, b->corr, sizeof(double) * 4 * total, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(s, b->seed, sizeof(uint32_t) * np, cudaMemcpyHostToDevice, st));
    d_off = o; d_corr = c; d_seed = s;
  } else {
    d_off = (const long long*)b->pair_offset; d_corr = b->corr; d_seed = b->seed; d_res = results; d_mask = inlier_mask;
  }
  int* d_idx = B.get<int>((size_t)total);
  if (!d_idx) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  // shared memory: control block + as many correspondences as fit (two CTAs per SM when the pair is small enough)
  const size_t ctrl = ((sizeof(RansacShared) + 31) / 32) * 32;
  size_t want = ctrl + (size_t)max_n * sizeof(double4);
  const size_t limit = 227 * 1024;
  int cap = max_n;
  if (want > limit) { cap = (int)((limit - ctrl) / sizeof(double4)); want = ctrl + (size_t)cap * sizeof(double4); }
  THB_CUDA_CHECK(cudaFuncSetAttribute(k_ransac_relpose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
  k_ransac_relpose<<<np, RT, want, st>>>(*p, np, d_off, d_corr, d_seed, d_res, d_mask, d_idx, cap);
  THB_CUDA_CHECK(cudaGetLastError());
  if (host) {
    THB_CUDA_CHECK(cudaMemcpyAsync(results, d_res, sizeof(ThbRelPoseResult) * np, cudaMemcpyDeviceToHost, st));
    if (inlier_mask) THB_CUDA_CHECK(cudaMemcpyAsync(inlier_mask, d_mask, (size_t)total, cudaMemcpyDeviceToHost, st));
  }
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

int thb_five_point_relative_pose(const double* x1, const double* x2, int32_t count, double* E_out, int32_t* num_solutions,
                                 void* cuda_stream) {
  if (!x1 || !x2 || !E_out || !num_solutions || count < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  if (count == 0) return THB_OK;
  int rc = check_device();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  Bufs B;
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
