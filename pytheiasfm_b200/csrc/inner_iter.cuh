// Inner iterations of the bundle-adjustment trust-region loop: BundleAdjustmentOptions::use_inner_iterations (the
// reference default, sfm/bundle_adjustment/bundle_adjustment.h:144) -> ceres::Solver::Options::use_inner_iterations with
// Theia's reversed ordering (bundle_adjuster.cc:329-334): after every trust-region candidate, each camera-extrinsics
// block, then each shared intrinsics block, then each point is minimised on its own with all other blocks constant
// (ceres CoordinateDescentMinimizer). The blocks of one group share no residual, so a group is one
// batched launch:
//   cameras     k_inner_cam   one CTA per camera runs the whole Levenberg-Marquardt loop of its 6 (or 3) coordinates over
//                             its camera-major observation range; J^T J (21), J^T r (6) and the cost are block-reduced
//                             in a fixed order (deterministic), thread 0 takes the trust-region decisions
//   intrinsics  k_inner_intr  one launch per evaluation over the group's observations (a group owns up to every
//                             observation of the problem); the <= 8 groups' loops are driven from the host
//   points      k_track_ba    (track_ba.cuh) one thread per point, the same kernel that stands behind BundleAdjustTrack
// Every inner solve is a TrustRegionMinimizer with default Minimizer::Options and a LevenbergMarquardtStrategy with
// default TrustRegionStrategy::Options, which is what CoordinateDescentMinimizer::Solve constructs: InnerLmParams.
#ifndef THB_INNER_ITER_CUH_
#define THB_INNER_ITER_CUH_

#include "ba_kernels.cuh"

namespace thb {

struct InnerLmParams {
  int max_num_iterations = 50, max_invalid = 5;
  double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, radius0 = 1e4, min_radius = 1e-32, max_radius = 1e32;
  double min_relative_decrease = 1e-3, min_diag = 1e-6, max_diag = 1e32;
};

// ---- cameras ------------------------------------------------------------------------------------------------------------
constexpr int IC_THREADS = 128;
constexpr int IC_VALS = 29;  // 21 (J^T J lower) + 6 (J^T r) + cost + failures

template <int PD, bool WANT_J>
__device__ __forceinline__ void inner_cam_pass(const BaConst& K, const BaState& S, const ObsSoA& O, int c, int q0, int q1, const double* rec,
                                               const double* scale, double (*red)[IC_VALS], double* tot, const double* prior = nullptr, int prior_kinds = 0) {
  double v[IC_VALS];
#pragma unroll
  for (int k = 0; k < IC_VALS; ++k) v[k] = 0.0;
  for (int q = q0 + threadIdx.x; q < q1; q += IC_THREADS) {
    const int p = O.pt[q];
    const double2 xy = O.xy[q], si = O.si[q];
    if (WANT_J) {
      double r[2], jc[12], jp[2 * PD], hc = 0.0;
      if (!eval_obs<-1, PD, 0, true>(K, S, c, p, xy, si, nullptr, nullptr, nullptr, r, jc, jp, nullptr, &hc, 0, 0, 0, 0, nullptr, rec)) { v[28] += 1.0; continue; }
#pragma unroll
      for (int k = 0; k < 6; ++k) { jc[k] *= scale[k]; jc[6 + k] *= scale[k]; }  // scale = 0 on constant coordinates
      int e = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) v[e++] += jc[a] * jc[b] + jc[6 + a] * jc[6 + b];
#pragma unroll
      for (int a = 0; a < 6; ++a) v[21 + a] += jc[a] * r[0] + jc[6 + a] * r[1];
      v[27] += hc;
    } else {
      double r[2];
      if (!eval_residual<-1>(K, S, c, p, xy, si, r, rec)) { v[28] += 1.0; continue; }
      const double sq = r[0] * r[0] + r[1] * r[1];
      if (K.loss_type == THB_LOSS_TRIVIAL) v[27] += 0.5 * sq;
      else { double rho[3]; eval_loss(K.loss_type, K.loss_width, sq, rho); v[27] += 0.5 * rho[0]; }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = WANT_J ? 0 : 27; k < IC_VALS; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) red[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < IC_VALS) {
    double s = 0.0;
    for (int ww = 0; ww < IC_THREADS / 32; ++ww) s += red[ww][threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (prior && threadIdx.x == 0) {  // the camera's prior blocks are residual blocks of this parameter block as well
    for (int kind = 0; kind < PRIOR_KINDS; ++kind) {
      if (!(prior_kinds & (1 << kind))) continue;
      double r[3], J[3][6];
      cam_prior_eval(kind, prior, rec, r, J);
      tot[27] += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      if (WANT_J) {
        for (int k = 0; k < 3; ++k) for (int a = 0; a < 6; ++a) J[k][a] *= scale[a];
        for (int a = 0; a < 6; ++a) {
          tot[21 + a] += J[0][a] * r[0] + J[1][a] * r[1] + J[2][a] * r[2];
          for (int b = 0; b <= a; ++b) tot[a * (a + 1) / 2 + b] += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b];
        }
      }
    }
  }
  if (prior) __syncthreads();
}

// S: the candidate state; S.cam[c] is refined in place (S.camd is re-derived by the caller afterwards).
template <int PD>
__global__ void __launch_bounds__(IC_THREADS) k_inner_cam(BaConst K, BaState S, ObsSoA O, const int* __restrict__ cam_start, InnerLmParams P,
                                                          const uint8_t* __restrict__ has_prior, const double* __restrict__ prior_all) {
  const int c = blockIdx.x;
  const int cc = K.cam_const[c];
  const int q0 = cam_start[c], q1 = cam_start[c + 1];
  if (cc == THB_CAM_CONST_ALL || q1 == q0) return;
  __shared__ double rec[CAMD], x[6], cand[6], scale[6], tot[IC_VALS], red[IC_THREADS / 32][IC_VALS];
  __shared__ int s_go;  // 0: finished, 1: evaluate the candidate, 2: re-evaluate the Jacobian at x
  const int t = threadIdx.x;
  if (t < 6) {
    x[t] = S.cam[6 * (size_t)c + t];
    scale[t] = (cc & (t < 3 ? THB_CAM_CONST_POSITION : THB_CAM_CONST_ORIENTATION)) ? 0.0 : 1.0;
  }
  __syncthreads();
  if (t == 0) cam_derive_record(x, rec);
  __syncthreads();
  const int prior_kinds = has_prior ? has_prior[c] : 0;
  const double* prior = prior_kinds ? prior_all + PRIOR_STRIDE * (size_t)c : nullptr;
  inner_cam_pass<PD, true>(K, S, O, c, q0, q1, rec, scale, red, tot, prior, prior_kinds);  // column norms of the unscaled Jacobian
  if (tot[28] > 0.0) return;  // IterationZero failed: the block stays as it is
  if (t == 0) {
    int e = 0;
    for (int a = 0; a < 6; ++a) { e += a; scale[a] = scale[a] == 0.0 ? 0.0 : 1.0 / (1.0 + sqrt(tot[e])); ++e; }
  }
  __syncthreads();
  inner_cam_pass<PD, true>(K, S, O, c, q0, q1, rec, scale, red, tot, prior, prior_kinds);
  // trust-region state (thread 0)
  double H[21], g[6], diag[6], x_cost = tot[27], x_norm = 0.0, radius = P.radius0, decrease_factor = 2.0, mcc = 0.0;
  bool step_ok = true, reuse_diag = false;
  int iteration = 0, invalid = 0;
  if (t == 0) {
    for (int k = 0; k < 21; ++k) H[k] = tot[k];
    for (int k = 0; k < 6; ++k) { g[k] = tot[21 + k]; x_norm += x[k] * x[k]; }
    x_norm = sqrt(x_norm);
  }
  for (;;) {
    if (t == 0) {
      int go = 1;
      for (;;) {  // until a valid step is found or the loop terminates
        if (iteration >= P.max_num_iterations) { go = 0; break; }
        if (step_ok) {
          double gmax = 0.0;
          for (int k = 0; k < 6; ++k) if (scale[k] != 0.0) gmax = fmax(gmax, fabs(g[k] / scale[k]));
          if (gmax <= P.gtol) { go = 0; break; }
        }
        if (radius <= P.min_radius) { go = 0; break; }
        ++iteration;
        step_ok = false;
        if (!reuse_diag) { int e = 0; for (int a = 0; a < 6; ++a) { e += a; diag[a] = fmin(fmax(H[e], P.min_diag), P.max_diag); ++e; } }
        reuse_diag = true;
        double M[36], y[6];
        { int e = 0; for (int a = 0; a < 6; ++a) for (int b = 0; b <= a; ++b) { M[a * 6 + b] = H[e]; M[b * 6 + a] = H[e]; ++e; } }
        for (int a = 0; a < 6; ++a) M[a * 6 + a] += diag[a] / radius;
        bool valid = spd_solve<6>(M, g, y);
        if (valid) {
          // model cost change -(J s)^T (r + J s / 2), s = -y: y^T g - y^T (J^T J) y / 2
          double yg = 0.0, yHy = 0.0;
          int e = 0;
          for (int a = 0; a < 6; ++a) {
            yg += y[a] * g[a];
            for (int b = 0; b <= a; ++b) { yHy += (a == b ? 1.0 : 2.0) * y[a] * H[e] * y[b]; ++e; }
          }
          mcc = yg - 0.5 * yHy;
          valid = isfinite(mcc) && mcc > 0.0;
        }
        if (!valid) {
          if (++invalid >= P.max_invalid) { go = 0; break; }
          radius /= decrease_factor; decrease_factor *= 2.0;
          continue;
        }
        invalid = 0;
        for (int k = 0; k < 6; ++k) cand[k] = x[k] + (-y[k] * scale[k]);
        cam_derive_record(cand, rec);
        break;
      }
      s_go = go;
    }
    __syncthreads();
    if (s_go == 0) break;
    inner_cam_pass<PD, false>(K, S, O, c, q0, q1, rec, scale, red, tot, prior, prior_kinds);
    if (t == 0) {
      int go = 1;  // 1: next trust-region iteration without a new Jacobian, 2: accepted, 0: converged
      const double cand_cost = tot[28] > 0.0 ? 1.7976931348623157e308 : tot[27];
      double sn = 0.0, cn = 0.0;
      for (int k = 0; k < 6; ++k) { sn += (cand[k] - x[k]) * (cand[k] - x[k]); cn += cand[k] * cand[k]; }
      const double cost_change = x_cost - cand_cost;
      if (sqrt(sn) <= P.ptol * (x_norm + P.ptol)) go = 0;
      else if (fabs(cost_change) <= P.ftol * x_cost) go = 0;
      else {
        const double rel = cand_cost >= 1.7976931348623157e308 ? -1.7976931348623157e308 : cost_change / mcc;
        if (rel > P.min_relative_decrease) {
          for (int k = 0; k < 6; ++k) { x[k] = cand[k]; S.cam[6 * (size_t)c + k] = cand[k]; }
          x_norm = sqrt(cn);
          const double u = 2.0 * rel - 1.0;
          radius = fmin(P.max_radius, radius / fmax(1.0 / 3.0, 1.0 - u * u * u));
          decrease_factor = 2.0; reuse_diag = false;
          go = 2;
        } else {
          radius /= decrease_factor; decrease_factor *= 2.0;
        }
      }
      if (go != 2) cam_derive_record(x, rec);  // the evaluation record follows x
      s_go = go;
    }
    __syncthreads();
    const int go2 = s_go;
    __syncthreads();  // thread 0 rewrites s_go at the top of the next round: everybody has read it by now
    if (go2 == 0) break;
    if (go2 == 2) {
      inner_cam_pass<PD, true>(K, S, O, c, q0, q1, rec, scale, red, tot, prior, prior_kinds);
      if (tot[28] > 0.0) break;  // EvaluateGradientAndJacobian failed: FAILURE, x keeps the accepted point
      if (t == 0) {
        for (int k = 0; k < 21; ++k) H[k] = tot[k];
        for (int k = 0; k < 6; ++k) g[k] = tot[21 + k];
        x_cost = tot[27];
        step_ok = true;
      }
    }
  }
}

// ---- shared intrinsics blocks ----------------------------------------------------------------------------------------------
constexpr int II_VALS = 56;  // 45 (J^T J lower, 9 x 9) + 9 (J^T r) + cost + failures
// One evaluation of the group in slot `slot` over all observations (point-major order): out[II_VALS] += sums.
template <int PD, bool WANT_J>
__global__ void __launch_bounds__(256) k_inner_intr(BaConst K, BaState S, ObsSoA O, const int8_t* __restrict__ o_slot, int slot,
                                                    const double* __restrict__ scale, double* __restrict__ out) {
  __shared__ double red[8][II_VALS];
  double v[II_VALS];
#pragma unroll
  for (int k = 0; k < II_VALS; ++k) v[k] = 0.0;
  double sc[NI];
#pragma unroll
  for (int k = 0; k < NI; ++k) sc[k] = scale[k];
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < K.no; q += gridDim.x * blockDim.x) {
    if (o_slot[q] != slot) continue;
    const int c = O.cam[q], p = O.pt[q];
    const double2 xy = O.xy[q], si = O.si[q];
    if (WANT_J) {
      double r[2], jc[12], jp[2 * PD], ji[2 * NI], hc = 0.0;
      if (!eval_obs<-1, PD, NI, true>(K, S, c, p, xy, si, nullptr, nullptr, nullptr, r, jc, jp, ji, &hc)) { v[55] += 1.0; continue; }
#pragma unroll
      for (int k = 0; k < NI; ++k) { ji[k] *= sc[k]; ji[NI + k] *= sc[k]; }
      int e = 0;
#pragma unroll
      for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) v[e++] += ji[a] * ji[b] + ji[NI + a] * ji[NI + b];
#pragma unroll
      for (int a = 0; a < NI; ++a) v[45 + a] += ji[a] * r[0] + ji[NI + a] * r[1];
      v[54] += hc;
    } else {
      double r[2];
      if (!eval_residual<-1>(K, S, c, p, xy, si, r)) { v[55] += 1.0; continue; }
      const double sq = r[0] * r[0] + r[1] * r[1];
      if (K.loss_type == THB_LOSS_TRIVIAL) v[54] += 0.5 * sq;
      else { double rho[3]; eval_loss(K.loss_type, K.loss_width, sq, rho); v[54] += 0.5 * rho[0]; }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = WANT_J ? 0 : 54; k < II_VALS; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) red[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < II_VALS && (WANT_J || threadIdx.x >= 54)) {
    double s = 0.0;
    for (int ww = 0; ww < 8; ++ww) s += red[ww][threadIdx.x];
    if (s != 0.0) atomicAdd(out + threadIdx.x, s);
  }
}

// |x - cand|^2 and |cand|^2 over the non-constant blocks (TrustRegionMinimizer::ParameterToleranceReached after the inner
// iterations changed the candidate)
__global__ void k_step_norms(int nc, int np, int ng, const uint8_t* __restrict__ cam_const, const uint8_t* __restrict__ pt_const,
                             const int* __restrict__ intr_slot, const int* __restrict__ intr_model, BaState X, BaState Xc,
                             double* __restrict__ step2, double* __restrict__ xnew2) {
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0, n = 0.0;
  auto acc = [&](double a, double b) { s += (a - b) * (a - b); n += b * b; };
  if (i < nc) {
    if (cam_const[i] != THB_CAM_CONST_ALL) for (int k = 0; k < 6; ++k) acc(X.cam[6 * (size_t)i + k], Xc.cam[6 * (size_t)i + k]);
  } else if (i - nc < np) {
    const int p = i - nc;
    if (!pt_const[p]) for (int k = 0; k < 4; ++k) acc(X.pts[(size_t)p * 4 + k], Xc.pts[(size_t)p * 4 + k]);
  } else if (i - nc - np < ng) {
    const int g = i - nc - np;
    if (intr_slot[g] >= 0) for (int k = 0; k < num_intrinsics(intr_model[g]); ++k) acc(X.intr[(size_t)g * KS + k], Xc.intr[(size_t)g * KS + k]);
  }
  s = block_sum(s, red);
  n = block_sum(n, red);
  if (threadIdx.x == 0) { atomicAdd(step2, s); atomicAdd(xnew2, n); }
}

}  // namespace thb
#endif  // THB_INNER_ITER_CUH_
