// Batched theia::BundleAdjustTrack (sfm/bundle_adjustment/bundle_adjustment.cc:261-285): one independent 3-DoF (4-DoF
// without the homogeneous parametrisation) trust-region solve per track, cameras and intrinsics constant
// (bundle_adjuster.cc:176-221: AddTrack adds the observing views with constant extrinsics / intrinsics). The reference
// builds one ceres::Problem per track and calls it once per track from TrackEstimator (estimate_track.cc:289), ~60k calls on
// C5; here one thread runs the whole LevenbergMarquardt loop of its track - the same loop ba_solver.cu::OneIteration runs
// on the host for the joint problem, with the "Schur complement" reduced to the track's own PD x PD block.
// Latency-bound (a track has 2-30 observations, each evaluated ~3 + iterations times); HBM traffic is one read of the
// track's observations per pass.
#ifndef THB_TRACK_BA_CUH_
#define THB_TRACK_BA_CUH_

#include "ba_kernels.cuh"

namespace thb {

struct TrackBaParams {
  int max_num_iterations, max_invalid, jacobi_scaling;
  double ftol, gtol, ptol, radius0, min_radius, max_radius, min_relative_decrease, min_diag, max_diag;
};

// One pass over the track's observations at state St: cost, and (WANT_J) V = sum Jp^T Jp (lower, row-major PD x PD) and
// g = sum Jp^T r with the column scales ps[p * PD + k] applied. Returns false when an evaluation fails (ReprojectionError returned false).
template <int PD, bool WANT_J>
__device__ __noinline__ bool track_pass(const BaConst& K, const BaState& St, const ObsSoA& O, int p, int q0, int q1, const double* ps,
                                        double* cost, double* V, double* g) {
  double c = 0.0;
  if (WANT_J) {
    for (int k = 0; k < PD * PD; ++k) V[k] = 0.0;
    for (int k = 0; k < PD; ++k) g[k] = 0.0;
  }
  bool ok = true;
  for (int q = q0; q < q1; ++q) {
    const int cam = O.cam[q];
    const double2 xy = O.xy[q], si = O.si[q];
    if (WANT_J) {
      double r[2], jc[12], jp[2 * PD], hc = 0.0;
      if (!eval_obs<-1, PD, 0, true>(K, St, cam, p, xy, si, nullptr, ps, nullptr, r, jc, jp, nullptr, &hc)) { ok = false; continue; }
      c += hc;
#pragma unroll
      for (int a = 0; a < PD; ++a) {
        g[a] += jp[a] * r[0] + jp[PD + a] * r[1];
#pragma unroll
        for (int b = 0; b <= a; ++b) V[a * PD + b] += jp[a] * jp[b] + jp[PD + a] * jp[PD + b];
      }
    } else {
      double r[2];
      if (!eval_residual<-1>(K, St, cam, p, xy, si, r)) { ok = false; continue; }
      const double sq = r[0] * r[0] + r[1] * r[1];
      if (K.loss_type == THB_LOSS_TRIVIAL) c += 0.5 * sq;
      else { double rho[3]; eval_loss(K.loss_type, K.loss_width, sq, rho); c += 0.5 * rho[0]; }
    }
  }
  *cost = c;
  return ok;
}

// X.pts: working copy of the points (refined in place), Xc.pts: candidate scratch, ps: [np * PD] column-scale scratch.
// Tracks that are constant or unobserved are left untouched with num_iterations = -1.
template <int PD>
__global__ void __launch_bounds__(128) k_track_ba(BaConst K, BaState X, BaState Xc, ObsSoA O, const int* __restrict__ pt_start, double* ps_all,
                                                  TrackBaParams P, ThbTrackBaResult* __restrict__ res) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= K.np) return;
  const int q0 = pt_start[p], q1 = pt_start[p + 1];
  ThbTrackBaResult out;
  out.initial_cost = 0.0; out.final_cost = 0.0; out.num_iterations = -1; out.termination_type = THB_TERM_CONVERGENCE;
  if (K.pt_const[p] || q1 == q0) { res[p] = out; return; }
  double* xp = X.pts + (size_t)p * 4;
  double* xc = Xc.pts + (size_t)p * 4;
  double* ps = ps_all + (size_t)p * PD;
  double x[4] = {xp[0], xp[1], xp[2], xp[3]};
  double V[PD * PD], g[PD], x_cost = 0.0;
  for (int k = 0; k < PD; ++k) ps[k] = 1.0;
  bool ok = true;
  if (P.jacobi_scaling) {  // column norms of the unscaled Jacobian, fixed for the whole solve
    ok = track_pass<PD, true>(K, X, O, p, q0, q1, ps_all, &x_cost, V, g);
    for (int k = 0; k < PD; ++k) ps[k] = 1.0 / (1.0 + sqrt(V[k * PD + k]));
  }
  ok = track_pass<PD, true>(K, X, O, p, q0, q1, ps_all, &x_cost, V, g) && ok;
  if (!ok) { out.num_iterations = 0; out.termination_type = THB_TERM_FAILURE; res[p] = out; return; }
  out.initial_cost = x_cost;
  double x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  double radius = P.radius0, decrease_factor = 2.0;
  bool step_is_successful = true;
  int iteration = 0, invalid = 0, term = THB_TERM_NO_CONVERGENCE;
  for (;;) {
    if (iteration >= P.max_num_iterations) { term = THB_TERM_NO_CONVERGENCE; break; }
    if (radius <= P.min_radius) { term = THB_TERM_CONVERGENCE; break; }
    const bool want_grad = step_is_successful;
    // LevenbergMarquardtStrategy::ComputeStep: (J^T J + D^2 / radius) y = g, step = -y
    double Vd[PD * PD], Vi[PD * PD], y[PD];
#pragma unroll
    for (int k = 0; k < PD * PD; ++k) Vd[k] = V[k];
#pragma unroll
    for (int a = 0; a < PD; ++a) Vd[a * PD + a] = V[a * PD + a] + fmin(fmax(V[a * PD + a], P.min_diag), P.max_diag) / radius;
    const bool spd = spd_inverse<PD>(Vd, Vi);
#pragma unroll
    for (int a = 0; a < PD; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < PD; ++b) s += Vi[a * PD + b] * g[b];
      y[a] = spd ? s : 0.0;
    }
    // model cost change -(J s)^T (r + J s / 2) with s = -y: y^T g - y^T (J^T J) y / 2
    double yg = 0.0, yVy = 0.0;
#pragma unroll
    for (int a = 0; a < PD; ++a) {
      yg += y[a] * g[a];
#pragma unroll
      for (int b = 0; b < PD; ++b) yVy += y[a] * V[(a >= b ? a * PD + b : b * PD + a)] * y[b];
    }
    const double model_cost_change = yg - 0.5 * yVy;
    // candidate and its cost
    double d[PD], o[4];
#pragma unroll
    for (int k = 0; k < PD; ++k) d[k] = -y[k] * ps[k];
    point_plus<PD>(x, d, o);
    xc[0] = o[0]; xc[1] = o[1]; xc[2] = o[2]; xc[3] = o[3];
    double cand_cost = 0.0;
    const bool cand_ok = track_pass<PD, false>(K, Xc, O, p, q0, q1, ps_all, &cand_cost, nullptr, nullptr);
    if (want_grad) {
      double gmax = 0.0;
#pragma unroll
      for (int k = 0; k < PD; ++k) gmax = fmax(gmax, fabs(g[k] / ps[k]));
      if (P.gtol >= 0.0 && gmax <= P.gtol) { term = THB_TERM_CONVERGENCE; break; }
    }
    ++iteration;
    step_is_successful = false;
    const bool step_valid = spd && isfinite(model_cost_change) && model_cost_change > 0.0;
    if (!step_valid) {
      if (++invalid >= P.max_invalid) { term = THB_TERM_FAILURE; break; }
      radius /= decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    if (!cand_ok) cand_cost = 1.7976931348623157e308;
    double s2 = 0.0, x2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { s2 += (o[k] - x[k]) * (o[k] - x[k]); x2 += o[k] * o[k]; }
    if (P.ptol >= 0.0 && sqrt(s2) <= P.ptol * (x_norm + P.ptol)) { term = THB_TERM_CONVERGENCE; break; }
    const double cost_change = x_cost - cand_cost;
    if (P.ftol >= 0.0 && fabs(cost_change) <= P.ftol * x_cost) { term = THB_TERM_CONVERGENCE; break; }
    const double relative_decrease = cand_cost >= 1.7976931348623157e308 ? -1.7976931348623157e308 : cost_change / model_cost_change;
    if (relative_decrease > P.min_relative_decrease) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { x[k] = o[k]; xp[k] = o[k]; }
      x_norm = sqrt(x2);
      if (!track_pass<PD, true>(K, X, O, p, q0, q1, ps_all, &x_cost, V, g)) { term = THB_TERM_FAILURE; break; }
      step_is_successful = true;
      const double t = 2.0 * relative_decrease - 1.0;
      radius = fmin(P.max_radius, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      decrease_factor = 2.0;
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
    }
  }
  out.final_cost = x_cost; out.num_iterations = iteration; out.termination_type = term;
  res[p] = out;
}

}  // namespace thb
#endif  // THB_TRACK_BA_CUH_
