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
  // TrackEstimator::EstimateTrack stages around the solve (estimate_track.cc:209-321); rays == nullptr: plain BundleAdjustTrack
  const double* rays;       // [no * 3] unit ray direction of every observation, grouped like the observations
  int* status;              // [np] THB_TRACK_*
  int bundle_adjustment;
  double cos_min_angle, sq_max_reprojection_error;
};

// TriangulateMidpoint (triangulation.cc:130-157) over the track's rays, origins = camera positions. As triangulation.cu.
__device__ __forceinline__ bool track_midpoint(const BaState& X, const ObsSoA& O, const double* rays, int q0, int q1, double z[4]) {
  double A[4][4] = {}, b[4] = {};
  for (int q = q0; q < q1; ++q) {
    const double* C = X.cam + (size_t)O.cam[q] * 6;
    const double d[4] = {rays[3 * (size_t)q], rays[3 * (size_t)q + 1], rays[3 * (size_t)q + 2], 0.0};
    const double o[4] = {C[0], C[1], C[2], 1.0};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) { const double T = (r == c ? 1.0 : 0.0) - d[r] * d[c]; A[r][c] += T; s += T * o[c]; }
      b[r] += s;
    }
  }
  bool good = true;
  double L[4][4] = {};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double x = A[k][k];
#pragma unroll
    for (int j = 0; j < k; ++j) x -= L[k][j] * L[k][j];
    if (!(x > 0.0)) good = false;
    x = sqrt(x);
    L[k][k] = x;
#pragma unroll
    for (int r = k + 1; r < 4; ++r) {
      double v = A[r][k];
#pragma unroll
      for (int j = 0; j < k; ++j) v -= L[r][j] * L[k][j];
      L[r][k] = v / x;
    }
  }
  double y[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { double v = b[r]; for (int j = 0; j < r; ++j) v -= L[r][j] * y[j]; y[r] = v / L[r][r]; }
#pragma unroll
  for (int r = 3; r >= 0; --r) { double v = y[r]; for (int j = r + 1; j < 4; ++j) v -= L[j][r] * z[j]; z[r] = v / L[r][r]; }
  return good;
}

// AcceptableReprojectionError (estimate_track.cc:93-119): every view sees the point in front (Camera::ProjectPoint depth
// >= 0, camera.cc:206-216) and the MEAN squared pixel error stays below the threshold.
__device__ __noinline__ bool track_reprojection_ok(const BaConst& K, const BaState& St, const ObsSoA& O, int p, int q0, int q1, double sq_max) {
  const double* Xp = St.pts + (size_t)p * 4;
  const double X[4] = {Xp[0], Xp[1], Xp[2], Xp[3]};
  double sum = 0.0;
  for (int q = q0; q < q1; ++q) {
    const int c = O.cam[q];
    const double* cd = St.camd + (size_t)c * CAMD;
    const double adj[3] = {X[0] - X[3] * cd[CD_C], X[1] - X[3] * cd[CD_C + 1], X[2] - X[3] * cd[CD_C + 2]};
    double pc[3];
    rot_apply(cd + CD_W, cd[CD_A], cd[CD_B], cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2], adj, pc);
    if (pc[2] / X[3] < 0.0) return false;
    double r[2];
    if (!eval_residual<-1>(K, St, c, p, O.xy[q], make_double2(1.0, 1.0), r)) return false;
    sum += r[0] * r[0] + r[1] * r[1];
  }
  return sum / (double)(q1 - q0) < sq_max;
}

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
  if (K.pt_const[p] || (q1 == q0 && !P.rays)) { res[p] = out; if (P.status) P.status[p] = THB_TRACK_SKIPPED; return; }
  double* xp = X.pts + (size_t)p * 4;
  double* xc = Xc.pts + (size_t)p * 4;
  double* ps = ps_all + (size_t)p * PD;
  if (P.rays) {  // EstimateTrack: angle test, then the midpoint as the starting point
    bool wide = false;
    for (int a = q0; a < q1 && !wide; ++a)
      for (int b = a + 1; b < q1; ++b) {
        const double* u = P.rays + 3 * (size_t)a; const double* v = P.rays + 3 * (size_t)b;
        if (u[0] * v[0] + u[1] * v[1] + u[2] * v[2] < P.cos_min_angle) { wide = true; break; }
      }
    if (q1 - q0 < 2 || !wide) { res[p] = out; P.status[p] = THB_TRACK_BAD_ANGLE; return; }
    double z[4];
    if (!track_midpoint(X, O, P.rays, q0, q1, z)) { res[p] = out; P.status[p] = THB_TRACK_FAILED_TRIANGULATION; return; }
    xp[0] = z[0]; xp[1] = z[1]; xp[2] = z[2]; xp[3] = z[3];
    if (!P.bundle_adjustment) {
      out.num_iterations = 0;
      res[p] = out;
      P.status[p] = track_reprojection_ok(K, X, O, p, q0, q1, P.sq_max_reprojection_error) ? THB_TRACK_ESTIMATED : THB_TRACK_BAD_REPROJECTION;
      return;
    }
  }
  double x[4] = {xp[0], xp[1], xp[2], xp[3]};
  double V[PD * PD], g[PD], x_cost = 0.0;
  for (int k = 0; k < PD; ++k) ps[k] = 1.0;
  bool ok = true;
  if (P.jacobi_scaling) {  // column norms of the unscaled Jacobian, fixed for the whole solve
    ok = track_pass<PD, true>(K, X, O, p, q0, q1, ps_all, &x_cost, V, g);
    for (int k = 0; k < PD; ++k) ps[k] = 1.0 / (1.0 + sqrt(V[k * PD + k]));
  }
  ok = track_pass<PD, true>(K, X, O, p, q0, q1, ps_all, &x_cost, V, g) && ok;
  if (!ok) { out.num_iterations = 0; out.termination_type = THB_TERM_FAILURE; res[p] = out; if (P.status) P.status[p] = THB_TRACK_BA_FAILED; return; }
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
  if (P.status) {
    if (term == THB_TERM_FAILURE) P.status[p] = THB_TRACK_BA_FAILED;
    else if (!P.rays) P.status[p] = THB_TRACK_ESTIMATED;
    else P.status[p] = track_reprojection_ok(K, X, O, p, q0, q1, P.sq_max_reprojection_error) ? THB_TRACK_ESTIMATED : THB_TRACK_BAD_REPROJECTION;
  }
}


// SetOutlierTracksToUnestimated (set_outlier_tracks_to_unestimated.cc:62-137), one thread per track; observations reached
// through perm (point-major order of the caller's observation arrays).
__global__ void __launch_bounds__(128) k_outlier_tracks(BaConst K, BaState St, const int* __restrict__ pt_start, const int* __restrict__ perm,
                                                        const int* __restrict__ obs_cam, const double2* __restrict__ obs_xy, double sq_max,
                                                        double cos_min, int* __restrict__ status, int* __restrict__ removed) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= K.np) return;
  if (K.pt_const[p]) { status[p] = THB_OUTLIER_SKIPPED; return; }
  const int q0 = pt_start[p], q1 = pt_start[p + 1];
  const double* Xp = St.pts + (size_t)p * 4;
  const double X[4] = {Xp[0], Xp[1], Xp[2], Xp[3]};
  int st = THB_OUTLIER_KEPT, nproj = 0;
  double sum = 0.0;
  for (int q = q0; q < q1; ++q) {
    const int i = perm[q], c = obs_cam[i];
    const double* cd = St.camd + (size_t)c * CAMD;
    const double adj[3] = {X[0] - X[3] * cd[CD_C], X[1] - X[3] * cd[CD_C + 1], X[2] - X[3] * cd[CD_C + 2]};
    double pc[3];
    rot_apply(cd + CD_W, cd[CD_A], cd[CD_B], cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2], adj, pc);
    if (pc[2] / X[3] < 0.0) { st = THB_OUTLIER_BAD_REPROJECTION; break; }   // depth < 0 (:102-106)
    const int g = K.ng > 1 ? K.cam_group[c] : 0;
    double pix[2] = {0.0, 0.0};
    project<-1, double, double>(K.intr_model[g], St.intr + (size_t)g * KS, pc, pix);
    const double2 xy = obs_xy[i];
    sum += (pix[0] - xy.x) * (pix[0] - xy.x) + (pix[1] - xy.y) * (pix[1] - xy.y);
    ++nproj;
  }
  if (st == THB_OUTLIER_KEPT && sum / (double)nproj > sq_max) st = THB_OUTLIER_BAD_REPROJECTION;  // 0 / 0 = NaN compares false
  if (st == THB_OUTLIER_KEPT) {  // SufficientTriangulationAngle over the rays (point - camera position), triangulation.cc:236-250
    bool wide = false;
    const double h[3] = {X[0] / X[3], X[1] / X[3], X[2] / X[3]};
    for (int a = q0; a < q1 && !wide; ++a) {
      const double* ca = St.camd + (size_t)obs_cam[perm[a]] * CAMD + CD_C;
      double u[3] = {h[0] - ca[0], h[1] - ca[1], h[2] - ca[2]};
      const double nu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
      u[0] /= nu; u[1] /= nu; u[2] /= nu;
      for (int b = a + 1; b < q1; ++b) {
        const double* cb = St.camd + (size_t)obs_cam[perm[b]] * CAMD + CD_C;
        double v[3] = {h[0] - cb[0], h[1] - cb[1], h[2] - cb[2]};
        const double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (u[0] * (v[0] / nv) + u[1] * (v[1] / nv) + u[2] * (v[2] / nv) < cos_min) { wide = true; break; }
      }
    }
    if (!wide) st = THB_OUTLIER_BAD_ANGLE;
  }
  status[p] = st;
  if (st > 0) atomicAdd(removed, 1);
}

// rays in the track-major order of the gathered observations
__global__ void k_gather_rays(int no, const int* __restrict__ perm, const double* __restrict__ raw, double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= no) return;
  const int i = perm[q];
  out[3 * (size_t)q] = raw[3 * (size_t)i]; out[3 * (size_t)q + 1] = raw[3 * (size_t)i + 1]; out[3 * (size_t)q + 2] = raw[3 * (size_t)i + 2];
}

// the caller's points receive the refined track only when the reference would keep it (status == nullptr: BundleAdjustTrack,
// a usable solution; else EstimateTrack: THB_TRACK_ESTIMATED)
__global__ void k_track_commit(int np, const int* __restrict__ status, const ThbTrackBaResult* __restrict__ res, const double* __restrict__ refined,
                               double* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const bool keep = status ? status[p] == THB_TRACK_ESTIMATED : (res[p].num_iterations >= 0 && res[p].termination_type != THB_TERM_FAILURE);
  if (keep) for (int k = 0; k < 4; ++k) out[4 * (size_t)p + k] = refined[4 * (size_t)p + k];
}

// ---- SelectGoodTracksForBundleAdjustment (select_good_tracks_for_bundle_adjustment.cc:263-325) --------------------------
// ComputeStatisticsForTrack (:80-107), one thread per track: truncated length and mean squared reprojection error.
__global__ void __launch_bounds__(128) k_track_stats(BaConst K, BaState St, const int* __restrict__ pt_start, const int* __restrict__ perm,
                                                     const int* __restrict__ obs_cam, const double2* __restrict__ obs_xy, int long_thr,
                                                     int* __restrict__ len, double* __restrict__ mean) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= K.np) return;
  const int q0 = pt_start[p], q1 = pt_start[p + 1];
  const double* Xp = St.pts + (size_t)p * 4;
  const double X[4] = {Xp[0], Xp[1], Xp[2], Xp[3]};
  double sum = 0.0;
  for (int q = q0; q < q1; ++q) {
    const int i = perm[q], c = obs_cam[i];
    const double* cd = St.camd + (size_t)c * CAMD;
    const double adj[3] = {X[0] - X[3] * cd[CD_C], X[1] - X[3] * cd[CD_C + 1], X[2] - X[3] * cd[CD_C + 2]};
    double pc[3];
    rot_apply(cd + CD_W, cd[CD_A], cd[CD_B], cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2], adj, pc);
    const int g = K.ng > 1 ? K.cam_group[c] : 0;
    double pix[2] = {0.0, 0.0};
    project<-1, double, double>(K.intr_model[g], St.intr + (size_t)g * KS, pc, pix);
    const double2 xy = obs_xy[i];
    sum += (pix[0] - xy.x) * (pix[0] - xy.x) + (pix[1] - xy.y) * (pix[1] - xy.y);
  }
  len[p] = min(q1 - q0, long_thr);
  mean[p] = sum / (double)(q1 - q0);
}

// sort keys of the two passes: (view, grid cell) for stage 1, (view, track) for stage 2; views outside the subset sort last
__global__ void k_select_keys(int no, const int* __restrict__ obs_cam, const int* __restrict__ obs_pt, const double2* __restrict__ obs_xy,
                              const uint8_t* __restrict__ cam_sel, double inv_cell, unsigned long long* __restrict__ cell_key,
                              unsigned long long* __restrict__ track_key, int* __restrict__ iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= no) return;
  const int c = obs_cam[i];
  iota[i] = i;
  track_key[i] = ((unsigned long long)(unsigned)c << 32) | (unsigned)obs_pt[i];
  if (cam_sel && !cam_sel[c]) { cell_key[i] = ~0ull; return; }
  const double2 xy = obs_xy[i];
  const int cx = (int)(xy.x * inv_cell), cy = (int)(xy.y * inv_cell);  // Eigen's cast<int>: truncation
  cell_key[i] = ((unsigned long long)(unsigned)c << 42) | ((unsigned long long)((unsigned)cx & 0x1fffffu) << 21) | ((unsigned)cy & 0x1fffffu);
}

// SelectBestTracksFromEachImageGridCell (:152-195): the first observation of every (view, cell) segment walks its segment
__global__ void k_select_cells(int no, const unsigned long long* __restrict__ key, const int* __restrict__ obs, const int* __restrict__ obs_pt,
                               const int* __restrict__ len, const double* __restrict__ mean, uint8_t* __restrict__ selected) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= no) return;
  const unsigned long long k = key[j];
  if (k == ~0ull || (j > 0 && key[j - 1] == k)) return;
  int best = obs_pt[obs[j]];
  for (int q = j + 1; q < no && key[q] == k; ++q) {
    const int t = obs_pt[obs[q]];
    if (len[t] < len[best] || (len[t] == len[best] && mean[t] < mean[best])) best = t;  // std::pair<int, double>::operator<
  }
  selected[best] = 1;
}

// SelectTopRankedTracksInView (:199-254). The views depend on each other through the growing set, so ONE CTA walks them in
// order; inside a view the count and the "first `needed` tracks not yet chosen, ascending track id" are block-parallel.
__global__ void __launch_bounds__(1024) k_select_top_ranked(int nc, const uint8_t* __restrict__ cam_sel, const int* __restrict__ cam_start,
                                                            const unsigned long long* __restrict__ track_key, int min_per_view,
                                                            volatile uint8_t* selected) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  auto block_count = [&](int flag, int* rank) {  // exclusive rank of this thread's flag, returns the block total
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    if (w == 0) {
      const int v = s_warp[lane];
      int inc = v;
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      s_warp[lane] = inc - v;
      if (lane == 31) s_total = inc;
    }
    __syncthreads();
    *rank = s_warp[w] + __popc(m & ((1u << lane) - 1));
    const int tot = s_total;
    __syncthreads();
    return tot;
  };
  for (int c = 0; c < nc; ++c) {
    if (cam_sel && !cam_sel[c]) continue;
    const int q0 = cam_start[c], q1 = cam_start[c + 1], n_est = q1 - q0;
    int n_opt = 0, rank;
    for (int base = q0; base < q1; base += 1024) {
      const int q = base + t;
      n_opt += block_count(q < q1 && selected[(unsigned)(track_key[q < q1 ? q : q0] & 0xffffffffu)] != 0, &rank);
    }
    if (n_opt >= min_per_view || n_opt == n_est) continue;
    const int needed = min(min_per_view - n_opt, n_est - n_opt);
    int taken = 0;
    for (int base = q0; base < q1 && taken < needed; base += 1024) {
      const int q = base + t;
      const unsigned pt = (unsigned)(track_key[q < q1 ? q : q0] & 0xffffffffu);
      const int flag = q < q1 && selected[pt] == 0;
      const int tot = block_count(flag, &rank);
      if (flag && taken + rank < needed) selected[pt] = 1;
      taken += tot;
    }
    __threadfence_block();
    __syncthreads();
  }
}

__global__ void k_count_selected(int np, const uint8_t* __restrict__ selected, int* __restrict__ count) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = p < np && selected[p] != 0;
  const unsigned m = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

}  // namespace thb
#endif  // THB_TRACK_BA_CUH_
