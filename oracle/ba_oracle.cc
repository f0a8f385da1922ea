// ORACLE — test infrastructure only. Nothing under pytheiasfm_b200/ may include, link or
// call this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker and the CPU baseline.
//
// CPU restatement of the reference's bundle-adjustment hot path:
//   theia::BundleAdjuster::Optimize -> ceres::Solve   (sfm/bundle_adjustment/bundle_adjuster.cc:315-355)
// over residual blocks ReprojectionError<Model>        (sfm/camera/reprojection_error.h:49-114)
// with the loss functions of create_loss_function.cc:44-76 / loss_functions.cc:40-44, the
// SubsetManifold / SphereManifold<4> parameterisations of bundle_adjuster.cc:357-460,538-545
// and Schur elimination of the point blocks (groups 0 | 1,2; bundle_adjuster.cc:547-577).
//
// PARITY UNPINNED: Ceres Solver (>= 2.2) and Eigen (>= 3.4) are un-vendored, un-pinned
// third-party dependencies absent from /root/reference and from this image.
// The trust-region policy below restates Ceres' published TrustRegionMinimizer /
// LevenbergMarquardtStrategy / Corrector / SphereManifold; it is pinned
// only by the reference's own property tests (bundle_adjustment_test.cc:76-258) and by an
// independent scipy.optimize.least_squares cross-check (tests/test_oracle_ba.py).
//
// Input/Output use the structs of include/theia_b200.h (host pointers only).

#include <omp.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "../include/theia_b200.h"
#include "camera_models.h"
#include "dense_linalg.h"

namespace oracle {
namespace {

constexpr int KS = THB_INTR_STRIDE;
constexpr double kMaxDouble = std::numeric_limits<double>::max();

// ceres::LossFunction::Evaluate for the losses selected by create_loss_function.cc:44-76.
// Formulas: ceres/loss_function.cc (external; Tukey in the Ceres >= 2.1 scaling).
inline void EvaluateLoss(int type, double a, double s, double rho[3]) {
  const double b = a * a;
  switch (type) {
    case THB_LOSS_HUBER:
      if (s > b) {
        const double r = std::sqrt(s);
        rho[0] = 2.0 * a * r - b;
        rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
        rho[2] = -rho[1] / (2.0 * s);
      } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
      return;
    case THB_LOSS_SOFTLONE: {
      const double c = 1.0 / b, sum = 1.0 + s * c, tmp = std::sqrt(sum);
      rho[0] = 2.0 * b * (tmp - 1.0);
      rho[1] = std::max(std::numeric_limits<double>::min(), 1.0 / tmp);
      rho[2] = -(c * rho[1]) / (2.0 * sum);
      return;
    }
    case THB_LOSS_CAUCHY: {
      const double c = 1.0 / b, sum = 1.0 + s * c, inv = 1.0 / sum;
      rho[0] = b * std::log(sum);
      rho[1] = std::max(std::numeric_limits<double>::min(), inv);
      rho[2] = -c * (inv * inv);
      return;
    }
    case THB_LOSS_ARCTAN: {
      const double bb = 1.0 / b, sum = 1.0 + s * s * bb, inv = 1.0 / sum;
      rho[0] = a * std::atan2(s, a);
      rho[1] = std::max(std::numeric_limits<double>::min(), inv);
      rho[2] = -2.0 * s * bb * (inv * inv);
      return;
    }
    case THB_LOSS_TUKEY:
      if (s <= b) {
        const double value = 1.0 - s / b, value_sq = value * value;
        rho[0] = b / 3.0 * (1.0 - value_sq * value);
        rho[1] = value_sq;
        rho[2] = -2.0 / b * value;
      } else { rho[0] = b / 3.0; rho[1] = 0.0; rho[2] = 0.0; }
      return;
    case THB_LOSS_TRUNCATED:  // loss_functions.cc:40-44
      rho[0] = std::min(s, b); rho[1] = s < b ? 1.0 : 0.0; rho[2] = 0.0;
      return;
    default:
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

// ceres::internal::ComputeHouseholderVector for the 4-vector point (SphereManifold<4>,
// bundle_adjuster.cc:538-545): H = I - beta v v^T, H x = |x| e_4, v[3] = 1.
inline void HouseholderVector4(const double x[4], double v[4], double* beta) {
  const double sigma = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  v[0] = x[0]; v[1] = x[1]; v[2] = x[2]; v[3] = 1.0;
  *beta = 0.0;
  const double xp = x[3];
  if (sigma <= std::numeric_limits<double>::epsilon()) {
    if (xp < 0.0) *beta = 2.0;
    return;
  }
  const double mu = std::sqrt(xp * xp + sigma);
  double vp = 1.0;
  if (xp <= 0.0) vp = xp - mu; else vp = -sigma / (xp + mu);
  *beta = 2.0 * vp * vp / (sigma + vp * vp);
  v[0] /= vp; v[1] /= vp; v[2] /= vp;
}

// SphereManifold<4>::PlusJacobian: |x| * H[:, 0:3]   (row-major 4x3)
inline void SpherePlusJacobian(const double x[4], double J[12]) {
  double v[4], beta;
  HouseholderVector4(x, v, &beta);
  const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  for (int i = 0; i < 3; ++i) {
    for (int r = 0; r < 4; ++r) J[r * 3 + i] = -beta * v[i] * v[r];
    J[i * 3 + i] += 1.0;
  }
  for (int k = 0; k < 12; ++k) J[k] *= nx;
}

// SphereManifold<4>::Plus
inline void SpherePlus(const double x[4], const double d[3], double out[4]) {
  const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd == 0.0) { for (int i = 0; i < 4; ++i) out[i] = x[i]; return; }
  double v[4], beta;
  HouseholderVector4(x, v, &beta);
  const double sbd = std::sin(nd) / nd;
  const double y[4] = {sbd * d[0], sbd * d[1], sbd * d[2], std::cos(nd)};
  const double vty = v[0] * y[0] + v[1] * y[1] + v[2] * y[2] + v[3] * y[3];
  const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  for (int i = 0; i < 4; ++i) out[i] = nx * (y[i] - v[i] * (beta * vty));
}

struct State {
  std::vector<double> cam, intr, pts;
};

class BaOracle {
 public:
  BaOracle(const ThbBaProblem& p, const ThbBaOptions& o) : P(p), O(o) {}

  int Init() {
    nc = P.num_cameras; ng = P.num_groups; np = P.num_points; no = P.num_observations;
    if (nc < 0 || ng < 0 || np < 0 || no < 0) return THB_E_INVALID_ARGUMENT;
    if (P.memory_space != THB_MEM_HOST) return THB_E_INVALID_ARGUMENT;
    if (no > 0 && (!P.cam_ext || !P.cam_group || !P.intr || !P.intr_model || !P.pts || !P.obs_cam ||
                   !P.obs_pt || !P.obs_xy)) return THB_E_INVALID_ARGUMENT;
    for (int g = 0; g < ng; ++g) if (NumIntrinsics(P.intr_model[g]) < 0) return THB_E_UNSUPPORTED;
    for (int c = 0; c < nc; ++c) if (P.cam_group[c] < 0 || P.cam_group[c] >= ng) return THB_E_INVALID_ARGUMENT;
    for (int i = 0; i < no; ++i)
      if (P.obs_cam[i] < 0 || P.obs_cam[i] >= nc || P.obs_pt[i] < 0 || P.obs_pt[i] >= np)
        return THB_E_INVALID_ARGUMENT;
    x.cam.assign(P.cam_ext, P.cam_ext + (size_t)nc * 6);
    x.intr.assign(P.intr, P.intr + (size_t)ng * KS);
    x.pts.assign(P.pts, P.pts + (size_t)np * 4);

    std::vector<char> cam_used(nc, 0), grp_used(ng, 0), pt_used(np, 0);
    for (int i = 0; i < no; ++i) { cam_used[P.obs_cam[i]] = 1; grp_used[P.cam_group[P.obs_cam[i]]] = 1; pt_used[P.obs_pt[i]] = 1; }

    // Tangent structure. SubsetManifold drops constant coordinates (bundle_adjuster.cc:429-441,
    // 478-507); a fully constant block is removed from the program.
    cam_td.assign(nc, 0); cam_idx.assign(nc, {});
    for (int c = 0; c < nc; ++c) {
      const int m = P.cam_const ? P.cam_const[c] : 0;
      int t = 0;
      if (cam_used[c]) {
        if (!(m & THB_CAM_CONST_POSITION)) for (int k = 0; k < 3; ++k) cam_idx[c][t++] = k;
        if (!(m & THB_CAM_CONST_ORIENTATION)) for (int k = 3; k < 6; ++k) cam_idx[c][t++] = k;
      }
      cam_td[c] = t;
    }
    intr_td.assign(ng, 0); intr_idx.assign(ng, {}); intr_K.assign(ng, 0);
    for (int g = 0; g < ng; ++g) {
      const int K = NumIntrinsics(P.intr_model[g]);
      intr_K[g] = K;
      int t = 0;
      if (grp_used[g] && P.intr_const) {
        for (int k = 0; k < K; ++k) if (!((P.intr_const[g] >> k) & 1)) intr_idx[g][t++] = k;
      }
      intr_td[g] = t;
    }
    pt_td.assign(np, 0);
    for (int p = 0; p < np; ++p) {
      const bool c = (P.pt_const && P.pt_const[p]) || !pt_used[p];
      pt_td[p] = c ? 0 : (O.use_homogeneous_point_parametrization ? 3 : 4);
    }
    // Tangent vector layout: [points | intrinsics | cameras] (Schur groups 0 | 1 | 2).
    int off = 0;
    pt_off.assign(np, -1); for (int p = 0; p < np; ++p) if (pt_td[p]) { pt_off[p] = off; off += pt_td[p]; }
    n_pt_tan = off;
    intr_off.assign(ng, -1); for (int g = 0; g < ng; ++g) if (intr_td[g]) { intr_off[g] = off; off += intr_td[g]; }
    cam_off.assign(nc, -1); for (int c = 0; c < nc; ++c) if (cam_td[c]) { cam_off[c] = off; off += cam_td[c]; }
    n_tan = off; n_red = n_tan - n_pt_tan;

    // Residual blocks whose parameter blocks are all constant are dropped by Ceres and their
    // cost goes to Summary::fixed_cost.
    fixed.assign(no, 0);
    for (int i = 0; i < no; ++i) {
      const int c = P.obs_cam[i];
      if (!cam_td[c] && !intr_td[P.cam_group[c]] && !pt_td[P.obs_pt[i]]) fixed[i] = 1;
    }
    // point -> observations
    pt_start.assign(np + 1, 0);
    for (int i = 0; i < no; ++i) pt_start[P.obs_pt[i] + 1]++;
    for (int p = 0; p < np; ++p) pt_start[p + 1] += pt_start[p];
    pt_list.resize(no);
    { std::vector<int> cur(pt_start.begin(), pt_start.end() - 1);
      for (int i = 0; i < no; ++i) pt_list[cur[P.obs_pt[i]]++] = i; }

    // Bounds (bundle_adjuster.cc:396-427) exist on every non-constant intrinsics block.
    is_constrained = false;
    for (int g = 0; g < ng; ++g) if (intr_td[g]) is_constrained = true;

    res.assign((size_t)no * 2, 0.0); Jc.assign((size_t)no * 12, 0.0); Ji.assign((size_t)no * 2 * KS, 0.0);
    Jp.assign((size_t)no * 8, 0.0);
    grad.assign(n_tan, 0.0); scale.assign(n_tan, 1.0);
    return THB_OK;
  }

  // ---- Evaluation of one observation with Jets of width 6+K+4 (as AutoDiffCostFunction
  // <ReprojectionError<Model>, 2, 6, K, 4>, create_reprojection_error_cost_function.h:63-128).
  template <int K>
  bool EvalObsJet(int model, const double* ext, const double* Kp, const double* X, const double obs[2],
                  const double si[2], double r[2], double jc[12], double ji[2 * KS], double jp[8]) const {
    constexpr int N = 6 + K + 4;
    typedef Jet<N> J;
    J e[6], k[K], xx[4], rr[2];
    for (int i = 0; i < 6; ++i) e[i] = J(ext[i], i);
    for (int i = 0; i < K; ++i) k[i] = J(Kp[i], 6 + i);
    for (int i = 0; i < 4; ++i) xx[i] = J(X[i], 6 + K + i);
    const bool ok = ReprojectionError<J>(model, e, k, xx, obs, si, rr);
    if (!ok) return false;
    for (int a = 0; a < 2; ++a) {
      r[a] = rr[a].a;
      for (int i = 0; i < 6; ++i) jc[a * 6 + i] = rr[a].v[i];
      for (int i = 0; i < KS; ++i) ji[a * KS + i] = i < K ? rr[a].v[6 + i] : 0.0;
      for (int i = 0; i < 4; ++i) jp[a * 4 + i] = rr[a].v[6 + K + i];
    }
    return true;
  }

  bool EvalObsAmbient(const State& s, int i, bool want_jac, double r[2], double jc[12], double ji[2 * KS],
                      double jp[8]) const {
    const int c = P.obs_cam[i], g = P.cam_group[c], p = P.obs_pt[i];
    const int model = P.intr_model[g];
    const double* ext = &s.cam[(size_t)c * 6];
    const double* Kp = &s.intr[(size_t)g * KS];
    const double* X = &s.pts[(size_t)p * 4];
    const double obs[2] = {P.obs_xy[2 * (size_t)i], P.obs_xy[2 * (size_t)i + 1]};
    const double si[2] = {P.obs_sqrt_info ? P.obs_sqrt_info[2 * (size_t)i] : 1.0,
                          P.obs_sqrt_info ? P.obs_sqrt_info[2 * (size_t)i + 1] : 1.0};
    if (!want_jac) return ReprojectionError<double>(model, ext, Kp, X, obs, si, r);
    switch (intr_K[g]) {
      case 5: return EvalObsJet<5>(model, ext, Kp, X, obs, si, r, jc, ji, jp);
      case 7: return EvalObsJet<7>(model, ext, Kp, X, obs, si, r, jc, ji, jp);
      case 9: return EvalObsJet<9>(model, ext, Kp, X, obs, si, r, jc, ji, jp);
      default: return false;
    }
  }

  // ProgramEvaluator::Evaluate: cost = sum 0.5*rho(|r|^2); tangent-space Jacobian blocks
  // (ambient * PlusJacobian), Corrector applied; gradient = J^T r (before Jacobi scaling).
  // Writes into res/Jc/Ji/Jp/grad members when want_jac. Returns false if any block fails.
  bool Evaluate(const State& s, bool want_jac, double* cost, bool only_fixed = false) {
    double total = 0.0;
    bool all_ok = true;
    if (want_jac) std::fill(grad.begin(), grad.end(), 0.0);
    const int nth = omp_get_max_threads();
    std::vector<std::vector<double>> gth;
    if (want_jac) gth.assign(nth, std::vector<double>(n_red, 0.0));
#pragma omp parallel for schedule(static) reduction(+ : total) reduction(&& : all_ok)
    for (int p = 0; p < np; ++p) {
      double PJ[12];
      bool have_pj = false;
      double gp[4] = {0, 0, 0, 0};
      std::vector<double>& gt = want_jac ? gth[omp_get_thread_num()] : grad;
      for (int q = pt_start[p]; q < pt_start[p + 1]; ++q) {
        const int i = pt_list[q];
        if ((fixed[i] != 0) != only_fixed) continue;
        double r[2], jc[12], ji[2 * KS], jp[8];
        const bool ok = EvalObsAmbient(s, i, want_jac, r, jc, ji, jp);
        if (!ok) { all_ok = false; continue; }
        const double sq = r[0] * r[0] + r[1] * r[1];
        double rho[3];
        EvaluateLoss(O.loss_function_type, O.robust_loss_width, sq, rho);
        total += 0.5 * rho[0];
        if (!want_jac) continue;
        const int c = P.obs_cam[i], g = P.cam_group[c];
        // tangent blocks
        double tc[12], ti[2 * KS], tp[8];
        const int dc = cam_td[c], di = intr_td[g], dp = pt_td[p];
        for (int a = 0; a < 2; ++a) {
          for (int t = 0; t < dc; ++t) tc[a * 6 + t] = jc[a * 6 + cam_idx[c][t]];
          for (int t = 0; t < di; ++t) ti[a * KS + t] = ji[a * KS + intr_idx[g][t]];
        }
        if (dp == 3) {
          if (!have_pj) { SpherePlusJacobian(&s.pts[(size_t)p * 4], PJ); have_pj = true; }
          for (int a = 0; a < 2; ++a)
            for (int t = 0; t < 3; ++t) {
              double v = 0.0;
              for (int k = 0; k < 4; ++k) v += jp[a * 4 + k] * PJ[k * 3 + t];
              tp[a * 4 + t] = v;
            }
        } else if (dp == 4) {
          for (int k = 0; k < 8; ++k) tp[k] = jp[k];
        }
        // ceres::internal::Corrector (corrector.cc; external)
        const double sqrt_rho1 = std::sqrt(rho[1]);
        double residual_scaling, alpha_sq_norm;
        if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
        else {
          const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
          const double alpha = 1.0 - std::sqrt(D);
          residual_scaling = sqrt_rho1 / (1 - alpha);
          alpha_sq_norm = alpha / sq;
        }
        auto correct = [&](double* blk, int stride, int d) {
          for (int t = 0; t < d; ++t) {
            if (alpha_sq_norm == 0.0) { blk[t] *= sqrt_rho1; blk[stride + t] *= sqrt_rho1; }
            else {
              const double rtj = blk[t] * r[0] + blk[stride + t] * r[1];
              blk[t] = sqrt_rho1 * (blk[t] - alpha_sq_norm * r[0] * rtj);
              blk[stride + t] = sqrt_rho1 * (blk[stride + t] - alpha_sq_norm * r[1] * rtj);
            }
          }
        };
        correct(tc, 6, dc); correct(ti, KS, di); correct(tp, 4, dp);
        r[0] *= residual_scaling; r[1] *= residual_scaling;
        res[2 * (size_t)i] = r[0]; res[2 * (size_t)i + 1] = r[1];
        for (int a = 0; a < 2; ++a) {
          for (int t = 0; t < 6; ++t) Jc[(size_t)i * 12 + a * 6 + t] = t < dc ? tc[a * 6 + t] : 0.0;
          for (int t = 0; t < KS; ++t) Ji[(size_t)i * 2 * KS + a * KS + t] = t < di ? ti[a * KS + t] : 0.0;
          for (int t = 0; t < 4; ++t) Jp[(size_t)i * 8 + a * 4 + t] = t < dp ? tp[a * 4 + t] : 0.0;
        }
        for (int t = 0; t < dp; ++t) gp[t] += tp[t] * r[0] + tp[4 + t] * r[1];
        for (int t = 0; t < dc; ++t) gt[cam_off[c] - n_pt_tan + t] += tc[t] * r[0] + tc[6 + t] * r[1];
        for (int t = 0; t < di; ++t) gt[intr_off[g] - n_pt_tan + t] += ti[t] * r[0] + ti[KS + t] * r[1];
      }
      if (want_jac && pt_td[p]) for (int t = 0; t < pt_td[p]; ++t) grad[pt_off[p] + t] = gp[t];
    }
    if (want_jac)
      for (int t = 0; t < nth; ++t)
        for (int k = 0; k < n_red; ++k) grad[n_pt_tan + k] += gth[t][k];
    if (!only_fixed && AnyPrior()) {
      if (want_jac) { prior_n.assign(nc, 0); prior_r.assign((size_t)nc * kPriorRows, 0.0); prior_J.assign((size_t)nc * kPriorRows * 6, 0.0); }
      for (int c = 0; c < nc; ++c) {
        double r[kPriorRows], Ja[kPriorRows][6];
        const int n = PriorRows(s, c, r, Ja);
        for (int k = 0; k < n; ++k) total += 0.5 * r[k] * r[k];
        if (!want_jac) continue;
        prior_n[c] = n;
        for (int k = 0; k < n; ++k) {
          prior_r[(size_t)c * kPriorRows + k] = r[k];
          for (int t = 0; t < cam_td[c]; ++t) {
            const double j = Ja[k][cam_idx[c][t]];
            prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + t] = j;
            grad[cam_off[c] + t] += j * r[k];
          }
        }
      }
    }
    *cost = total;
    return all_ok;
  }

  void ScaleJacobianColumns() {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < no; ++i) {
      if (fixed[i]) continue;
      const int c = P.obs_cam[i], g = P.cam_group[c], p = P.obs_pt[i];
      for (int a = 0; a < 2; ++a) {
        for (int t = 0; t < cam_td[c]; ++t) Jc[(size_t)i * 12 + a * 6 + t] *= scale[cam_off[c] + t];
        for (int t = 0; t < intr_td[g]; ++t) Ji[(size_t)i * 2 * KS + a * KS + t] *= scale[intr_off[g] + t];
        for (int t = 0; t < pt_td[p]; ++t) Jp[(size_t)i * 8 + a * 4 + t] *= scale[pt_off[p] + t];
      }
    }
  }

  void ScalePriorColumns() {
    if (prior_J.empty()) return;
    for (int c = 0; c < nc; ++c)
      for (int k = 0; k < prior_n[c]; ++k) for (int t = 0; t < cam_td[c]; ++t) prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + t] *= scale[cam_off[c] + t];
  }

  void SquaredColumnNorm(std::vector<double>* out) const {
    out->assign(n_tan, 0.0);
    if (!prior_J.empty())
      for (int c = 0; c < nc; ++c)
        for (int k = 0; k < prior_n[c]; ++k) for (int t = 0; t < cam_td[c]; ++t) { const double v = prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + t]; (*out)[cam_off[c] + t] += v * v; }
    for (int i = 0; i < no; ++i) {
      if (fixed[i]) continue;
      const int c = P.obs_cam[i], g = P.cam_group[c], p = P.obs_pt[i];
      for (int a = 0; a < 2; ++a) {
        for (int t = 0; t < cam_td[c]; ++t) { const double v = Jc[(size_t)i * 12 + a * 6 + t]; (*out)[cam_off[c] + t] += v * v; }
        for (int t = 0; t < intr_td[g]; ++t) { const double v = Ji[(size_t)i * 2 * KS + a * KS + t]; (*out)[intr_off[g] + t] += v * v; }
        for (int t = 0; t < pt_td[p]; ++t) { const double v = Jp[(size_t)i * 8 + a * 4 + t]; (*out)[pt_off[p] + t] += v * v; }
      }
    }
  }

  // EvaluateGradientAndJacobian (trust_region_minimizer.cc; external)
  bool EvaluateGradientAndJacobian(int iteration) {
    if (!Evaluate(x, true, &x_cost)) return false;
    ++n_jac_eval;
    if (O.jacobi_scaling) {
      if (iteration == 0) {
        std::vector<double> cn; SquaredColumnNorm(&cn);
        for (int k = 0; k < n_tan; ++k) scale[k] = 1.0 / (1.0 + std::sqrt(cn[k]));
      }
      ScaleJacobianColumns();
      ScalePriorColumns();
    }
    // unconstrained: max-norm of the gradient; constrained: max |Plus(x,-g) - x|.
    if (!is_constrained) {
      gradient_max_norm = 0.0;
      for (double g : grad) gradient_max_norm = std::max(gradient_max_norm, std::fabs(g));
    } else {
      std::vector<double> ng(grad.size());
      for (size_t k = 0; k < grad.size(); ++k) ng[k] = -grad[k];
      State y; Plus(x, ng, &y);
      gradient_max_norm = MaxAbsDiff(x, y);
    }
    return true;
  }

  double MaxAbsDiff(const State& a, const State& b) const {
    double m = 0.0;
    for (int c = 0; c < nc; ++c) if (cam_td[c]) for (int k = 0; k < 6; ++k) m = std::max(m, std::fabs(a.cam[c * 6 + k] - b.cam[c * 6 + k]));
    for (int g = 0; g < ng; ++g) if (intr_td[g]) for (int k = 0; k < intr_K[g]; ++k) m = std::max(m, std::fabs(a.intr[g * KS + k] - b.intr[g * KS + k]));
    for (int p = 0; p < np; ++p) if (pt_td[p]) for (int k = 0; k < 4; ++k) m = std::max(m, std::fabs(a.pts[(size_t)p * 4 + k] - b.pts[(size_t)p * 4 + k]));
    return m;
  }
  double NormDiff(const State& a, const State* b) const {  // |a - b| (b == nullptr: |a|), non-constant blocks
    double s = 0.0;
    auto acc = [&](double u, double v) { const double d = u - v; s += d * d; };
    for (int c = 0; c < nc; ++c) if (cam_td[c]) for (int k = 0; k < 6; ++k) acc(a.cam[c * 6 + k], b ? b->cam[c * 6 + k] : 0.0);
    for (int g = 0; g < ng; ++g) if (intr_td[g]) for (int k = 0; k < intr_K[g]; ++k) acc(a.intr[g * KS + k], b ? b->intr[g * KS + k] : 0.0);
    for (int p = 0; p < np; ++p) if (pt_td[p]) for (int k = 0; k < 4; ++k) acc(a.pts[(size_t)p * 4 + k], b ? b->pts[(size_t)p * 4 + k] : 0.0);
    return std::sqrt(s);
  }

  // ParameterBlock::Plus: manifold Plus then projection on the box constraints
  // (bounds of bundle_adjuster.cc:396-427).
  void Plus(const State& s, const std::vector<double>& d, State* out) const {
    *out = s;
    for (int c = 0; c < nc; ++c) for (int t = 0; t < cam_td[c]; ++t) out->cam[c * 6 + cam_idx[c][t]] += d[cam_off[c] + t];
    for (int g = 0; g < ng; ++g) {
      if (!intr_td[g]) continue;
      double* K = &out->intr[g * KS];
      for (int t = 0; t < intr_td[g]; ++t) K[intr_idx[g][t]] += d[intr_off[g] + t];
      K[0] = std::max(K[0], 1.0);  // focal length >= 1 (index 0 in every model)
      if (P.intr_model[g] == DOUBLE_SPHERE) {
        K[5] = std::min(std::max(K[5], -1.0), 1.0);
        K[6] = std::min(std::max(K[6], 0.0), 1.0);
      } else if (P.intr_model[g] == EXTENDED_UNIFIED) {
        K[5] = std::min(std::max(K[5], 0.0), 1.0);
        K[6] = std::max(K[6], 0.1);
      }
    }
    for (int p = 0; p < np; ++p) {
      if (pt_td[p] == 3) SpherePlus(&s.pts[(size_t)p * 4], &d[pt_off[p]], &out->pts[(size_t)p * 4]);
      else if (pt_td[p] == 4) for (int k = 0; k < 4; ++k) out->pts[(size_t)p * 4 + k] += d[pt_off[p] + k];
    }
  }

  // Exact solve of (J^T J + D^2) y = J^T r by Schur elimination of the point blocks.
  // Returns false on a non-positive-definite block (LINEAR_SOLVER_FAILURE).
  bool SchurSolve(const std::vector<double>& D, std::vector<double>* y) {
    ++n_solves;
    y->assign(n_tan, 0.0);
    const int nr = n_red;
    std::vector<double> S((size_t)nr * nr, 0.0), rhs(nr, 0.0);
    for (int k = 0; k < nr; ++k) S[(size_t)k * nr + k] = D[n_pt_tan + k] * D[n_pt_tan + k];
    if (!prior_J.empty())
      for (int c = 0; c < nc; ++c) {
        const int o = cam_off[c] - n_pt_tan;
        for (int k = 0; k < prior_n[c]; ++k)
          for (int a = 0; a < cam_td[c]; ++a) {
            const double ja = prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + a];
            rhs[o + a] += ja * prior_r[(size_t)c * kPriorRows + k];
            for (int b = 0; b <= a; ++b) S[(size_t)(o + a) * nr + o + b] += ja * prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + b];
          }
      }
    std::vector<omp_lock_t> locks(std::max(1, nc + ng));
    for (auto& l : locks) omp_init_lock(&l);
    std::vector<double> Vinv((size_t)np * 16, 0.0), gpv((size_t)np * 4, 0.0);
    bool ok = true;
    // reduced-space view of an observation: columns [intr(di) | cam(dc)] with global offsets
#pragma omp parallel for schedule(dynamic, 64)
    for (int p = 0; p < np; ++p) {
      const int n_obs = pt_start[p + 1] - pt_start[p];
      if (!n_obs) continue;
      const int dp = pt_td[p];
      double V[16] = {0}, gp[4] = {0};
      if (dp) {
        for (int t = 0; t < dp; ++t) V[t * dp + t] = D[pt_off[p] + t] * D[pt_off[p] + t];
        for (int q = pt_start[p]; q < pt_start[p + 1]; ++q) {
          const int i = pt_list[q];
          if (fixed[i]) continue;
          const double* jp = &Jp[(size_t)i * 8];
          const double* r = &res[2 * (size_t)i];
          for (int a = 0; a < dp; ++a) {
            gp[a] += jp[a] * r[0] + jp[4 + a] * r[1];
            for (int b = 0; b < dp; ++b) V[a * dp + b] += jp[a] * jp[b] + jp[4 + a] * jp[4 + b];
          }
        }
        // V^-1 by solving for the identity columns
        double Vi[16];
        bool pd = true;
        for (int col = 0; col < dp && pd; ++col) {
          double e[4] = {0, 0, 0, 0}, xcol[4];
          e[col] = 1.0;
          pd = SmallCholeskySolve(dp, V, e, xcol);
          for (int a = 0; a < dp; ++a) Vi[a * dp + col] = xcol[a];
        }
        if (!pd) {
#pragma omp atomic write
          ok = false;
          continue;
        }
        for (int k = 0; k < dp * dp; ++k) Vinv[(size_t)p * 16 + k] = Vi[k];
        for (int a = 0; a < dp; ++a) gpv[(size_t)p * 4 + a] = gp[a];
      }
      if (!nr) continue;
      // per-observation reduced blocks
      struct RB { int dim; int idx[6 + KS]; double Jr[2][6 + KS]; double E[(6 + KS) * 4]; double T[(6 + KS) * 4]; int lock_id; };
      std::vector<RB> rb;
      rb.reserve(n_obs);
      for (int q = pt_start[p]; q < pt_start[p + 1]; ++q) {
        const int i = pt_list[q];
        if (fixed[i]) continue;
        const int c = P.obs_cam[i], g = P.cam_group[c];
        RB b; b.dim = 0; b.lock_id = c;
        for (int t = 0; t < intr_td[g]; ++t) { b.idx[b.dim] = intr_off[g] - n_pt_tan + t; b.Jr[0][b.dim] = Ji[(size_t)i * 2 * KS + t]; b.Jr[1][b.dim] = Ji[(size_t)i * 2 * KS + KS + t]; ++b.dim; }
        for (int t = 0; t < cam_td[c]; ++t) { b.idx[b.dim] = cam_off[c] - n_pt_tan + t; b.Jr[0][b.dim] = Jc[(size_t)i * 12 + t]; b.Jr[1][b.dim] = Jc[(size_t)i * 12 + 6 + t]; ++b.dim; }
        if (!b.dim) continue;
        const double* r = &res[2 * (size_t)i];
        const double* jp = &Jp[(size_t)i * 8];
        // U and g_red contributions
        omp_set_lock(&locks[0]);  // rhs is tiny: one lock
        for (int a = 0; a < b.dim; ++a) rhs[b.idx[a]] += b.Jr[0][a] * r[0] + b.Jr[1][a] * r[1];
        omp_unset_lock(&locks[0]);
        for (int a = 0; a < b.dim; ++a) {
          // rows of S are protected per row owner: use lock of (row mod locks)
          omp_lock_t* lk = &locks[b.idx[a] % locks.size()];
          omp_set_lock(lk);
          for (int bb = 0; bb < b.dim; ++bb)
            if (b.idx[bb] <= b.idx[a]) S[(size_t)b.idx[a] * nr + b.idx[bb]] += b.Jr[0][a] * b.Jr[0][bb] + b.Jr[1][a] * b.Jr[1][bb];
          omp_unset_lock(lk);
        }
        if (dp) {
          const double* Vi = &Vinv[(size_t)p * 16];
          for (int a = 0; a < b.dim; ++a)
            for (int t = 0; t < dp; ++t) b.E[a * 4 + t] = b.Jr[0][a] * jp[t] + b.Jr[1][a] * jp[4 + t];
          for (int a = 0; a < b.dim; ++a)
            for (int t = 0; t < dp; ++t) {
              double v = 0.0;
              for (int u = 0; u < dp; ++u) v += b.E[a * 4 + u] * Vi[u * dp + t];
              b.T[a * 4 + t] = v;
            }
          rb.push_back(b);
        }
      }
      if (!dp) continue;
      for (size_t i1 = 0; i1 < rb.size(); ++i1) {
        const RB& A = rb[i1];
        omp_set_lock(&locks[0]);
        for (int a = 0; a < A.dim; ++a) {
          double v = 0.0;
          for (int t = 0; t < dp; ++t) v += A.T[a * 4 + t] * gp[t];
          rhs[A.idx[a]] -= v;
        }
        omp_unset_lock(&locks[0]);
        for (int a = 0; a < A.dim; ++a) {
          omp_lock_t* lk = &locks[A.idx[a] % locks.size()];
          omp_set_lock(lk);
          double* srow = &S[(size_t)A.idx[a] * nr];
          for (size_t i2 = 0; i2 < rb.size(); ++i2) {
            const RB& B = rb[i2];
            for (int bb = 0; bb < B.dim; ++bb) {
              if (B.idx[bb] > A.idx[a]) continue;
              double v = 0.0;
              for (int t = 0; t < dp; ++t) v += A.T[a * 4 + t] * B.E[bb * 4 + t];
              srow[B.idx[bb]] -= v;
            }
          }
          omp_unset_lock(lk);
        }
      }
    }
    for (auto& l : locks) omp_destroy_lock(&l);
    if (!ok) return false;
    std::vector<double> yr(rhs);
    if (nr) {
      if (O.linear_solver == THB_SOLVER_SCHUR_PCG) {
        if (!ConjugateGradients(S, rhs, &yr)) return false;
      } else {
        if (!CholeskyLower(S.data(), nr)) return false;
        CholeskySolveLower(S.data(), nr, yr.data());
      }
      for (int k = 0; k < nr; ++k) (*y)[n_pt_tan + k] = yr[k];
    }
    // back substitution: y_p = V^-1 (g_p - sum_i E_i^T y_red)
#pragma omp parallel for schedule(static)
    for (int p = 0; p < np; ++p) {
      const int dp = pt_td[p];
      if (!dp) continue;
      double b[4];
      for (int a = 0; a < dp; ++a) b[a] = gpv[(size_t)p * 4 + a];
      for (int q = pt_start[p]; q < pt_start[p + 1]; ++q) {
        const int i = pt_list[q];
        if (fixed[i]) continue;
        const int c = P.obs_cam[i], g = P.cam_group[c];
        const double* jp = &Jp[(size_t)i * 8];
        double jy[2] = {0.0, 0.0};
        for (int t = 0; t < intr_td[g]; ++t) { const double v = yr[intr_off[g] - n_pt_tan + t]; jy[0] += Ji[(size_t)i * 2 * KS + t] * v; jy[1] += Ji[(size_t)i * 2 * KS + KS + t] * v; }
        for (int t = 0; t < cam_td[c]; ++t) { const double v = yr[cam_off[c] - n_pt_tan + t]; jy[0] += Jc[(size_t)i * 12 + t] * v; jy[1] += Jc[(size_t)i * 12 + 6 + t] * v; }
        for (int a = 0; a < dp; ++a) b[a] -= jp[a] * jy[0] + jp[4 + a] * jy[1];
      }
      const double* Vi = &Vinv[(size_t)p * 16];
      for (int a = 0; a < dp; ++a) {
        double v = 0.0;
        for (int u = 0; u < dp; ++u) v += Vi[a * dp + u] * b[u];
        (*y)[pt_off[p] + a] = v;
      }
    }
    for (double v : *y) if (!std::isfinite(v)) return false;
    return true;
  }

  int Covariance(double* cam_cov, uint8_t* cam_ok, double* pt_cov, uint8_t* pt_ok) {
    bool cams_free = false, pts_free = false, intr_free = false;
    for (int c = 0; c < nc; ++c) cams_free |= cam_td[c] > 0;
    for (int q = 0; q < np; ++q) pts_free |= pt_td[q] > 0;
    for (int g = 0; g < ng; ++g) intr_free |= intr_td[g] > 0;
    if (intr_free || (cams_free && pts_free)) return THB_E_UNSUPPORTED;
    if (pts_free && !O.use_homogeneous_point_parametrization) return THB_E_INVALID_ARGUMENT;
    if (cam_ok) std::fill(cam_ok, cam_ok + nc, 0);
    if (pt_ok) std::fill(pt_ok, pt_ok + np, 0);
    if (!cams_free && !pts_free) return THB_OK;
    double cost = 0.0;
    if (!Evaluate(x, true, &cost)) return THB_E_NUMERICAL;
    if (cams_free) {
      std::fill(cam_cov, cam_cov + (size_t)nc * 36, 0.0);
      std::vector<double> N((size_t)nc * 36, 0.0);
      std::vector<int> count(nc, 0);
      for (int i = 0; i < no; ++i) {
        if (fixed[i]) continue;
        const int c = P.obs_cam[i], d = cam_td[c];
        ++count[c];
        for (int a = 0; a < 2; ++a)
          for (int u = 0; u < d; ++u)
            for (int v = 0; v < d; ++v) N[(size_t)c * 36 + u * d + v] += Jc[(size_t)i * 12 + a * 6 + u] * Jc[(size_t)i * 12 + a * 6 + v];
      }
      if (!prior_J.empty())
        for (int c = 0; c < nc; ++c)
          for (int k = 0; k < prior_n[c]; ++k)
            for (int u = 0; u < cam_td[c]; ++u)
              for (int v = 0; v < cam_td[c]; ++v) N[(size_t)c * 36 + u * cam_td[c] + v] += prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + u] * prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + v];
      for (int c = 0; c < nc; ++c) {
        const int d = cam_td[c];
        if (!d || !count[c]) continue;
        double inv[36];
        bool good = true;
        for (int col = 0; col < d && good; ++col) {
          double e[6] = {0, 0, 0, 0, 0, 0}, xcol[6];
          e[col] = 1.0;
          good = DenseCholeskySolve(d, &N[(size_t)c * 36], e, xcol);
          for (int r = 0; r < d; ++r) inv[r * d + col] = xcol[r];
        }
        for (int u = 0; u < d && good; ++u) good = inv[u * d + u] > 0.0 && inv[u * d + u] < 1e28;
        if (!good) continue;
        for (int u = 0; u < d; ++u) for (int v = 0; v < d; ++v) cam_cov[(size_t)c * 36 + cam_idx[c][u] * 6 + cam_idx[c][v]] = inv[u * d + v];
        cam_ok[c] = 1;
      }
    } else {
      std::fill(pt_cov, pt_cov + (size_t)np * 9, 0.0);
      std::vector<double> N((size_t)np * 9, 0.0);
      std::vector<int> count(np, 0);
      for (int i = 0; i < no; ++i) {
        if (fixed[i]) continue;
        const int q = P.obs_pt[i];
        if (pt_td[q] != 3) continue;
        ++count[q];
        for (int a = 0; a < 2; ++a)
          for (int u = 0; u < 3; ++u)
            for (int v = 0; v < 3; ++v) N[(size_t)q * 9 + u * 3 + v] += Jp[(size_t)i * 8 + a * 4 + u] * Jp[(size_t)i * 8 + a * 4 + v];
      }
      for (int q = 0; q < np; ++q) {
        if (pt_td[q] != 3 || !count[q]) continue;
        double inv[9];
        bool good = true;
        for (int col = 0; col < 3 && good; ++col) {
          double e[3] = {0, 0, 0}, xcol[3];
          e[col] = 1.0;
          good = DenseCholeskySolve(3, &N[(size_t)q * 9], e, xcol);
          for (int r = 0; r < 3; ++r) inv[r * 3 + col] = xcol[r];
        }
        for (int u = 0; u < 3 && good; ++u) good = inv[u * 3 + u] > 0.0 && inv[u * 3 + u] < 1e28;
        if (!good) continue;
        for (int k = 0; k < 9; ++k) pt_cov[(size_t)q * 9 + k] = inv[k];
        pt_ok[q] = 1;
      }
    }
    return THB_OK;
  }

  // ceres ITERATIVE_SCHUR with the SCHUR_JACOBI preconditioner on the reduced camera system (iterative_schur_complement_solver.cc,
  // conjugate_gradients_solver.h of Ceres 2.2, restated): preconditioned CG from x = 0 on S y = rhs, S given by its lower triangle.
  // The preconditioner is the inverse of the block diagonal of S, one block per reduced parameter block (camera extrinsics,
  // shared intrinsics), each inverted through its Cholesky factor. Termination: Ceres' Q-test with q_tolerance = eta (the
  // residual test is disabled by the Levenberg-Marquardt strategy, r_tolerance = -1), max_linear_solver_iterations, the residual
  // recomputed from scratch every 10th iteration. Ceres multiplies with the implicit Schur complement; the product with the
  // explicit S is the same operator up to round-off. Returns false on LINEAR_SOLVER_FAILURE; a run that hits the iteration
  // limit or an indefinite direction keeps its current iterate (NO_CONVERGENCE is a usable step for the trust-region loop).
  bool ConjugateGradients(const std::vector<double>& S, const std::vector<double>& b, std::vector<double>* xout) {
    const int nr = n_red;
    struct Blk { int off, dim; };
    std::vector<Blk> blocks;
    for (int g = 0; g < ng; ++g) if (intr_td[g]) blocks.push_back({intr_off[g] - n_pt_tan, intr_td[g]});
    for (int c = 0; c < nc; ++c) if (cam_td[c]) blocks.push_back({cam_off[c] - n_pt_tan, cam_td[c]});
    std::vector<double> Minv(blocks.size() * 81, 0.0);
    bool pd = true;
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < (int)blocks.size(); ++bi) {
      const int o = blocks[bi].off, d = blocks[bi].dim;
      double Bk[81], e[9], col[9];
      for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) Bk[i * d + j] = i >= j ? S[(size_t)(o + i) * nr + o + j] : S[(size_t)(o + j) * nr + o + i];
      for (int j = 0; j < d; ++j) {
        for (int i = 0; i < d; ++i) e[i] = i == j ? 1.0 : 0.0;
        if (!DenseCholeskySolve(d, Bk, e, col)) {
#pragma omp atomic write
          pd = false;
        }
        for (int i = 0; i < d; ++i) Minv[(size_t)bi * 81 + i * d + j] = col[i];
      }
    }
    if (!pd) return false;
    auto multiply = [&](const std::vector<double>& v, std::vector<double>* out) {  // out = S v, symmetric, lower triangle stored
      out->assign(nr, 0.0);
      std::vector<double> upper(nr, 0.0);
#pragma omp parallel
      {
        std::vector<double> loc(nr, 0.0);
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < nr; ++i) {
          const double* row = &S[(size_t)i * nr];
          double acc = 0.0;
          const double vi = v[i];
          for (int j = 0; j < i; ++j) { acc += row[j] * v[j]; loc[j] += row[j] * vi; }
          (*out)[i] = acc + row[i] * vi;
        }
#pragma omp critical
        for (int i = 0; i < nr; ++i) upper[i] += loc[i];
      }
      for (int i = 0; i < nr; ++i) (*out)[i] += upper[i];
    };
    auto dot = [&](const std::vector<double>& u, const std::vector<double>& v) { double a = 0.0; for (int i = 0; i < nr; ++i) a += u[i] * v[i]; return a; };
    auto zero_or_inf = [](double v) { return v == 0.0 || std::isinf(v); };
    std::vector<double>& x = *xout;
    x.assign(nr, 0.0);
    double norm_b = std::sqrt(dot(b, b));
    if (norm_b == 0.0) return true;
    std::vector<double> r(b), z(nr), p(nr), q, tmp(nr);
    double rho = 1.0, Q0 = 0.0;  // Q0 = -x.(b + r) at x = 0
    const int max_it = std::max(1, O.pcg_max_iterations);
    for (int it = 1;; ++it) {
      for (size_t bi = 0; bi < blocks.size(); ++bi) {
        const int o = blocks[bi].off, d = blocks[bi].dim;
        for (int i = 0; i < d; ++i) {
          double a = 0.0;
          for (int j = 0; j < d; ++j) a += Minv[bi * 81 + i * d + j] * r[o + j];
          z[o + i] = a;
        }
      }
      const double last_rho = rho;
      rho = dot(r, z);
      if (zero_or_inf(rho)) return false;
      if (it == 1) p = z;
      else {
        const double beta = rho / last_rho;
        if (zero_or_inf(beta)) return false;
        for (int i = 0; i < nr; ++i) p[i] = z[i] + beta * p[i];
      }
      multiply(p, &q);
      const double pq = dot(p, q);
      if (pq <= 0.0 || std::isinf(pq)) break;  // NO_CONVERGENCE: keep x
      const double alpha = rho / pq;
      if (std::isinf(alpha)) return false;
      for (int i = 0; i < nr; ++i) x[i] += alpha * p[i];
      if (it % 10 == 0) { multiply(x, &tmp); for (int i = 0; i < nr; ++i) r[i] = b[i] - tmp[i]; }
      else for (int i = 0; i < nr; ++i) r[i] -= alpha * q[i];
      double Q1 = 0.0;
      for (int i = 0; i < nr; ++i) Q1 -= x[i] * (b[i] + r[i]);
      const double zeta = it * (Q1 - Q0) / Q1;
      ++n_cg_iterations;
      if (zeta < O.pcg_eta) break;
      Q0 = Q1;
      if (it >= max_it) break;
    }
    return true;
  }

  // model_cost_change = -(J step)^T (r + J step / 2)
  double ModelCostChange(const std::vector<double>& step) const {
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (int i = 0; i < no; ++i) {
      if (fixed[i]) continue;
      const int c = P.obs_cam[i], g = P.cam_group[c], p = P.obs_pt[i];
      double m[2] = {0.0, 0.0};
      for (int a = 0; a < 2; ++a) {
        for (int t = 0; t < cam_td[c]; ++t) m[a] += Jc[(size_t)i * 12 + a * 6 + t] * step[cam_off[c] + t];
        for (int t = 0; t < intr_td[g]; ++t) m[a] += Ji[(size_t)i * 2 * KS + a * KS + t] * step[intr_off[g] + t];
        for (int t = 0; t < pt_td[p]; ++t) m[a] += Jp[(size_t)i * 8 + a * 4 + t] * step[pt_off[p] + t];
      }
      acc += -(m[0] * (res[2 * (size_t)i] + m[0] / 2.0) + m[1] * (res[2 * (size_t)i + 1] + m[1] / 2.0));
    }
    if (!prior_J.empty())
      for (int c = 0; c < nc; ++c)
        for (int k = 0; k < prior_n[c]; ++k) {
          double m = 0.0;
          for (int t = 0; t < cam_td[c]; ++t) m += prior_J[(size_t)c * (kPriorRows * 6) + k * 6 + t] * step[cam_off[c] + t];
          acc += -(m * (prior_r[(size_t)c * kPriorRows + k] + m / 2.0));
        }
    return acc;
  }

  // ---- inner iterations: ceres::internal::CoordinateDescentMinimizer (external) ---------------------
  // Theia reverses the Schur ordering for them (bundle_adjuster.cc:329-334): extrinsics blocks first, then the shared
  // intrinsics blocks, then the points. Every non-constant parameter block is minimised on its own - all other blocks
  // constant at their current values - over the residual blocks that depend on it, by a TrustRegionMinimizer with
  // default Minimizer::Options (50 iterations, tolerances 1e-6 / 1e-10 / 1e-8, Jacobi scaling, no bounds line search),
  // a LevenbergMarquardtStrategy with default TrustRegionStrategy::Options (radius 1e4, maximum 1e32) and DENSE_QR.
  // The blocks of one group share no residual, so solving them in parallel equals solving them one after the other.
  enum { BLK_CAM = 0, BLK_INTR = 1, BLK_PT = 2 };

  void BuildInnerLists() {
    if (!cam_start.empty()) return;
    cam_start.assign(nc + 1, 0); grp_start.assign(ng + 1, 0);
    for (int i = 0; i < no; ++i) { cam_start[P.obs_cam[i] + 1]++; grp_start[P.cam_group[P.obs_cam[i]] + 1]++; }
    for (int c = 0; c < nc; ++c) cam_start[c + 1] += cam_start[c];
    for (int g = 0; g < ng; ++g) grp_start[g + 1] += grp_start[g];
    cam_list.resize(no); grp_list.resize(no);
    std::vector<int> cc(cam_start.begin(), cam_start.end() - 1), gc(grp_start.begin(), grp_start.end() - 1);
    for (int i = 0; i < no; ++i) { cam_list[cc[P.obs_cam[i]]++] = i; grp_list[gc[P.cam_group[P.obs_cam[i]]]++] = i; }
  }

  // Residuals (robustified) and tangent-space Jacobian of the block over its own residual blocks.
  bool EvalBlock(const State& s, int kind, int idx, const int* list, int n, bool want_jac, double* cost,
                 std::vector<double>* r_out, std::vector<double>* J_out, int dim) const {
    double total = 0.0;
    if (want_jac) { r_out->assign((size_t)2 * n, 0.0); J_out->assign((size_t)2 * n * dim, 0.0); }
    double PJ[12];
    if (want_jac && kind == BLK_PT && dim == 3) SpherePlusJacobian(&s.pts[(size_t)idx * 4], PJ);
    for (int q = 0; q < n; ++q) {
      const int i = list[q];
      double r[2], jc[12], ji[2 * KS], jp[8];
      if (!EvalObsAmbient(s, i, want_jac, r, jc, ji, jp)) return false;
      const double sq = r[0] * r[0] + r[1] * r[1];
      double rho[3];
      EvaluateLoss(O.loss_function_type, O.robust_loss_width, sq, rho);
      total += 0.5 * rho[0];
      if (!want_jac) continue;
      double t[2][KS];
      const int c = P.obs_cam[i], g = P.cam_group[c];
      for (int a = 0; a < 2; ++a) {
        if (kind == BLK_CAM) for (int k = 0; k < dim; ++k) t[a][k] = jc[a * 6 + cam_idx[c][k]];
        else if (kind == BLK_INTR) for (int k = 0; k < dim; ++k) t[a][k] = ji[a * KS + intr_idx[g][k]];
        else if (dim == 3) for (int k = 0; k < 3; ++k) { double v = 0.0; for (int u = 0; u < 4; ++u) v += jp[a * 4 + u] * PJ[u * 3 + k]; t[a][k] = v; }
        else for (int k = 0; k < 4; ++k) t[a][k] = jp[a * 4 + k];
      }
      const double sqrt_rho1 = std::sqrt(rho[1]);
      double residual_scaling = sqrt_rho1, alpha_sq_norm = 0.0;
      if (!(sq == 0.0 || rho[2] <= 0.0)) {
        const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / sq;
      }
      for (int k = 0; k < dim; ++k) {
        if (alpha_sq_norm == 0.0) { t[0][k] *= sqrt_rho1; t[1][k] *= sqrt_rho1; }
        else {
          const double rtj = t[0][k] * r[0] + t[1][k] * r[1];
          t[0][k] = sqrt_rho1 * (t[0][k] - alpha_sq_norm * r[0] * rtj);
          t[1][k] = sqrt_rho1 * (t[1][k] - alpha_sq_norm * r[1] * rtj);
        }
      }
      (*r_out)[2 * q] = r[0] * residual_scaling; (*r_out)[2 * q + 1] = r[1] * residual_scaling;
      for (int a = 0; a < 2; ++a) for (int k = 0; k < dim; ++k) (*J_out)[((size_t)2 * q + a) * dim + k] = t[a][k];
    }
    if (kind == BLK_CAM && AnyPrior()) {  // the camera's prior blocks are residual blocks of this parameter block as well
      double r[kPriorRows], Ja[kPriorRows][6];
      const int n = PriorRows(s, idx, r, Ja);
      for (int k = 0; k < n; ++k) {
        total += 0.5 * r[k] * r[k];
        if (want_jac) {
          r_out->push_back(r[k]);
          for (int t = 0; t < dim; ++t) J_out->push_back(Ja[k][cam_idx[idx][t]]);
        }
      }
    }
    *cost = total;
    return true;
  }

  // ParameterBlock::Plus of one block (manifold Plus, then the box projection), in place on a copy of the block.
  void BlockPlus(int kind, int idx, const double* xb, const double* d, int dim, double* out) const {
    if (kind == BLK_CAM) {
      for (int k = 0; k < 6; ++k) out[k] = xb[k];
      for (int k = 0; k < dim; ++k) out[cam_idx[idx][k]] += d[k];
    } else if (kind == BLK_INTR) {
      for (int k = 0; k < KS; ++k) out[k] = xb[k];
      for (int k = 0; k < dim; ++k) out[intr_idx[idx][k]] += d[k];
      out[0] = std::max(out[0], 1.0);
      if (P.intr_model[idx] == DOUBLE_SPHERE) { out[5] = std::min(std::max(out[5], -1.0), 1.0); out[6] = std::min(std::max(out[6], 0.0), 1.0); }
      else if (P.intr_model[idx] == EXTENDED_UNIFIED) { out[5] = std::min(std::max(out[5], 0.0), 1.0); out[6] = std::max(out[6], 0.1); }
    } else if (dim == 3) {
      SpherePlus(xb, d, out);
    } else {
      for (int k = 0; k < 4; ++k) out[k] = xb[k] + d[k];
    }
  }

  // TrustRegionMinimizer::Minimize on one parameter block; s is updated in place on success.
  void InnerSolve(State* s, int kind, int idx) const {
    const int* list; int n, dim, bs; double* xb;
    if (kind == BLK_CAM) { list = &cam_list[cam_start[idx]]; n = cam_start[idx + 1] - cam_start[idx]; dim = cam_td[idx]; bs = 6; xb = &s->cam[(size_t)idx * 6]; }
    else if (kind == BLK_INTR) { list = &grp_list[grp_start[idx]]; n = grp_start[idx + 1] - grp_start[idx]; dim = intr_td[idx]; bs = intr_K[idx]; xb = &s->intr[(size_t)idx * KS]; }
    else { list = &pt_list[pt_start[idx]]; n = pt_start[idx + 1] - pt_start[idx]; dim = pt_td[idx]; bs = 4; xb = &s->pts[(size_t)idx * 4]; }
    if (dim == 0 || n == 0) return;
    const int store = kind == BLK_INTR ? KS : bs;
    double x0[KS], cand[KS];
    for (int k = 0; k < store; ++k) x0[k] = xb[k];  // the minimiser's x; xb follows the best point
    auto norm_of = [&](const double* v) { double q = 0.0; for (int k = 0; k < bs; ++k) q += v[k] * v[k]; return std::sqrt(q); };
    std::vector<double> r, J;
    double x_cost = 0.0, scale[KS], g[KS], diag[KS];
    auto evaluate_gj = [&](bool first) {
      if (!EvalBlock(*s, kind, idx, list, n, true, &x_cost, &r, &J, dim)) return false;
      for (int k = 0; k < dim; ++k) { double a = 0.0; for (int q = 0; q < (int)r.size(); ++q) a += J[(size_t)q * dim + k] * r[q]; g[k] = a; }
      if (first) for (int k = 0; k < dim; ++k) { double a = 0.0; for (int q = 0; q < (int)r.size(); ++q) a += J[(size_t)q * dim + k] * J[(size_t)q * dim + k]; scale[k] = 1.0 / (1.0 + std::sqrt(a)); }
      for (int q = 0; q < (int)r.size(); ++q) for (int k = 0; k < dim; ++k) J[(size_t)q * dim + k] *= scale[k];
      return true;
    };
    if (!evaluate_gj(true)) return;  // IterationZero fails: FAILURE, parameters untouched
    double x_norm = norm_of(x0), radius = 1e4, decrease_factor = 2.0, gmax = 0.0;
    bool reuse_diagonal = false, step_ok = true;
    int iteration = 0, invalid = 0;
    for (;;) {
      if (iteration >= 50) break;
      if (step_ok) { gmax = 0.0; for (int k = 0; k < dim; ++k) gmax = std::max(gmax, std::fabs(g[k])); if (gmax <= 1e-10) break; }
      if (radius <= 1e-32) break;
      ++iteration;
      step_ok = false;
      if (!reuse_diagonal)
        for (int k = 0; k < dim; ++k) { double a = 0.0; for (int q = 0; q < (int)r.size(); ++q) a += J[(size_t)q * dim + k] * J[(size_t)q * dim + k]; diag[k] = std::min(std::max(a, 1e-6), 1e32); }
      reuse_diagonal = true;
      // (J^T J + D^2) y = J^T r with D^2 = diag / radius; step = -y (DENSE_QR on [J; D] in Ceres: the same least squares)
      double M[KS * KS], b[KS], y[KS];
      for (int a = 0; a < dim; ++a) {
        double bb = 0.0; for (int q = 0; q < (int)r.size(); ++q) bb += J[(size_t)q * dim + a] * r[q]; b[a] = bb;
        for (int c = 0; c <= a; ++c) { double v = 0.0; for (int q = 0; q < (int)r.size(); ++q) v += J[(size_t)q * dim + a] * J[(size_t)q * dim + c]; M[a * dim + c] = v; M[c * dim + a] = v; }
        M[a * dim + a] += diag[a] / radius;
      }
      bool valid = DenseCholeskySolve(dim, M, b, y);
      double step[KS], mcc = 0.0;
      if (valid) {
        for (int k = 0; k < dim; ++k) step[k] = -y[k];
        for (int q = 0; q < (int)r.size(); ++q) { double m = 0.0; for (int k = 0; k < dim; ++k) m += J[(size_t)q * dim + k] * step[k]; mcc += -(m * (r[q] + m / 2.0)); }
        valid = std::isfinite(mcc) && mcc > 0.0;
      }
      if (!valid) {
        if (++invalid >= 5) break;
        radius /= decrease_factor; decrease_factor *= 2.0;
        continue;
      }
      invalid = 0;
      double delta[KS];
      for (int k = 0; k < dim; ++k) delta[k] = step[k] * scale[k];
      BlockPlus(kind, idx, x0, delta, dim, cand);
      for (int k = 0; k < store; ++k) xb[k] = cand[k];  // evaluate the candidate in place
      double cand_cost;
      std::vector<double> dummy_r, dummy_J;
      if (!EvalBlock(*s, kind, idx, list, n, false, &cand_cost, &dummy_r, &dummy_J, dim)) cand_cost = kMaxDouble;
      for (int k = 0; k < store; ++k) xb[k] = x0[k];
      double sn = 0.0; for (int k = 0; k < bs; ++k) sn += (cand[k] - x0[k]) * (cand[k] - x0[k]);
      if (std::sqrt(sn) <= 1e-8 * (x_norm + 1e-8)) break;
      const double cost_change = x_cost - cand_cost;
      if (std::fabs(cost_change) <= 1e-6 * x_cost) break;
      const double rel = cand_cost >= kMaxDouble ? std::numeric_limits<double>::lowest() : cost_change / mcc;
      if (rel > 1e-3) {
        for (int k = 0; k < store; ++k) { x0[k] = cand[k]; xb[k] = cand[k]; }
        x_norm = norm_of(x0);
        if (!evaluate_gj(false)) break;
        step_ok = true;
        radius = std::min(1e32, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3)));
        decrease_factor = 2.0; reuse_diagonal = false;
      } else {
        radius /= decrease_factor; decrease_factor *= 2.0;
      }
    }
    for (int k = 0; k < store; ++k) xb[k] = x0[k];
  }

  static bool DenseCholeskySolve(int n, const double* A, const double* b, double* x) {
    double L[KS * KS];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        double v = A[i * n + j];
        for (int k = 0; k < j; ++k) v -= L[i * n + k] * L[j * n + k];
        if (i == j) { if (!(v > 0.0) || !std::isfinite(v)) return false; L[i * n + i] = std::sqrt(v); }
        else L[i * n + j] = v / L[j * n + j];
      }
    double y[KS];
    for (int i = 0; i < n; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[i * n + k] * y[k]; y[i] = v / L[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < n; ++k) v -= L[k * n + i] * x[k]; x[i] = v / L[i * n + i]; }
    for (int i = 0; i < n; ++i) if (!std::isfinite(x[i])) return false;
    return true;
  }

  void CoordinateDescent(State* s) {
    BuildInnerLists();
#pragma omp parallel for schedule(dynamic, 4)
    for (int c = 0; c < nc; ++c) InnerSolve(s, BLK_CAM, c);
    for (int g = 0; g < ng; ++g) InnerSolve(s, BLK_INTR, g);
#pragma omp parallel for schedule(dynamic, 64)
    for (int p = 0; p < np; ++p) InnerSolve(s, BLK_PT, p);
  }

  void Log(ThbBaSummary* s, double cost, double radius) {
    if (s->iter_log_count < THB_MAX_ITER_LOG) {
      s->iter_cost[s->iter_log_count] = cost;
      s->iter_radius[s->iter_log_count] = radius;
      ++s->iter_log_count;
    }
  }

  // TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy (external).
  int Solve(ThbBaSummary* sum) {
    const auto t0 = std::chrono::steady_clock::now();
    std::memset(sum, 0, sizeof(*sum));
    double fixed_cost = 0.0;
    bool any_fixed = false;
    for (int i = 0; i < no; ++i) any_fixed |= fixed[i] != 0;
    if (any_fixed && !Evaluate(x, false, &fixed_cost, true)) { sum->termination_type = THB_TERM_FAILURE; return THB_E_NUMERICAL; }
    double radius = O.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    std::vector<double> diagonal(n_tan, 0.0), lm_diag(n_tan, 0.0), step, delta(n_tan);
    int num_consecutive_invalid = 0;
    sum->termination_type = THB_TERM_NO_CONVERGENCE;
    // IterationZero
    if (is_constrained) { std::vector<double> z(n_tan, 0.0); State y; Plus(x, z, &y); x = y; }
    double x_norm = NormDiff(x, nullptr);
    if (n_tan == 0) {  // nothing to optimise: Ceres returns CONVERGENCE with cost = fixed cost
      sum->initial_cost = sum->final_cost = fixed_cost; sum->success = 1; sum->termination_type = THB_TERM_CONVERGENCE;
      return THB_OK;
    }
    if (!EvaluateGradientAndJacobian(0)) { sum->termination_type = THB_TERM_FAILURE; sum->success = 0; return THB_E_NUMERICAL; }
    sum->initial_cost = x_cost + fixed_cost;
    double min_cost = x_cost;
    Log(sum, x_cost + fixed_cost, radius);
    bool step_is_successful = true;  // iteration 0 counts as successful for the gradient test
    int iteration = 0;
    // the preprocessor disables inner iterations on programs with fewer than two parameter blocks
    int num_blocks = 0;
    for (int c = 0; c < nc; ++c) num_blocks += cam_td[c] > 0;
    for (int g = 0; g < ng; ++g) num_blocks += intr_td[g] > 0;
    for (int p = 0; p < np; ++p) num_blocks += pt_td[p] > 0;
    bool inner_enabled = O.use_inner_iterations != 0 && num_blocks >= 2;
    const auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    while (true) {
      // FinalizeIterationAndCheckIfMinimizerCanContinue
      if (elapsed() >= O.max_solver_time_in_seconds) { sum->termination_type = THB_TERM_NO_CONVERGENCE; break; }
      if (iteration >= O.max_num_iterations) { sum->termination_type = THB_TERM_NO_CONVERGENCE; break; }
      if (O.gradient_tolerance >= 0.0 && step_is_successful && gradient_max_norm <= O.gradient_tolerance) { sum->termination_type = THB_TERM_CONVERGENCE; break; }
      if (radius <= O.min_trust_region_radius) { sum->termination_type = THB_TERM_CONVERGENCE; break; }
      ++iteration;
      step_is_successful = false;
      // ComputeTrustRegionStep / LevenbergMarquardtStrategy::ComputeStep
      if (!reuse_diagonal) {
        SquaredColumnNorm(&diagonal);
        for (double& d : diagonal) d = std::min(std::max(d, O.min_lm_diagonal), O.max_lm_diagonal);
      }
      for (int k = 0; k < n_tan; ++k) lm_diag[k] = std::sqrt(diagonal[k] / radius);
      bool step_valid = SchurSolve(lm_diag, &step);
      reuse_diagonal = true;
      double model_cost_change = 0.0;
      if (step_valid) {
        for (double& v : step) v = -v;
        model_cost_change = ModelCostChange(step);
        step_valid = model_cost_change > 0.0;
      }
      if (!step_valid) {
        // HandleInvalidStep
        if (++num_consecutive_invalid >= O.max_num_consecutive_invalid_steps) { sum->termination_type = THB_TERM_FAILURE; break; }
        radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        Log(sum, x_cost + fixed_cost, radius);
        continue;
      }
      num_consecutive_invalid = 0;
      for (int k = 0; k < n_tan; ++k) delta[k] = step[k] * scale[k];
      State cand;
      if (is_constrained) {
        // DoLineSearch: Armijo, first probe at step size 1. PARTIAL RESTATEMENT: Ceres
        // contracts the step by polynomial interpolation when the probe fails; here a failed
        // probe halves delta until the sufficient-decrease test holds (max 20 probes).
        double gtd = 0.0;
        for (int k = 0; k < n_tan; ++k) gtd += grad[k] * delta[k];
        double alpha = 1.0;
        for (int it = 0; it < 20; ++it) {
          std::vector<double> sd(delta);
          for (double& v : sd) v *= alpha;
          Plus(x, sd, &cand);
          double c;
          ++n_cost_eval;
          if (Evaluate(cand, false, &c) && std::isfinite(c) && c <= x_cost + 1e-4 * gtd * alpha) break;
          alpha *= 0.5;
          if (it == 19) alpha = 1.0;  // line search failed: delta unchanged
        }
        for (double& v : delta) v *= alpha;
      }
      // ComputeCandidatePointAndEvaluateCost
      Plus(x, delta, &cand);
      double cand_cost;
      ++n_cost_eval;
      if (!Evaluate(cand, false, &cand_cost)) cand_cost = kMaxDouble;
      // DoInnerIterationsIfNeeded (trust_region_minimizer.cc; external)
      bool inner_useful = false;
      if (inner_enabled && cand_cost < kMaxDouble) {
        State inner = cand;
        CoordinateDescent(&inner);
        double inner_cost;
        ++n_cost_eval;
        if (Evaluate(inner, false, &inner_cost)) {
          cand = inner;
          model_cost_change += cand_cost - inner_cost;
          inner_useful = inner_cost < x_cost;
          inner_enabled = 1.0 - inner_cost / cand_cost > 1e-3;  // inner_iteration_tolerance
          cand_cost = inner_cost;
        }
      }
      // ParameterToleranceReached
      const double step_norm = NormDiff(x, &cand);
      if (O.parameter_tolerance >= 0.0 && step_norm <= O.parameter_tolerance * (x_norm + O.parameter_tolerance)) { sum->termination_type = THB_TERM_CONVERGENCE; break; }
      // FunctionToleranceReached
      const double cost_change = x_cost - cand_cost;
      if (O.function_tolerance >= 0.0 && std::fabs(cost_change) <= O.function_tolerance * x_cost) { sum->termination_type = THB_TERM_CONVERGENCE; break; }
      // IsStepSuccessful (monotonic steps)
      const double relative_decrease = cand_cost >= kMaxDouble ? std::numeric_limits<double>::lowest() : cost_change / model_cost_change;
      if (inner_useful || relative_decrease > O.min_relative_decrease) {
        x = cand; x_norm = NormDiff(x, nullptr);
        if (!EvaluateGradientAndJacobian(iteration)) { sum->termination_type = THB_TERM_FAILURE; break; }
        step_is_successful = true;
        ++sum->num_successful_steps;
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
        radius = std::min(O.max_trust_region_radius, radius);
        decrease_factor = 2.0; reuse_diagonal = false;
        min_cost = std::min(min_cost, x_cost);
      } else {
        radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      }
      Log(sum, x_cost + fixed_cost, radius);
    }
    sum->num_iterations = iteration;
    sum->final_cost = min_cost + fixed_cost;
    sum->success = sum->termination_type != THB_TERM_FAILURE;
    sum->num_jacobian_evaluations = n_jac_eval; sum->num_cost_evaluations = n_cost_eval; sum->num_linear_solves = n_solves; sum->num_linear_solver_iterations = n_cg_iterations;
    sum->solve_time_in_seconds = elapsed();
    if (sum->success) {
      std::memcpy(P.cam_ext, x.cam.data(), sizeof(double) * x.cam.size());
      std::memcpy(P.intr, x.intr.data(), sizeof(double) * x.intr.size());
      std::memcpy(P.pts, x.pts.data(), sizeof(double) * x.pts.size());
    }
    return THB_OK;
  }

  ThbBaProblem P; ThbBaOptions O;
  int nc = 0, ng = 0, np = 0, no = 0, n_tan = 0, n_pt_tan = 0, n_red = 0;
  State x;
  std::vector<int> cam_td, intr_td, intr_K, pt_td, pt_off, intr_off, cam_off, pt_start, pt_list;
  std::vector<int> cam_start, cam_list, grp_start, grp_list;
  std::vector<std::array<int, 6>> cam_idx;
  std::vector<std::array<int, KS>> intr_idx;
  std::vector<char> fixed;
  bool is_constrained = false;
  std::vector<double> res, Jc, Ji, Jp, grad, scale;
  double x_cost = 0.0, gradient_max_norm = 0.0;
  int n_jac_eval = 0, n_cost_eval = 0, n_solves = 0, n_cg_iterations = 0;
  // Camera prior residual blocks, no loss function: position (bundle_adjuster.cc:160-163, position_error.h:44-80)
  // sqrt_info * (prior - position), gravity (:165-168, gravity_error.h:44-86) sqrt_info * (R(aa) (0,0,-1) - prior) and orientation (:170-172,
  // orientation_error.h:44-80) sqrt_info * log(exp(aa) exp(prior)^-1), differentiated
  // with Jets like ceres::AutoDiffCostFunction does. prior_n[c] rows (0, 3, 6 or 9), prior_r their residuals, prior_J the rows x cam_td
  // tangent Jacobian (stride 6) as of the last Evaluate with want_jac.
  static constexpr int kPriorRows = 9;  // position, gravity, orientation: 3 rows each
  std::vector<int> prior_n;
  std::vector<double> prior_r, prior_J;
  bool HasPositionPrior(int c) const { return P.cam_has_position_prior && P.cam_position_prior && P.cam_position_prior_sqrt_info && P.cam_has_position_prior[c] && cam_td[c] > 0; }
  bool HasGravityPrior(int c) const { return P.cam_has_gravity_prior && P.cam_gravity_prior && P.cam_gravity_prior_sqrt_info && P.cam_has_gravity_prior[c] && cam_td[c] > 0; }
  bool HasOrientationPrior(int c) const { return P.cam_has_orientation_prior && P.cam_orientation_prior && P.cam_orientation_prior_sqrt_info && P.cam_has_orientation_prior[c] && cam_td[c] > 0; }
  bool AnyPrior() const { return P.cam_has_position_prior || P.cam_has_gravity_prior || P.cam_has_orientation_prior; }
  // residuals r[<= 9] and ambient Jacobian Ja[row][6] of camera c's prior blocks at state s; returns the number of rows
  int PriorRows(const State& s, int c, double* r, double (*Ja)[6]) const {
    int n = 0;
    if (HasPositionPrior(c)) {
      const double* A = P.cam_position_prior_sqrt_info + 9 * (size_t)c;
      const double* pr = P.cam_position_prior + 3 * (size_t)c;
      const double d[3] = {pr[0] - s.cam[(size_t)c * 6], pr[1] - s.cam[(size_t)c * 6 + 1], pr[2] - s.cam[(size_t)c * 6 + 2]};
      for (int k = 0; k < 3; ++k, ++n) {
        r[n] = A[3 * k] * d[0] + A[3 * k + 1] * d[1] + A[3 * k + 2] * d[2];
        for (int a = 0; a < 6; ++a) Ja[n][a] = a < 3 ? -A[3 * k + a] : 0.0;
      }
    }
    if (HasGravityPrior(c)) {
      const double* A = P.cam_gravity_prior_sqrt_info + 9 * (size_t)c;
      const double* gp = P.cam_gravity_prior + 3 * (size_t)c;
      typedef Jet<3> J3;
      const J3 aa[3] = {J3(s.cam[(size_t)c * 6 + 3], 0), J3(s.cam[(size_t)c * 6 + 4], 1), J3(s.cam[(size_t)c * 6 + 5], 2)};
      const J3 gw[3] = {J3(0.0), J3(0.0), J3(-1.0)};
      J3 gc[3];
      AngleAxisRotatePoint(aa, gw, gc);
      for (int k = 0; k < 3; ++k, ++n) {
        J3 res = J3(A[3 * k]) * (gc[0] - gp[0]) + J3(A[3 * k + 1]) * (gc[1] - gp[1]) + J3(A[3 * k + 2]) * (gc[2] - gp[2]);
        r[n] = res.a;
        for (int a = 0; a < 6; ++a) Ja[n][a] = a < 3 ? 0.0 : res.v[a - 3];
      }
    }
    if (HasOrientationPrior(c)) {  // bundle_adjuster.cc:170-172, orientation_error.h:44-80
      const double* A = P.cam_orientation_prior_sqrt_info + 9 * (size_t)c;
      const double* op = P.cam_orientation_prior + 3 * (size_t)c;
      typedef Jet<3> J3;
      const J3 aa[3] = {J3(s.cam[(size_t)c * 6 + 3], 0), J3(s.cam[(size_t)c * 6 + 4], 1), J3(s.cam[(size_t)c * 6 + 5], 2)};
      const J3 pa[3] = {J3(op[0]), J3(op[1]), J3(op[2])};
      J3 qa[4], qb[4], q[4], e[3];
      So3Exp(aa, qa);
      So3Exp(pa, qb);
      qb[1] = -qb[1]; qb[2] = -qb[2]; qb[3] = -qb[3];
      So3Mul(qa, qb, q);
      So3Log(q, e);
      for (int k = 0; k < 3; ++k, ++n) {
        J3 res = J3(A[3 * k]) * e[0] + J3(A[3 * k + 1]) * e[1] + J3(A[3 * k + 2]) * e[2];
        r[n] = res.a;
        for (int a = 0; a < 6; ++a) Ja[n][a] = a < 3 ? 0.0 : res.v[a - 3];
      }
    }
    return n;
  }
  // Sophus::SO3 (so3.hpp, a third-party dependency absent from /root/reference: restated from its published source, PARITY
  // UNPINNED): exp to a unit quaternion (w, x, y, z) with the Taylor branch below |omega|^2 < 1e-20, the group product with its
  // first-order renormalisation, log with the atan2 branch on the sign of w.
  template <typename T> static void So3Exp(const T w[3], T q[4]) {
    const T th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    T imag, real;
    if (th2 < 1e-10 * 1e-10) {
      const T th4 = th2 * th2;
      imag = T(0.5) - T(1.0 / 48.0) * th2 + T(1.0 / 3840.0) * th4;
      real = T(1.0) - T(1.0 / 8.0) * th2 + T(1.0 / 384.0) * th4;
    } else {
      const T th = sqrt_(th2), half = T(0.5) * th;
      imag = sin_(half) / th;
      real = cos_(half);
    }
    q[0] = real; q[1] = imag * w[0]; q[2] = imag * w[1]; q[3] = imag * w[2];
  }
  template <typename T> static void So3Mul(const T a[4], const T b[4], T q[4]) {
    q[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    q[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    q[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    q[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    const T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n2 != 1.0) { const T sc = T(2.0) / (T(1.0) + n2); for (int k = 0; k < 4; ++k) q[k] = q[k] * sc; }
  }
  template <typename T> static void So3Log(const T q[4], T e[3]) {
    const T n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    T f;
    if (n2 < 1e-10 * 1e-10) {
      f = T(2.0) / q[0] - T(2.0 / 3.0) * n2 / (q[0] * q[0] * q[0]);
    } else {
      const T n = sqrt_(n2);
      const T at = q[0] < 0.0 ? atan2_(-n, -q[0]) : atan2_(n, q[0]);
      f = T(2.0) * at / n;
    }
    for (int k = 0; k < 3; ++k) e[k] = f * q[1 + k];
  }
};

}  // namespace
}  // namespace oracle

extern "C" {

void oracle_ba_default_options(ThbBaOptions* o) {
  std::memset(o, 0, sizeof(*o));
  // bundle_adjustment.h:87-167 defaults, except use_inner_iterations
  o->loss_function_type = THB_LOSS_TRIVIAL; o->robust_loss_width = 2.0;
  o->linear_solver = THB_SOLVER_SCHUR_CHOLESKY;
  o->use_homogeneous_point_parametrization = 1; o->use_inner_iterations = 0;
  o->max_num_iterations = 100; o->jacobi_scaling = 1; o->verbose = 0;
  o->max_num_consecutive_invalid_steps = 5;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->max_trust_region_radius = 1e12; o->initial_trust_region_radius = 1e4;
  o->min_trust_region_radius = 1e-32; o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->max_solver_time_in_seconds = 3600.0;
  o->pcg_eta = 0.1; o->pcg_max_iterations = 500;
}

int oracle_ba_solve(const ThbBaProblem* p, const ThbBaOptions* o, ThbBaSummary* s) {
  if (!p || !o || !s) return THB_E_INVALID_ARGUMENT;
  oracle::BaOracle ba(*p, *o);
  const int rc = ba.Init();
  if (rc != THB_OK) return rc;
  return ba.Solve(s);
}

// Ambient residuals/Jacobians of every block (same contract as thb_ba_evaluate).
int oracle_ba_evaluate(const ThbBaProblem* p, double* residuals, double* jac_cam, double* jac_intr,
                       double* jac_pt, uint8_t* ok) {
  if (!p) return THB_E_INVALID_ARGUMENT;
  ThbBaOptions o; oracle_ba_default_options(&o);
  oracle::BaOracle ba(*p, o);
  const int rc = ba.Init();
  if (rc != THB_OK) return rc;
  for (int i = 0; i < p->num_observations; ++i) {
    double r[2] = {0, 0}, jc[12] = {0}, ji[2 * THB_INTR_STRIDE] = {0}, jp[8] = {0};
    const bool good = ba.EvalObsAmbient(ba.x, i, true, r, jc, ji, jp);
    if (ok) ok[i] = good ? 1 : 0;
    if (!good) { std::memset(r, 0, sizeof(r)); std::memset(jc, 0, sizeof(jc)); std::memset(ji, 0, sizeof(ji)); std::memset(jp, 0, sizeof(jp)); }
    if (residuals) std::memcpy(residuals + 2 * (size_t)i, r, sizeof(r));
    if (jac_cam) std::memcpy(jac_cam + 12 * (size_t)i, jc, sizeof(jc));
    if (jac_intr) std::memcpy(jac_intr + 2 * THB_INTR_STRIDE * (size_t)i, ji, sizeof(ji));
    if (jac_pt) std::memcpy(jac_pt + 8 * (size_t)i, jp, sizeof(jp));
  }
  return THB_OK;
}

// Cost 0.5*sum rho(|r|^2) at the given parameters (all blocks, fixed ones included).
int oracle_ba_cost(const ThbBaProblem* p, const ThbBaOptions* o, double* cost) {
  if (!p || !o || !cost) return THB_E_INVALID_ARGUMENT;
  oracle::BaOracle ba(*p, *o);
  const int rc = ba.Init();
  if (rc != THB_OK) return rc;
  double c0 = 0.0, c1 = 0.0;
  const bool ok0 = ba.Evaluate(ba.x, false, &c0, false);
  const bool ok1 = ba.Evaluate(ba.x, false, &c1, true);
  *cost = c0 + c1;
  return ok0 && ok1 ? THB_OK : THB_E_NUMERICAL;
}

// SetOutlierTracksToUnestimated (sfm/set_outlier_tracks_to_unestimated.cc:62-137) over the flattened problem: the
// observations are those of estimated views, pt_const marks tracks that are not estimated. status as THB_OUTLIER_*.
int oracle_set_outlier_tracks(const ThbBaProblem* p, double max_err, double min_angle_deg, int32_t* status) {
  if (!p || !status) return THB_E_INVALID_ARGUMENT;
  const int np = p->num_points, no = p->num_observations;
  std::vector<std::vector<int>> obs(np);
  for (int i = 0; i < no; ++i) obs[p->obs_pt[i]].push_back(i);
  const double sq_max = max_err * max_err, cos_min = std::cos(min_angle_deg * 3.14159265358979323846 / 180.0);
  int removed = 0;
  for (int t = 0; t < np; ++t) {
    if (p->pt_const && p->pt_const[t]) { status[t] = THB_OUTLIER_SKIPPED; continue; }
    const double* X = p->pts + 4 * (size_t)t;
    int st = THB_OUTLIER_KEPT, nproj = 0;
    double sum = 0.0;
    std::vector<std::array<double, 3>> rays;
    for (int i : obs[t]) {
      const int c = p->obs_cam[i], g = p->cam_group[c];
      const double* ext = p->cam_ext + 6 * (size_t)c;
      std::array<double, 3> ray = {X[0] / X[3] - ext[0], X[1] / X[3] - ext[1], X[2] / X[3] - ext[2]};
      const double nr = std::sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
      for (double& v : ray) v /= nr;
      rays.push_back(ray);
      const double adj[3] = {X[0] - X[3] * ext[0], X[1] - X[3] * ext[1], X[2] - X[3] * ext[2]};
      double pc[3], pix[2] = {0, 0};
      oracle::AngleAxisRotatePoint(ext + 3, adj, pc);
      if (pc[2] / X[3] < 0.0) { st = THB_OUTLIER_BAD_REPROJECTION; break; }
      oracle::ProjectByModel<double>(p->intr_model[g], p->intr + (size_t)g * THB_INTR_STRIDE, pc, pix);
      const double ex = pix[0] - p->obs_xy[2 * (size_t)i], ey = pix[1] - p->obs_xy[2 * (size_t)i + 1];
      sum += ex * ex + ey * ey;
      ++nproj;
    }
    if (st == THB_OUTLIER_KEPT && sum / static_cast<double>(nproj) > sq_max) st = THB_OUTLIER_BAD_REPROJECTION;
    if (st == THB_OUTLIER_KEPT) {
      bool wide = false;
      for (size_t a = 0; a < rays.size() && !wide; ++a)
        for (size_t b = a + 1; b < rays.size(); ++b)
          if (rays[a][0] * rays[b][0] + rays[a][1] * rays[b][1] + rays[a][2] * rays[b][2] < cos_min) { wide = true; break; }
      if (!wide) st = THB_OUTLIER_BAD_ANGLE;
    }
    status[t] = st;
    removed += st > 0;
  }
  return removed;
}

// ceres::Covariance (default options: loss applied, tangent space) for the BundleAdjustView(s) / BundleAdjustTrack(s) problems
// (bundle_adjuster.cc:660-773): with only cameras or only points free every block's covariance is (J_b^T J_b)^-1 of its own
// tangent-space Jacobian. Same contract as thb_ba_covariance.
int oracle_ba_covariance(const ThbBaProblem* p, const ThbBaOptions* o, double* cam_cov, uint8_t* cam_ok, double* pt_cov, uint8_t* pt_ok) {
  if (!p || !o) return THB_E_INVALID_ARGUMENT;
  ThbBaOptions opt = *o; opt.jacobi_scaling = 0;
  oracle::BaOracle ba(*p, opt);
  const int rc = ba.Init();
  if (rc != THB_OK) return rc;
  return ba.Covariance(cam_cov, cam_ok, pt_cov, pt_ok);
}

// theia::SelectGoodTracksForBundleAdjustment (sfm/select_good_tracks_for_bundle_adjustment.cc:263-325) over the flat problem
// (estimated views and tracks only). Unspecified orders of the reference (unordered containers) are fixed to ascending index:
// see include/theia_b200.h. Returns the number of chosen tracks.
int oracle_select_good_tracks(const ThbBaProblem* p, const uint8_t* cam_selected, int long_thr, int cell_size, int min_per_view, uint8_t* selected) {
  if (!p || !selected || cell_size <= 0) return THB_E_INVALID_ARGUMENT;
  const int nc = p->num_cameras, np = p->num_points, no = p->num_observations;
  // ComputeStatisticsForTrack (:80-107): truncated length, mean squared reprojection error over every view of the track
  std::vector<int> len(np, 0);
  std::vector<double> sum(np, 0.0), mean(np, 0.0);
  for (int i = 0; i < no; ++i) {
    const int c = p->obs_cam[i], g = p->cam_group[c], t = p->obs_pt[i];
    const double* X = p->pts + 4 * (size_t)t;
    const double* ext = p->cam_ext + 6 * (size_t)c;
    const double adj[3] = {X[0] - X[3] * ext[0], X[1] - X[3] * ext[1], X[2] - X[3] * ext[2]};
    double pc[3], pix[2] = {0, 0};
    oracle::AngleAxisRotatePoint(ext + 3, adj, pc);
    oracle::ProjectByModel<double>(p->intr_model[g], p->intr + (size_t)g * THB_INTR_STRIDE, pc, pix);
    const double ex = pix[0] - p->obs_xy[2 * (size_t)i], ey = pix[1] - p->obs_xy[2 * (size_t)i + 1];
    sum[t] += ex * ex + ey * ey;
    ++len[t];
  }
  for (int t = 0; t < np; ++t) { mean[t] = sum[t] / static_cast<double>(len[t]); len[t] = std::min(len[t], long_thr); }
  std::vector<std::vector<int>> by_cam(nc);
  for (int i = 0; i < no; ++i) by_cam[p->obs_cam[i]].push_back(i);
  const double inv = 1.0 / cell_size;
  // SelectBestTracksFromEachImageGridCell (:152-195): min over std::pair<int, double> in every occupied cell
  for (int c = 0; c < nc; ++c) {
    if (cam_selected && !cam_selected[c]) continue;
    std::map<std::pair<int, int>, int> best;  // cell -> observation
    for (int i : by_cam[c]) {
      const std::pair<int, int> cell(static_cast<int>(p->obs_xy[2 * (size_t)i] * inv), static_cast<int>(p->obs_xy[2 * (size_t)i + 1] * inv));
      auto it = best.find(cell);
      if (it == best.end()) { best[cell] = i; continue; }
      const int a = p->obs_pt[i], b = p->obs_pt[it->second];
      if (std::make_pair(len[a], mean[a]) < std::make_pair(len[b], mean[b])) it->second = i;
    }
    for (const auto& kv : best) selected[p->obs_pt[kv.second]] = 1;
  }
  // SelectTopRankedTracksInView (:199-254): partial_sort on pair<TrackId, statistics> = ascending track id
  for (int c = 0; c < nc; ++c) {
    if (cam_selected && !cam_selected[c]) continue;
    std::vector<int> cand;
    int n_opt = 0;
    const int n_est = (int)by_cam[c].size();
    for (int i : by_cam[c]) { if (selected[p->obs_pt[i]]) ++n_opt; else cand.push_back(p->obs_pt[i]); }
    if (n_opt >= min_per_view || n_opt == n_est) continue;
    const int needed = std::min(min_per_view - n_opt, n_est - n_opt);
    std::sort(cand.begin(), cand.end());
    for (int k = 0; k < needed; ++k) selected[cand[k]] = 1;
  }
  int count = 0;
  for (int t = 0; t < np; ++t) count += selected[t] != 0;
  return count;
}

// Thread control for the timed baselines: torchrun exports OMP_NUM_THREADS=1, the CPU arm must say how many it used.
int oracle_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  int used = 0;
#pragma omp parallel
  {
#pragma omp single
    used = omp_get_num_threads();
  }
  return used;
}

// SphereManifold<4> helpers exposed for unit tests.
void oracle_sphere_plus(const double* x, const double* d, double* out) { oracle::SpherePlus(x, d, out); }
void oracle_sphere_plus_jacobian(const double* x, double* J) { oracle::SpherePlusJacobian(x, J); }

}  // extern "C"
