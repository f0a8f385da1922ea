// Two-view geometry of image pairs from PIXEL correspondences, batched: EstimateTwoViewInfo (calibrated branch) and
// TwoViewMatchGeometricVerification::VerifyMatches. Included by ransac_kernels.cu (it chains the RANSAC launches of that
// file with the kernels below on one stream; nothing is copied back between the stages).
//   k_tv_prepare    per pair: calibrated?, the pair's Sampson threshold scaled by image size and focal lengths
//                   (estimate_twoview_info.cc:155-167, reconstruction_estimator_utils.cc:97-110), size gates
//   k_tv_normalize  NormalizeFeatures (estimate_twoview_info.cc:67-103): pixel -> Camera::PixelToNormalizedCoordinates ->
//                   hnormalized, through the inverse camera model of each view (camera_models.cuh::pixel_to_camera)
//   k_tv_info       TwoViewInfo from the RANSAC result (:173-189)
//   k_tv_verify     one CTA per pair: SetupCameras, TriangulatePoints, BundleAdjustTwoViews, the reprojection filter
//                   (two_view_match_geometric_verification.cc:186-327)
#ifndef THB_TWO_VIEW_CUH_
#define THB_TWO_VIEW_CUH_

#include "ba_device.cuh"

namespace thb {
namespace {

constexpr int TV_THREADS = 256;
#ifndef THB_TV_CTAS
#define THB_TV_CTAS 1
#endif
constexpr int TV_CTAS = THB_TV_CTAS;  // resident verification CTAs per SM; 2 (128 registers, 1.3 KB stack) measured slower: 22.8 vs 20 ms on the C5 batch

// ComputeResolutionScaledThreshold (reconstruction_estimator_utils.cc:97-110)
__device__ __host__ inline double resolution_scaled_threshold(double threshold_pixels, int w, int h) {
  if (w == 0 && h == 0) return threshold_pixels;
  return threshold_pixels * (double)(w > h ? w : h) / 1024.0;
}

// flags[pair]: bit 0 = skip the pair (not calibrated / fewer matches than the gate)
__global__ void k_tv_prepare(int np, const long long* __restrict__ off, const ThbViewIntrinsics* __restrict__ i1,
                             const ThbViewIntrinsics* __restrict__ i2, ThbTwoViewOptions O, int verify, double* __restrict__ thresh,
                             uint8_t* __restrict__ skip, int* __restrict__ uncalibrated) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const ThbViewIntrinsics a = i1[p], b = i2[p];
  const double s1 = resolution_scaled_threshold(O.max_sampson_error_pixels, a.image_width, a.image_height);
  const double s2 = resolution_scaled_threshold(O.max_sampson_error_pixels, b.image_width, b.image_height);
  thresh[p] = s1 * s2 / (a.params[0] * b.params[0]);
  const bool calibrated = a.focal_length_is_set && b.focal_length_is_set;
  if (!calibrated || num_intrinsics(a.model) < 0 || num_intrinsics(b.model) < 0) atomicExch(uncalibrated, 1);
  const long long n = off[p + 1] - off[p];
  skip[p] = (!calibrated || (verify && n < O.min_num_inlier_matches)) ? 1 : 0;  // VerifyMatches :117-119
}

__global__ void k_tv_normalize(const long long* __restrict__ off, const double* __restrict__ px, const ThbViewIntrinsics* __restrict__ i1,
                               const ThbViewIntrinsics* __restrict__ i2, double* __restrict__ out) {
  const int p = blockIdx.x;
  __shared__ ThbViewIntrinsics a, b;
  if (threadIdx.x == 0) { a = i1[p]; b = i2[p]; }
  __syncthreads();
  for (long long i = off[p] + threadIdx.x; i < off[p + 1]; i += blockDim.x) {
    const double u[2] = {px[4 * i], px[4 * i + 1]}, v[2] = {px[4 * i + 2], px[4 * i + 3]};
    double q1[3], q2[3];
    pixel_to_camera(a.model, a.params, u, q1);
    pixel_to_camera(b.model, b.params, v, q2);
    out[4 * i] = q1[0] / q1[2]; out[4 * i + 1] = q1[1] / q1[2];
    out[4 * i + 2] = q2[0] / q2[2]; out[4 * i + 3] = q2[1] / q2[2];
  }
}

__global__ void k_tv_info(int np, const ThbRelPoseResult* __restrict__ res, const ThbViewIntrinsics* __restrict__ i1,
                          const ThbViewIntrinsics* __restrict__ i2, const uint8_t* __restrict__ skip, ThbTwoViewInfo* __restrict__ info) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  ThbTwoViewInfo o;
  memset(&o, 0, sizeof(o));
  const ThbRelPoseResult r = res[p];
  o.success = (!skip[p] && r.success) ? 1 : 0;
  if (o.success) {
    eigen_matrix_to_angle_axis(r.rotation, o.rotation_2);
    for (int k = 0; k < 3; ++k) o.position_2[k] = r.position[k];
    o.focal_length_1 = i1[p].params[0]; o.focal_length_2 = i2[p].params[0];
    o.num_verified_matches = r.num_inliers;
    o.visibility_score = 0;  // computed from the still-empty inlier list upstream (estimate_twoview_info.cc:186-189)
    o.num_ransac_iterations = r.num_iterations;
  }
  info[p] = o;
}

// ---- VerifyMatches after the two RANSACs ---------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void tv_block_sum(double (&v)[NV], double (*red)[48], double* tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) red[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int ww = 0; ww < TV_THREADS / 32; ++ww) s += red[ww][threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
}

// TriangulateMidpoint for two rays (triangulation.cc:130-157): origins o1, o2, unit directions d1, d2
__device__ bool tv_midpoint(const double* o1, const double* d1, const double* o2, const double* d2, double z[4]) {
  double A[4][4] = {}, b[4] = {};
  for (int ray = 0; ray < 2; ++ray) {
    const double* dd = ray ? d2 : d1; const double* oo = ray ? o2 : o1;
    const double d[4] = {dd[0], dd[1], dd[2], 0.0}, o[4] = {oo[0], oo[1], oo[2], 1.0};
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

// AcceptableReprojectionError (two_view_match_geometric_verification.cc:70-83): Camera::ProjectPoint depth >= 0 and squared
// pixel error below the bound
__device__ bool tv_reprojection_ok(int model, const double* Kp, const double* rec, const double* X, const double* feat, double sq_max) {
  const double adj[3] = {X[0] - X[3] * rec[CD_C], X[1] - X[3] * rec[CD_C + 1], X[2] - X[3] * rec[CD_C + 2]};
  double pc[3];
  rot_apply(rec + CD_W, rec[CD_A], rec[CD_B], rec[0] * rec[0] + rec[1] * rec[1] + rec[2] * rec[2], adj, pc);
  if (pc[2] / X[3] < 0.0) return false;
  double pix[2];
  if (!project<-1, double, double>(model, Kp, pc, pix)) return false;
  const double ex = feat[0] - pix[0], ey = feat[1] - pix[1];
  return ex * ex + ey * ey < sq_max;
}

struct TvShared {
  double rec1[CAMD], rec2[CAMD], rec2c[CAMD], cam2[6], cam2c[6], K1[KS], K2[KS];
  double tot[48], red[TV_THREADS / 32][48], gm[TV_THREADS / 32];
  int model1, model2, go, cur, n_tri;
};

// d_res / d_mask: result and inlier flags of the relative-pose RANSAC on the normalised correspondences. pts0 / pts1: two
// point buffers [total * 4]; pscale [total * 4]; tri [total] (1 = triangulated match).
__global__ void __launch_bounds__(TV_THREADS, TV_CTAS) k_tv_verify(const long long* __restrict__ off_all, const double* __restrict__ px_all,
                                                          const ThbViewIntrinsics* __restrict__ i1, const ThbViewIntrinsics* __restrict__ i2,
                                                          ThbTwoViewOptions O, const ThbRelPoseResult* __restrict__ res_all,
                                                          const uint8_t* __restrict__ inl_all, const ThbRelPoseResult* __restrict__ hom_all,
                                                          const uint8_t* __restrict__ skip, double* __restrict__ pts0, double* __restrict__ pts1,
                                                          double* __restrict__ pscale_all, uint8_t* __restrict__ tri_all,
                                                          const uint8_t* __restrict__ zeros, const int* __restrict__ two_groups,
                                                          ThbTwoViewInfo* __restrict__ info_all, uint8_t* __restrict__ verified_all) {
  __shared__ TvShared S;
  const int pair = blockIdx.x, t = threadIdx.x;
  const long long off = off_all[pair];
  const int n = (int)(off_all[pair + 1] - off);
  const double* px = px_all + 4 * off;
  const uint8_t* inl = inl_all + off;
  uint8_t* tri = tri_all + off;
  uint8_t* verified = verified_all + off;
  double* pbuf[2] = {pts0 + 4 * off, pts1 + 4 * off};
  double* pscale = pscale_all + 4 * off;
  ThbTwoViewInfo* info = info_all + pair;
  const ThbRelPoseResult R = res_all[pair];
  for (int i = t; i < n; i += TV_THREADS) verified[i] = 0;
  if (t == 0) {
    ThbTwoViewInfo o;
    memset(&o, 0, sizeof(o));
    if (!skip[pair]) o.num_homography_inliers = hom_all[pair].num_inliers;  // twoview_info->num_homography_inliers (:127)
    if (!skip[pair] && R.success) {
      eigen_matrix_to_angle_axis(R.rotation, o.rotation_2);
      for (int k = 0; k < 3; ++k) o.position_2[k] = R.position[k];
      o.focal_length_1 = i1[pair].params[0]; o.focal_length_2 = i2[pair].params[0];
      o.num_verified_matches = R.num_inliers;
      o.num_ransac_iterations = R.num_iterations;
    }
    *info = o;
  }
  __syncthreads();
  if (skip[pair] || !R.success || R.num_inliers < O.min_num_inlier_matches) return;  // :137-146 (success stays 0)
  const int n_inl = R.num_inliers;
  if (!(O.bundle_adjustment && n_inl > O.min_num_inlier_matches)) {  // :172-177 skipped: the RANSAC inliers are the verified matches
    for (int i = t; i < n; i += TV_THREADS) verified[i] = inl[i];
    if (t == 0) info->success = n_inl > O.min_num_inlier_matches ? 1 : 0;
    return;
  }
  // ---- SetupCameras (:56-67): camera 1 at the origin, camera 2 at (rotation_2, position_2)
  if (t == 0) {
    const ThbViewIntrinsics a = i1[pair], b = i2[pair];
    S.model1 = a.model; S.model2 = b.model;
    for (int k = 0; k < KS; ++k) { S.K1[k] = a.params[k]; S.K2[k] = b.params[k]; }
    const double c1[6] = {0, 0, 0, 0, 0, 0};
    cam_derive_record(c1, S.rec1);
    for (int k = 0; k < 3; ++k) { S.cam2[k] = info->position_2[k]; S.cam2[3 + k] = info->rotation_2[k]; }
    cam_derive_record(S.cam2, S.rec2);
    for (int k = 0; k < 6; ++k) { S.rec1[CD_SCALE + k] = 0.0; S.rec2[CD_SCALE + k] = 1.0; }
    S.cur = 0;
  }
  __syncthreads();
  // ---- TriangulatePoints (:186-257)
  const double cos_min = cos(O.min_triangulation_angle_degrees * 3.14159265358979323846 / 180.0);
  const double sq_tri = O.triangulation_max_reprojection_error * O.triangulation_max_reprojection_error;
  double cnt[1] = {0.0};
  for (int i = t; i < n; i += TV_THREADS) {
    uint8_t ok = 0;
    if (inl[i]) {
      const double f1[2] = {px[4 * i], px[4 * i + 1]}, f2[2] = {px[4 * i + 2], px[4 * i + 3]};
      double q1[3], q2[3], d2[3];
      pixel_to_camera(S.model1, S.K1, f1, q1);
      pixel_to_camera(S.model2, S.K2, f2, q2);
      const double* w2 = S.rec2 + CD_W;
      rot_apply(w2, -S.rec2[CD_A], S.rec2[CD_B], w2[0] * w2[0] + w2[1] * w2[1] + w2[2] * w2[2], q2, d2);  // R^T q2
      const double n1 = sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2]), n2 = sqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]);
      const double d1[3] = {q1[0] / n1, q1[1] / n1, q1[2] / n1};
      d2[0] /= n2; d2[1] /= n2; d2[2] /= n2;
      if (d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2] < cos_min) {
        const double o1[3] = {0, 0, 0};
        double X[4];
        if (tv_midpoint(o1, d1, S.cam2, d2, X) && tv_reprojection_ok(S.model1, S.K1, S.rec1, X, f1, sq_tri) &&
            tv_reprojection_ok(S.model2, S.K2, S.rec2, X, f2, sq_tri)) {
          ok = 1;
          for (int k = 0; k < 4; ++k) pbuf[0][4 * i + k] = X[k];
        }
      }
    }
    tri[i] = ok;
    cnt[0] += ok;
  }
  tv_block_sum<1>(cnt, S.red, S.tot);
  const int n_tri = (int)S.tot[0];
  if (t == 0) info->num_triangulated = n_tri;
  if (n_tri < O.min_num_inlier_matches) return;  // :269-271
  // ---- BundleAdjustTwoViews (bundle_adjust_two_views.cc:112-193): camera 1 constant, camera 2 extrinsics + 4-vector points
  // free, no loss, exact Schur solve; Ceres defaults (100 iterations, 1e-6 / 1e-10 / 1e-8, radius 1e4 .. 1e16)
  BaConst K;
  K.nc = 2; K.ng = 2; K.np = n; K.no = 2 * n;
  K.cam_group = two_groups; K.intr_model = &S.model1;  // model1, model2 are adjacent ints
  K.cam_const = zeros; K.intr_const = nullptr; K.pt_const = zeros; K.intr_slot = nullptr;
  K.loss_type = THB_LOSS_TRIVIAL; K.loss_width = 1.0;
  BaState St;
  St.cam = nullptr; St.camd = nullptr; St.intr = S.K1;  // K1, K2 are adjacent blocks of KS doubles
  const double2 one2 = make_double2(1.0, 1.0);
  // one pass at x: mode 0 = column norms of the unscaled Jacobian; mode 1 = Schur complement + rhs with the current radius
  auto pass_j = [&](int mode, double radius) {
    double v[42];
#pragma unroll
    for (int k = 0; k < 42; ++k) v[k] = 0.0;
    St.pts = pbuf[S.cur];
    for (int i = t; i < n; i += TV_THREADS) {
      if (!tri[i]) continue;
      double r1[2], r2[2], jc1[12], jc2[12], jp1[8], jp2[8], hc1 = 0.0, hc2 = 0.0;
      const double2 xy1 = make_double2(px[4 * i], px[4 * i + 1]), xy2 = make_double2(px[4 * i + 2], px[4 * i + 3]);
      const double* cs = mode ? S.rec2 : nullptr;
      const double* ps = mode ? pscale : nullptr;
      const bool ok1 = eval_obs<-1, 4, 0, false>(K, St, 0, i, xy1, one2, cs, ps, nullptr, r1, jc1, jp1, nullptr, &hc1, 0, 0, 0, 0, nullptr, S.rec1);
      const bool ok2 = eval_obs<-1, 4, 0, false>(K, St, 1, i, xy2, one2, cs, ps, nullptr, r2, jc2, jp2, nullptr, &hc2, 0, 0, 0, 0, nullptr, S.rec2);
      if (!ok1 || !ok2) { v[41] += 1.0; continue; }
      v[39] += hc1 + hc2;
      if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) v[27 + k] += jc2[k] * jc2[k] + jc2[6 + k] * jc2[6 + k];
#pragma unroll
        for (int k = 0; k < 4; ++k) pscale[4 * i + k] = 1.0 / (1.0 + sqrt(jp1[k] * jp1[k] + jp1[4 + k] * jp1[4 + k] + jp2[k] * jp2[k] + jp2[4 + k] * jp2[4 + k]));
        continue;
      }
      double V[16], gp[4], W[24], gc[6];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        gp[a] = jp1[a] * r1[0] + jp1[4 + a] * r1[1] + jp2[a] * r2[0] + jp2[4 + a] * r2[1];
#pragma unroll
        for (int b = 0; b < 4; ++b) V[a * 4 + b] = jp1[a] * jp1[b] + jp1[4 + a] * jp1[4 + b] + jp2[a] * jp2[b] + jp2[4 + a] * jp2[4 + b];
      }
      double gm = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) gm = fmax(gm, fabs(gp[a] / pscale[4 * i + a]));
      v[40] = fmax(v[40], gm);
#pragma unroll
      for (int a = 0; a < 4; ++a) V[a * 4 + a] += fmin(fmax(V[a * 4 + a], 1e-6), 1e32) / radius;
      double Vi[16];
      if (!spd_inverse<4>(V, Vi)) { v[41] += 1.0; continue; }
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        gc[a] = jc2[a] * r2[0] + jc2[6 + a] * r2[1];
#pragma unroll
        for (int b = 0; b < 4; ++b) W[a * 4 + b] = jc2[a] * jp2[b] + jc2[6 + a] * jp2[4 + b];
      }
      double T[24];  // W V^-1
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) T[a * 4 + b] = W[a * 4] * Vi[b] + W[a * 4 + 1] * Vi[4 + b] + W[a * 4 + 2] * Vi[8 + b] + W[a * 4 + 3] * Vi[12 + b];
      int e = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          v[e++] += jc2[a] * jc2[b] + jc2[6 + a] * jc2[6 + b] - (T[a * 4] * W[b * 4] + T[a * 4 + 1] * W[b * 4 + 1] + T[a * 4 + 2] * W[b * 4 + 2] + T[a * 4 + 3] * W[b * 4 + 3]);
        }
        v[21 + a] += gc[a] - (T[a * 4] * gp[0] + T[a * 4 + 1] * gp[1] + T[a * 4 + 2] * gp[2] + T[a * 4 + 3] * gp[3]);
        v[27 + a] += jc2[a] * jc2[a] + jc2[6 + a] * jc2[6 + a];
        v[33 + a] += gc[a];
      }
    }
    // the maximum rides through the sums as a separate warp reduction
    const double gm = warp_max(v[40]);
    v[40] = 0.0;
    tv_block_sum<42>(v, S.red, S.tot);
    if ((t & 31) == 0) S.gm[t >> 5] = gm;
    __syncthreads();
    if (t == 0) { double m = 0.0; for (int w = 0; w < TV_THREADS / 32; ++w) m = fmax(m, S.gm[w]); S.tot[40] = m; }
    __syncthreads();
  };
  // back-substitution with the camera step yc (scaled), candidate points into the other buffer, model cost change, candidate cost
  auto pass_b = [&](double radius, const double* yc) {
    double v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = 0.0;
    St.pts = pbuf[S.cur];
    double* cand = pbuf[S.cur ^ 1];
    for (int i = t; i < n; i += TV_THREADS) {
      if (!tri[i]) continue;
      double r1[2], r2[2], jc1[12], jc2[12], jp1[8], jp2[8], hc1 = 0.0, hc2 = 0.0;
      const double2 xy1 = make_double2(px[4 * i], px[4 * i + 1]), xy2 = make_double2(px[4 * i + 2], px[4 * i + 3]);
      const bool ok1 = eval_obs<-1, 4, 0, false>(K, St, 0, i, xy1, one2, S.rec2, pscale, nullptr, r1, jc1, jp1, nullptr, &hc1, 0, 0, 0, 0, nullptr, S.rec1);
      const bool ok2 = eval_obs<-1, 4, 0, false>(K, St, 1, i, xy2, one2, S.rec2, pscale, nullptr, r2, jc2, jp2, nullptr, &hc2, 0, 0, 0, 0, nullptr, S.rec2);
      if (!ok1 || !ok2) { v[5] += 1.0; continue; }
      double V[16], b4[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        b4[a] = jp1[a] * r1[0] + jp1[4 + a] * r1[1] + jp2[a] * r2[0] + jp2[4 + a] * r2[1];
#pragma unroll
        for (int b = 0; b < 4; ++b) V[a * 4 + b] = jp1[a] * jp1[b] + jp1[4 + a] * jp1[4 + b] + jp2[a] * jp2[b] + jp2[4 + a] * jp2[4 + b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) V[a * 4 + a] += fmin(fmax(V[a * 4 + a], 1e-6), 1e32) / radius;
      double Vi[16];
      if (!spd_inverse<4>(V, Vi)) { v[5] += 1.0; continue; }
      // jy = Jc2 yc (2); b4 -= Jp2^T jy; yp = V^-1 b4
      double jy[2] = {0.0, 0.0};
#pragma unroll
      for (int a = 0; a < 6; ++a) { jy[0] += jc2[a] * yc[a]; jy[1] += jc2[6 + a] * yc[a]; }
      double yp[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) b4[a] -= jp2[a] * jy[0] + jp2[4 + a] * jy[1];
#pragma unroll
      for (int a = 0; a < 4; ++a) yp[a] = Vi[a * 4] * b4[0] + Vi[a * 4 + 1] * b4[1] + Vi[a * 4 + 2] * b4[2] + Vi[a * 4 + 3] * b4[3];
      // step = -y: model cost change -(J s)^T (r + J s / 2)
      double m1[2] = {0.0, 0.0}, m2[2] = {-jy[0], -jy[1]};
#pragma unroll
      for (int a = 0; a < 4; ++a) { m1[0] -= jp1[a] * yp[a]; m1[1] -= jp1[4 + a] * yp[a]; m2[0] -= jp2[a] * yp[a]; m2[1] -= jp2[4 + a] * yp[a]; }
      v[0] += -(m1[0] * (r1[0] + m1[0] / 2.0) + m1[1] * (r1[1] + m1[1] / 2.0) + m2[0] * (r2[0] + m2[0] / 2.0) + m2[1] * (r2[1] + m2[1] / 2.0));
      double xo[4], xn[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        xo[a] = St.pts[4 * i + a];
        xn[a] = xo[a] + (-yp[a] * pscale[4 * i + a]);
        cand[4 * i + a] = xn[a];
        v[2] += (xn[a] - xo[a]) * (xn[a] - xo[a]);
        v[3] += xn[a] * xn[a];
      }
    }
    __syncthreads();  // the candidate buffer is complete
    BaState Sc = St;
    Sc.pts = cand;
    for (int i = t; i < n; i += TV_THREADS) {
      if (!tri[i]) continue;
      double r1[2], r2[2];
      const double2 xy1 = make_double2(px[4 * i], px[4 * i + 1]), xy2 = make_double2(px[4 * i + 2], px[4 * i + 3]);
      if (!eval_residual<-1>(K, Sc, 0, i, xy1, one2, r1, S.rec1) || !eval_residual<-1>(K, Sc, 1, i, xy2, one2, r2, S.rec2c)) { v[4] += 1.0; continue; }
      v[1] += 0.5 * (r1[0] * r1[0] + r1[1] * r1[1]) + 0.5 * (r2[0] * r2[0] + r2[1] * r2[1]);
    }
    tv_block_sum<6>(v, S.red, S.tot);
  };

  pass_j(0, 1.0);
  if (S.tot[41] > 0.0) { return; }  // evaluation failed at the start: BundleAdjustTwoViews reports FAILURE -> return false (:296-298)
  if (t == 0) for (int k = 0; k < 6; ++k) S.rec2[CD_SCALE + k] = 1.0 / (1.0 + sqrt(S.tot[27 + k]));
  __syncthreads();
  double radius = 1e4, decrease_factor = 2.0, x_cost = 0.0, x_norm = 0.0, min_cost = 0.0, initial_cost = 0.0, mcc_cam = 0.0;
  double yc[6] = {0, 0, 0, 0, 0, 0};
  bool step_ok = true, failure = false;
  int iteration = 0, invalid = 0;
  {  // |x| over the free blocks: camera 2 and the triangulated points
    double v[1] = {0.0};
    for (int i = t; i < n; i += TV_THREADS) if (tri[i]) for (int k = 0; k < 4; ++k) v[0] += pbuf[0][4 * i + k] * pbuf[0][4 * i + k];
    tv_block_sum<1>(v, S.red, S.tot);
    x_norm = S.tot[0];
    for (int k = 0; k < 6; ++k) x_norm += S.cam2[k] * S.cam2[k];
    x_norm = sqrt(x_norm);
  }
  for (;;) {
    // every thread follows the same control flow: the decisions below only use values every thread holds (S.tot, registers)
    if (iteration >= 100) break;
    if (radius <= 1e-32) break;
    pass_j(1, radius);
    if (step_ok) {  // x changed (or first iteration): cost and gradient at x
      x_cost = S.tot[39];
      if (iteration == 0) { initial_cost = x_cost; min_cost = x_cost; }
      double gmax = S.tot[40];
      for (int k = 0; k < 6; ++k) gmax = fmax(gmax, fabs(S.tot[33 + k] / S.rec2[CD_SCALE + k]));
      if (gmax <= 1e-10) break;
    }
    ++iteration;
    step_ok = false;
    bool valid = S.tot[41] == 0.0;
    if (valid) {
      double M[36];
      int e = 0;
      for (int a = 0; a < 6; ++a) for (int b = 0; b <= a; ++b) { M[a * 6 + b] = S.tot[e]; M[b * 6 + a] = S.tot[e]; ++e; }
      for (int a = 0; a < 6; ++a) M[a * 6 + a] += fmin(fmax(S.tot[27 + a], 1e-6), 1e32) / radius;
      double rhs[6];
      for (int a = 0; a < 6; ++a) rhs[a] = S.tot[21 + a];
      valid = spd_solve<6>(M, rhs, yc);
    }
    if (valid) {
      __syncthreads();
      if (t == 0) {
        for (int k = 0; k < 6; ++k) S.cam2c[k] = S.cam2[k] + (-yc[k] * S.rec2[CD_SCALE + k]);
        cam_derive_record(S.cam2c, S.rec2c);
        for (int k = 0; k < 6; ++k) S.rec2c[CD_SCALE + k] = S.rec2[CD_SCALE + k];
      }
      __syncthreads();
      pass_b(radius, yc);
      const double mcc = S.tot[0];
      valid = S.tot[5] == 0.0 && isfinite(mcc) && mcc > 0.0;
      mcc_cam = mcc;
    }
    if (!valid) {
      if (++invalid >= 5) { failure = true; break; }
      radius /= decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    const double cand_cost = S.tot[4] > 0.0 ? 1.7976931348623157e308 : S.tot[1];
    double sn = S.tot[2], cn = S.tot[3];
    for (int k = 0; k < 6; ++k) { sn += (S.cam2c[k] - S.cam2[k]) * (S.cam2c[k] - S.cam2[k]); cn += S.cam2c[k] * S.cam2c[k]; }
    if (sqrt(sn) <= 1e-8 * (x_norm + 1e-8)) break;
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= 1e-6 * x_cost) break;
    const double rel = cand_cost >= 1.7976931348623157e308 ? -1.7976931348623157e308 : cost_change / mcc_cam;
    __syncthreads();  // everybody has read S.tot / S.cam2 of this round
    if (rel > 1e-3) {
      if (t == 0) {
        for (int k = 0; k < 6; ++k) S.cam2[k] = S.cam2c[k];
        for (int k = 0; k < CAMD; ++k) S.rec2[k] = S.rec2c[k];
        S.cur ^= 1;
      }
      x_norm = sqrt(cn);
      step_ok = true;
      min_cost = fmin(min_cost, cand_cost);
      const double u = 2.0 * rel - 1.0;
      radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - u * u * u));
      decrease_factor = 2.0;
      // untriangulated slots of the new current buffer are never read; nothing to copy
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
    }
    __syncthreads();
  }
  __syncthreads();
  if (t == 0) { info->ba_iterations = iteration; info->ba_initial_cost = initial_cost; info->ba_final_cost = min_cost; }
  if (failure) return;  // summary.success == false (:296-298)
  // ---- reprojection filter after BA (:300-316) and the updated relative pose (:318-324)
  const double sq_fin = O.final_max_reprojection_error * O.final_max_reprojection_error;
  const double* P = pbuf[S.cur];
  double kept[1] = {0.0};
  for (int i = t; i < n; i += TV_THREADS) {
    if (!tri[i]) continue;
    const double f1[2] = {px[4 * i], px[4 * i + 1]}, f2[2] = {px[4 * i + 2], px[4 * i + 3]};
    const bool ok = tv_reprojection_ok(S.model1, S.K1, S.rec1, P + 4 * i, f1, sq_fin) && tv_reprojection_ok(S.model2, S.K2, S.rec2, P + 4 * i, f2, sq_fin);
    verified[i] = ok ? 1 : 0;
    kept[0] += ok ? 1.0 : 0.0;
  }
  tv_block_sum<1>(kept, S.red, S.tot);
  if (t == 0) {
    const int nv = (int)S.tot[0];
    const double pn = sqrt(S.cam2[0] * S.cam2[0] + S.cam2[1] * S.cam2[1] + S.cam2[2] * S.cam2[2]);
    for (int k = 0; k < 3; ++k) { info->rotation_2[k] = S.cam2[3 + k]; info->position_2[k] = S.cam2[k] / pn; }
    info->num_verified_matches = nv;
    info->success = nv > O.min_num_inlier_matches ? 1 : 0;
  }
}

}  // namespace
}  // namespace thb
#endif  // THB_TWO_VIEW_CUH_
