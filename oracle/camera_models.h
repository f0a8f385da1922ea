// ORACLE — test infrastructure only (see jet.h header).
//
// CPU restatement of the reference's camera projections and of ReprojectionError.
// Paths are relative to /root/reference/src/theia/sfm/camera/.
// T is double or oracle::Jet<N>; branch conditions use the scalar part, exactly like a
// ceres::Jet (so the derivative of the branch TAKEN is produced).
#ifndef ORACLE_CAMERA_MODELS_H_
#define ORACLE_CAMERA_MODELS_H_

#include <limits>

#include "jet.h"

namespace oracle {

enum ModelType {  // camera_intrinsics_model_type.h:46-56
  PINHOLE = 0, PINHOLE_RADIAL_TANGENTIAL = 1, FISHEYE = 2, FOV = 3,
  DIVISION_UNDISTORTION = 4, DOUBLE_SPHERE = 5, EXTENDED_UNIFIED = 6, ORTHOGRAPHIC = 7
};

inline int NumIntrinsics(int model) {
  switch (model) {
    case PINHOLE: return 7;                // pinhole_camera_model.h:86-94
    case FISHEYE: return 9;                // fisheye_camera_model.h:67-77
    case FOV: return 5;                    // fov_camera_model.h:69-75
    case DIVISION_UNDISTORTION: return 5;  // division_undistortion_camera_model.h:76-82
    case DOUBLE_SPHERE: return 7;          // double_sphere_camera_model.h:66-74
    case EXTENDED_UNIFIED: return 7;       // extended_unified_camera_model.h:66-74
    default: return -1;
  }
}

// ceres::AngleAxisRotatePoint (ceres/rotation.h; external), called at
// reprojection_error.h:84-86 and camera.cc:209-210.
template <typename T>
inline void AngleAxisRotatePoint(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const T theta = sqrt_(theta2);
    const T costheta = cos_(theta);
    const T sintheta = sin_(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    for (int i = 0; i < 3; ++i) out[i] = pt[i] * costheta + w_cross_pt[i] * sintheta + w[i] * tmp;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2],
                             aa[0] * pt[1] - aa[1] * pt[0]};
    for (int i = 0; i < 3; ++i) out[i] = pt[i] + w_cross_pt[i];
  }
}

// K = [f, a, s, cx, cy, ...] -> pixel, shared tail of the models with skew.
template <typename T>
inline void ApplyCalibrationWithSkew(const T* K, const T d[2], T pixel[2]) {
  pixel[0] = K[0] * d[0] + K[2] * d[1] + K[3];
  pixel[1] = K[0] * K[1] * d[1] + K[4];
}

// pinhole_camera_model.h:182-211 (projection), :244-260 (DistortPoint).
// K = [f, a, s, cx, cy, k1, k2]
template <typename T>
inline bool PinholeProject(const T* K, const T p[3], T pixel[2]) {
  const T n[2] = {p[0] / p[2], p[1] / p[2]};
  const T r_sq = n[0] * n[0] + n[1] * n[1];
  const T d = 1.0 + r_sq * (K[5] + K[6] * r_sq);
  const T dist[2] = {n[0] * d, n[1] * d};
  ApplyCalibrationWithSkew(K, dist, pixel);
  return true;
}

// double_sphere_camera_model.h:161-187 (projection), :215-249 (DistortPoint).
// K = [f, a, s, cx, cy, xi, alpha]. Returns false outside the valid cone; the pixel is
// then left unset in the reference (:233-235) — callers must not read it.
template <typename T>
inline bool DoubleSphereProject(const T* K, const T p[3], T pixel[2]) {
  const T& xi = K[5];
  const T& alpha = K[6];
  const T xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
  const T r2 = xx + yy;
  const T d1_2 = r2 + zz;
  const T d1 = sqrt_(d1_2);
  const T w1 = alpha > 0.5 ? (T(1.0) - alpha) / alpha : alpha / (T(1.0) - alpha);
  const T w2 = (w1 + xi) / sqrt_(T(2.0) * w1 * xi + xi * xi + T(1.0));
  if (p[2] <= -w2 * d1) return false;
  const T k = xi * d1 + p[2];
  const T kk = k * k;
  const T d2_2 = r2 + kk;
  const T d2 = sqrt_(d2_2);
  const T norm = alpha * d2 + (T(1.0) - alpha) * k;
  const T dist[2] = {p[0] / norm, p[1] / norm};
  ApplyCalibrationWithSkew(K, dist, pixel);
  return true;
}

// extended_unified_camera_model.h:161-187 (projection), :215-249 (DistortPoint).
// K = [f, a, s, cx, cy, alpha, beta]. Zero distorted point (and true) when norm<1e-3 or
// outside the hemisphere.
template <typename T>
inline bool ExtendedUnifiedProject(const T* K, const T p[3], T pixel[2]) {
  const T& alpha = K[5];
  const T& beta = K[6];
  const T xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
  const T r2 = xx + yy;
  const T rho2 = beta * r2 + zz;
  const T rho = sqrt_(rho2);
  const T norm = alpha * rho + (T(1.0) - alpha) * p[2];
  T dist[2] = {T(0.0), T(0.0)};
  bool zero = false;
  if (norm < 1e-3) zero = true;
  if (!zero && alpha > 0.5) {
    const T zn = p[2] / norm;
    const T C = (alpha - T(1.0)) / (alpha + alpha - T(1.0));
    if (zn < C) zero = true;
  }
  if (!zero) { dist[0] = p[0] / norm; dist[1] = p[1] / norm; }
  ApplyCalibrationWithSkew(K, dist, pixel);
  return true;
}

// fisheye_camera_model.h:163-189 (projection), :227-272 (DistortPoint).
// K = [f, a, s, cx, cy, k1, k2, k3, k4]
template <typename T>
inline bool FisheyeProject(const T* K, const T p[3], T pixel[2]) {
  const T r_sq = p[0] * p[0] + p[1] * p[1];
  T dist[2];
  if (r_sq < 1e-8) {
    dist[0] = p[0]; dist[1] = p[1];
  } else {
    const T r_num = sqrt_(r_sq);
    const T theta = atan2_(r_num, abs_(p[2]));
    const T theta_sq = theta * theta;
    const T theta_d = theta * (1.0 + K[5] * theta_sq + K[6] * theta_sq * theta_sq +
                               K[7] * theta_sq * theta_sq * theta_sq +
                               K[8] * theta_sq * theta_sq * theta_sq * theta_sq);
    dist[0] = theta_d * p[0] / r_num;
    dist[1] = theta_d * p[1] / r_num;
    if (p[2] < 0.0) { dist[0] = -dist[0]; dist[1] = -dist[1]; }
  }
  ApplyCalibrationWithSkew(K, dist, pixel);
  return true;
}

// fov_camera_model.h:156-181 (projection), :209-258 (DistortPoint). K = [f, a, cx, cy, omega]
template <typename T>
inline bool FovProject(const T* K, const T p[3], T pixel[2]) {
  const T n[2] = {p[0] / p[2], p[1] / p[2]};
  const T& omega = K[4];
  const T r_u_sq = n[0] * n[0] + n[1] * n[1];
  T r_d;
  if (omega < 1e-3) {
    r_d = (omega * omega * r_u_sq) / 3.0 - omega * omega / 12.0 + 1.0;
  } else if (r_u_sq < 1e-3) {
    const T tan_half_omega = tan_(omega / 2.0);
    r_d = (-2.0 * tan_half_omega * (4.0 * r_u_sq * tan_half_omega * tan_half_omega - 3.0)) /
          (3.0 * omega);
  } else {
    const T r_u = sqrt_(r_u_sq);
    r_d = atan_(2.0 * r_u * tan_(omega / 2.0)) / (r_u * omega);
  }
  const T focal_length_y = K[0] * K[1];
  pixel[0] = K[0] * (r_d * n[0]) + K[2];
  pixel[1] = focal_length_y * (r_d * n[1]) + K[3];
  return true;
}

// division_undistortion_camera_model.h:173-203 (projection), :263-297 (DistortPoint).
// K = [f, a, cx, cy, k]; the focal length is applied BEFORE the distortion.
template <typename T>
inline bool DivisionUndistortionProject(const T* K, const T p[3], T pixel[2]) {
  const T n[2] = {p[0] / p[2], p[1] / p[2]};
  const T focal_length_y = K[0] * K[1];
  const T u[2] = {K[0] * n[0], focal_length_y * n[1]};
  const T r_u_sq = u[0] * u[0] + u[1] * u[1];
  const T& k = K[4];
  const T denom = 2.0 * k * r_u_sq;
  const T inner_sqrt = 1.0 - 4.0 * k * r_u_sq;
  if (abs_(denom) < std::numeric_limits<double>::epsilon() || inner_sqrt < 0.0) {
    pixel[0] = u[0]; pixel[1] = u[1];
  } else {
    const T scale = (1.0 - sqrt_(inner_sqrt)) / denom;
    pixel[0] = u[0] * scale; pixel[1] = u[1] * scale;
  }
  pixel[0] += K[2];
  pixel[1] += K[3];
  return true;
}

template <typename T>
inline bool ProjectByModel(int model, const T* K, const T p[3], T pixel[2]) {
  switch (model) {  // create_reprojection_error_cost_function.h:54-136
    case PINHOLE: return PinholeProject(K, p, pixel);
    case FISHEYE: return FisheyeProject(K, p, pixel);
    case FOV: return FovProject(K, p, pixel);
    case DIVISION_UNDISTORTION: return DivisionUndistortionProject(K, p, pixel);
    case DOUBLE_SPHERE: return DoubleSphereProject(K, p, pixel);
    case EXTENDED_UNIFIED: return ExtendedUnifiedProject(K, p, pixel);
    default: return false;
  }
}

// reprojection_error.h:49-114. ext = [C(3), angle-axis(3)] (camera.h:202-204),
// X homogeneous [x,y,z,w]. Returns false if |X - wC|^2 < 1e-8 (:78-80) or if the camera
// model rejects the point (:109).
template <typename T>
inline bool ReprojectionError(int model, const T* ext, const T* K, const T* X,
                              const double obs[2], const double sqrt_info[2], T res[2]) {
  const T adj[3] = {X[0] - X[3] * ext[0], X[1] - X[3] * ext[1], X[2] - X[3] * ext[2]};
  const T sq = adj[0] * adj[0] + adj[1] * adj[1] + adj[2] * adj[2];
  if (sq < 1e-8) return false;
  T rot[3];
  AngleAxisRotatePoint(ext + 3, adj, rot);
  T pix[2];
  const bool ok = ProjectByModel(model, K, rot, pix);
  if (!ok) { res[0] = T(0.0); res[1] = T(0.0); return false; }
  res[0] = T(sqrt_info[0]) * (pix[0] - obs[0]);
  res[1] = T(sqrt_info[1]) * (pix[1] - obs[1]);
  return true;
}


// ---- inverse models (ORACLE): Model::PixelToCameraCoordinates + UndistortPoint, cited in csrc/camera_models.cuh ----------
inline bool PixelToCamera(int model, const double* K, const double px[2], double out[3]) {
  out[0] = out[1] = out[2] = 0.0;
  auto skewed_normalise = [&](double d[2]) {
    d[1] = (px[1] - K[4]) / (K[0] * K[1]);
    d[0] = (px[0] - K[3] - d[1] * K[2]) / K[0];
  };
  if (model == PINHOLE || model == FISHEYE) {
    double d[2]; skewed_normalise(d);
    double u[2] = {d[0], d[1]};
    for (int it = 0; it < 100; ++it) {
      const double prev[2] = {u[0], u[1]};
      const double r_sq = u[0] * u[0] + u[1] * u[1];
      if (model == PINHOLE) {
        const double f = 1.0 + r_sq * (K[5] + K[6] * r_sq);
        u[0] = d[0] / f; u[1] = d[1] / f;
      } else {
        const double r = std::sqrt(r_sq);
        if (r < 1e-8) { u[0] = d[0]; u[1] = d[1]; break; }
        const double th = std::atan2(r, 1.0), th2 = th * th;
        const double thd = th * (1.0 + K[5] * th2 + K[6] * th2 * th2 + K[7] * th2 * th2 * th2 + K[8] * th2 * th2 * th2 * th2);
        u[0] = r * d[0] / thd; u[1] = r * d[1] / thd;
      }
      if (std::fabs(u[0] - prev[0]) < 1e-10 && std::fabs(u[1] - prev[1]) < 1e-10) break;
    }
    out[0] = u[0]; out[1] = u[1]; out[2] = 1.0;
    return true;
  }
  if (model == FOV) {
    const double d[2] = {(px[0] - K[2]) / K[0], (px[1] - K[3]) / (K[0] * K[1])};
    const double w = K[4], rd2 = d[0] * d[0] + d[1] * d[1];
    double ru;
    if (w < 1e-3) ru = (w * w * rd2) / 3.0 - w * w / 12.0 + 1.0;
    else if (rd2 < 1e-3) ru = (w * (w * w * rd2 + 3.0)) / (6.0 * std::tan(w / 2.0));
    else { const double rd = std::sqrt(rd2); ru = std::tan(rd * w) / (2.0 * rd * std::tan(w / 2.0)); }
    out[0] = ru * d[0]; out[1] = ru * d[1]; out[2] = 1.0;
    return true;
  }
  if (model == DIVISION_UNDISTORTION) {
    const double d[2] = {px[0] - K[2], px[1] - K[3]};
    const double und = 1.0 / (1.0 + K[4] * (d[0] * d[0] + d[1] * d[1]));
    out[0] = d[0] * und / K[0]; out[1] = d[1] * und / (K[0] * K[1]); out[2] = 1.0;
    return true;
  }
  if (model == DOUBLE_SPHERE) {
    double d[2]; skewed_normalise(d);
    const double xi = K[5], al = K[6], r2 = d[0] * d[0] + d[1] * d[1];
    if (al > 0.5 && r2 >= 1.0 / (2.0 * al - 1.0)) return false;
    const double s2 = std::sqrt(1.0 - (2.0 * al - 1.0) * r2), n2 = al * s2 + 1.0 - al;
    const double mz = (1.0 - al * al * r2) / n2, mz2 = mz * mz;
    const double k = (mz * xi + std::sqrt(mz2 + (1.0 - xi * xi) * r2)) / (mz2 + r2);
    out[0] = k * d[0]; out[1] = k * d[1]; out[2] = k * mz - xi;
    return true;
  }
  if (model == EXTENDED_UNIFIED) {
    double d[2]; skewed_normalise(d);
    const double al = K[5], be = K[6], r2 = d[0] * d[0] + d[1] * d[1], ga = 1.0 - al;
    if (al > 0.5 && r2 >= 1.0 / ((al - ga) * be)) return false;
    const double k = (1.0 - al * al * be * r2) / (al * std::sqrt(1.0 - (al - ga) * be * r2) + ga);
    double nrm = std::sqrt(r2 + k * k);
    if (nrm < 1e-12) nrm = 1e-12;
    out[0] = d[0] / nrm; out[1] = d[1] / nrm; out[2] = k / nrm;
    return true;
  }
  return false;
}

}  // namespace oracle
#endif  // ORACLE_CAMERA_MODELS_H_
