// Device-side camera projections for the six camera models on the BA hot path.
//
// pixel = Model::CameraToPixelCoordinates(K, p) for a point p in the camera frame, written
// for T in {double, Dual<N>}; TK is the type of the intrinsics (double when they are held
// constant, Dual<N> when they are refined). Branches are taken on the scalar value so the
// derivative of the branch TAKEN is produced, which is what the reference's Jet autodiff
// yields. Behaviour follows (paths under /root/reference/src/theia/sfm/camera/):
//   pinhole_camera_model.h:182-211,244-260        double_sphere_camera_model.h:161-187,215-249
//   extended_unified_camera_model.h:161-187,215-249  fisheye_camera_model.h:163-189,227-272
//   fov_camera_model.h:156-181,209-258             division_undistortion_camera_model.h:173-203,263-297
#ifndef THB_CAMERA_MODELS_CUH_
#define THB_CAMERA_MODELS_CUH_

#include "common.cuh"

namespace thb {

THB_HD int num_intrinsics(int model) {
  switch (model) {
    case THB_MODEL_PINHOLE: return 7;
    case THB_MODEL_FISHEYE: return 9;
    case THB_MODEL_FOV: return 5;
    case THB_MODEL_DIVISION_UNDISTORTION: return 5;
    case THB_MODEL_DOUBLE_SPHERE: return 7;
    case THB_MODEL_EXTENDED_UNIFIED: return 7;
    default: return -1;
  }
}

// [f, a, s, cx, cy] tail shared by the models with skew.
template <typename TK, typename T>
THB_HD void apply_fasc(const TK* K, const T& dx, const T& dy, T pix[2]) {
  pix[0] = K[0] * dx + K[2] * dy + K[3];
  pix[1] = (K[0] * K[1]) * dy + K[4];
}

template <typename TK, typename T>
THB_HD bool project_pinhole(const TK* K, const T p[3], T pix[2]) {
  const T iz = 1.0 / p[2];
  const T nx = p[0] * iz, ny = p[1] * iz;
  const T r2 = nx * nx + ny * ny;
  const T d = 1.0 + r2 * (K[5] + K[6] * r2);
  apply_fasc(K, T(nx * d), T(ny * d), pix);
  return true;
}

template <typename TK, typename T>
THB_HD bool project_double_sphere(const TK* K, const T p[3], T pix[2]) {
  const TK xi = K[5], alpha = K[6];
  const T r2 = p[0] * p[0] + p[1] * p[1];
  const T d1 = d_sqrt(T(r2 + p[2] * p[2]));
  const TK one_m_alpha = 1.0 - alpha;
  const TK w1 = val(alpha) > 0.5 ? TK(one_m_alpha / alpha) : TK(alpha / one_m_alpha);
  const TK w2 = (w1 + xi) / d_sqrt(TK(2.0 * w1 * xi + xi * xi + 1.0));
  if (val(p[2]) <= -val(w2) * val(d1)) return false;  // outside the valid cone: evaluation fails
  const T k = xi * d1 + p[2];
  const T d2 = d_sqrt(T(r2 + k * k));
  const T inv = 1.0 / (alpha * d2 + one_m_alpha * k);
  apply_fasc(K, T(p[0] * inv), T(p[1] * inv), pix);
  return true;
}

template <typename TK, typename T>
THB_HD bool project_extended_unified(const TK* K, const T p[3], T pix[2]) {
  const TK alpha = K[5], beta = K[6];
  const T r2 = p[0] * p[0] + p[1] * p[1];
  const T rho = d_sqrt(T(beta * r2 + p[2] * p[2]));
  const T norm = alpha * rho + (1.0 - alpha) * p[2];
  bool zero = val(norm) < 1e-3;
  if (!zero && val(alpha) > 0.5) {
    const double zn = val(p[2]) / val(norm);
    const double Cc = (val(alpha) - 1.0) / (val(alpha) + val(alpha) - 1.0);
    zero = zn < Cc;
  }
  if (zero) {
    // distorted point is the constant (0,0): pixel = principal point, derivative only w.r.t. cx, cy
    apply_fasc(K, T(0.0), T(0.0), pix);
    return true;
  }
  const T inv = 1.0 / norm;
  apply_fasc(K, T(p[0] * inv), T(p[1] * inv), pix);
  return true;
}

template <typename TK, typename T>
THB_HD bool project_fisheye(const TK* K, const T p[3], T pix[2]) {
  const T r2 = p[0] * p[0] + p[1] * p[1];
  if (val(r2) < 1e-8) {
    apply_fasc(K, p[0], p[1], pix);
    return true;
  }
  const T r = d_sqrt(r2);
  const T theta = d_atan2(r, d_abs(p[2]));
  const T t2 = theta * theta;
  const T t4 = t2 * t2;
  const T theta_d = theta * (1.0 + K[5] * t2 + K[6] * t4 + K[7] * (t4 * t2) + K[8] * (t4 * t4));
  T s = theta_d / r;
  if (val(p[2]) < 0.0) s = -s;
  apply_fasc(K, T(p[0] * s), T(p[1] * s), pix);
  return true;
}

// K = [f, a, cx, cy, omega]
template <typename TK, typename T>
THB_HD bool project_fov(const TK* K, const T p[3], T pix[2]) {
  const T iz = 1.0 / p[2];
  const T nx = p[0] * iz, ny = p[1] * iz;
  const TK omega = K[4];
  const T r2 = nx * nx + ny * ny;
  T rd;
  if (val(omega) < 1e-3) {
    rd = (omega * omega) * r2 / 3.0 - (omega * omega) / 12.0 + 1.0;
  } else if (val(r2) < 1e-3) {
    const TK th = d_tan(TK(omega / 2.0));
    rd = (-2.0 * th * (4.0 * r2 * (th * th) - 3.0)) / (3.0 * omega);
  } else {
    const T ru = d_sqrt(r2);
    rd = d_atan(T(2.0 * ru * d_tan(TK(omega / 2.0)))) / (ru * omega);
  }
  pix[0] = K[0] * (rd * nx) + K[2];
  pix[1] = (K[0] * K[1]) * (rd * ny) + K[3];
  return true;
}

// K = [f, a, cx, cy, k]; focal length applied before the distortion.
template <typename TK, typename T>
THB_HD bool project_division(const TK* K, const T p[3], T pix[2]) {
  const T iz = 1.0 / p[2];
  const T ux = K[0] * (p[0] * iz), uy = (K[0] * K[1]) * (p[1] * iz);
  const T r2 = ux * ux + uy * uy;
  const TK k = K[4];
  const T denom = 2.0 * k * r2;
  const T inner = 1.0 - 4.0 * k * r2;
  if (fabs(val(denom)) < 2.220446049250313e-16 || val(inner) < 0.0) {
    pix[0] = ux + K[2];
    pix[1] = uy + K[3];
  } else {
    const T scale = (1.0 - d_sqrt(inner)) / denom;
    pix[0] = ux * scale + K[2];
    pix[1] = uy * scale + K[3];
  }
  return true;
}

// MODEL >= 0: compile-time model (single-model reconstructions, no switch in the kernel).
// MODEL == -1: run-time dispatch on `model` (mixed intrinsics groups).
template <int MODEL, typename TK, typename T>
THB_HD bool project(int model, const TK* K, const T p[3], T pix[2]) {
  if (MODEL >= 0) model = MODEL;
  switch (model) {
    case THB_MODEL_PINHOLE: return project_pinhole(K, p, pix);
    case THB_MODEL_FISHEYE: return project_fisheye(K, p, pix);
    case THB_MODEL_FOV: return project_fov(K, p, pix);
    case THB_MODEL_DIVISION_UNDISTORTION: return project_division(K, p, pix);
    case THB_MODEL_DOUBLE_SPHERE: return project_double_sphere(K, p, pix);
    case THB_MODEL_EXTENDED_UNIFIED: return project_extended_unified(K, p, pix);
    default: pix[0] = pix[1] = T(0.0); return false;
  }
}


// ---- inverse models: Model::PixelToCameraCoordinates (pixel -> point on the z = 1 plane, or a ray for the two omnidirectional
// models), what Camera::PixelToNormalizedCoordinates / PixelToUnitDepthRay call (camera.cc:218-230). Paths under
// /root/reference/src/theia/sfm/camera/: pinhole_camera_model.h:212-300, fisheye_camera_model.h:192-345,
// fov_camera_model.h:183-206,260-306, division_undistortion_camera_model.h:231-260,299-320,
// double_sphere_camera_model.h:188-212,251-286, extended_unified_camera_model.h:188-212,251-285. Returns false where the
// reference's UndistortPoint returns false (the reference then leaves the point uninitialised; here it is zero).
THB_HD bool pixel_to_camera(int model, const double* K, const double pixel[2], double pt[3]) {
  pt[0] = 0.0; pt[1] = 0.0; pt[2] = 0.0;
  switch (model) {
    case THB_MODEL_PINHOLE: {
      const double dy = (pixel[1] - K[4]) / (K[0] * K[1]);
      const double dx = (pixel[0] - K[3] - dy * K[2]) / K[0];
      double ux = dx, uy = dy;
      for (int i = 0; i < 100; ++i) {
        const double px = ux, py = uy;
        const double r_sq = ux * ux + uy * uy;
        const double d = 1.0 + r_sq * (K[5] + K[6] * r_sq);
        ux = dx / d; uy = dy / d;
        if (fabs(ux - px) < 1e-10 && fabs(uy - py) < 1e-10) break;
      }
      pt[0] = ux; pt[1] = uy; pt[2] = 1.0;
      return true;
    }
    case THB_MODEL_FISHEYE: {
      const double dy = (pixel[1] - K[4]) / (K[0] * K[1]);
      const double dx = (pixel[0] - K[3] - dy * K[2]) / K[0];
      double ux = dx, uy = dy;
      for (int i = 0; i < 100; ++i) {
        const double px = ux, py = uy;
        const double r = sqrt(ux * ux + uy * uy);
        if (r < 1e-8) { ux = dx; uy = dy; break; }
        const double theta = atan2(r, 1.0), t2 = theta * theta;
        const double theta_d = theta * (1.0 + K[5] * t2 + K[6] * t2 * t2 + K[7] * t2 * t2 * t2 + K[8] * t2 * t2 * t2 * t2);
        ux = r * dx / theta_d; uy = r * dy / theta_d;
        if (fabs(ux - px) < 1e-10 && fabs(uy - py) < 1e-10) break;
      }
      pt[0] = ux; pt[1] = uy; pt[2] = 1.0;
      return true;
    }
    case THB_MODEL_FOV: {
      const double dx = (pixel[0] - K[2]) / K[0], dy = (pixel[1] - K[3]) / (K[0] * K[1]);
      const double omega = K[4], r_d_sq = dx * dx + dy * dy;
      double r_u;
      if (omega < 1e-3) r_u = (omega * omega * r_d_sq) / 3.0 - omega * omega / 12.0 + 1.0;
      else if (r_d_sq < 1e-3) r_u = (omega * (omega * omega * r_d_sq + 3.0)) / (6.0 * tan(omega / 2.0));
      else { const double r_d = sqrt(r_d_sq); r_u = tan(r_d * omega) / (2.0 * r_d * tan(omega / 2.0)); }
      pt[0] = r_u * dx; pt[1] = r_u * dy; pt[2] = 1.0;
      return true;
    }
    case THB_MODEL_DIVISION_UNDISTORTION: {
      const double dx = pixel[0] - K[2], dy = pixel[1] - K[3];
      const double u = 1.0 / (1.0 + K[4] * (dx * dx + dy * dy));
      pt[0] = dx * u / K[0]; pt[1] = dy * u / (K[0] * K[1]); pt[2] = 1.0;
      return true;
    }
    case THB_MODEL_DOUBLE_SPHERE: {
      const double dy = (pixel[1] - K[4]) / (K[0] * K[1]);
      const double dx = (pixel[0] - K[3] - dy * K[2]) / K[0];
      const double xi = K[5], alpha = K[6], r2 = dx * dx + dy * dy;
      if (alpha > 0.5 && r2 >= 1.0 / (2.0 * alpha - 1.0)) return false;
      const double sqrt2 = sqrt(1.0 - (2.0 * alpha - 1.0) * r2);
      const double norm2 = alpha * sqrt2 + 1.0 - alpha;
      const double mz = (1.0 - alpha * alpha * r2) / norm2, mz2 = mz * mz;
      const double norm1 = mz2 + r2;
      const double sqrt1 = sqrt(mz2 + (1.0 - xi * xi) * r2);
      const double k = (mz * xi + sqrt1) / norm1;
      pt[0] = k * dx; pt[1] = k * dy; pt[2] = k * mz - xi;
      return true;
    }
    case THB_MODEL_EXTENDED_UNIFIED: {
      const double dy = (pixel[1] - K[4]) / (K[0] * K[1]);
      const double dx = (pixel[0] - K[3] - dy * K[2]) / K[0];
      const double alpha = K[5], beta = K[6], r2 = dx * dx + dy * dy, gamma = 1.0 - alpha;
      if (alpha > 0.5 && r2 >= 1.0 / ((alpha - gamma) * beta)) return false;
      const double tmp1 = 1.0 - alpha * alpha * beta * r2;
      const double tmp2 = alpha * sqrt(1.0 - (alpha - gamma) * beta * r2) + gamma;
      const double k = tmp1 / tmp2;
      double norm = sqrt(r2 + k * k);
      if (norm < 1e-12) norm = 1e-12;
      pt[0] = dx / norm; pt[1] = dy / norm; pt[2] = k / norm;
      return true;
    }
    default: return false;
  }
}

}  // namespace thb
#endif  // THB_CAMERA_MODELS_CUH_
