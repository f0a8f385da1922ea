// Device data model and the per-observation residual/Jacobian evaluation shared by the BA kernels.
//
// Layout in HBM (all FP64 / int32, structure-of-arrays):
//   parameters        cam[nc][6]   camd[nc][CAMD] (rotation matrix + SO(3) left Jacobian, derived)
//                     intr[ng][10] pts[np][4]
//   observations      stored twice, point-major (Schur order) and camera-major (camera-block
//                     order): cam/pt index, xy (double2), sqrt_info (double2)
//   Jacobian planes   r[2][no], Jc[12][no], Jp[2*PD][no], Ji[2*NK][no] — plane k of block B is
//                     the contiguous array B[k*no .. (k+1)*no): every warp store/load is one
//                     fully coalesced 256-byte segment.
#ifndef THB_BA_DEVICE_CUH_
#define THB_BA_DEVICE_CUH_

#include "camera_models.cuh"

namespace thb {

// Per-camera record gathered by every observation: ONE 128-byte, 128-byte-aligned L1 line read with four 256-bit loads
// (LDG.E.ENL2.256). Every lane of a point-major warp reads a different camera, so each load instruction costs 32 L1 tag
// lookups whatever its width: the r01 ncu capture showed K1 bound by exactly that (17 128-bit gathers of a 256-byte
// record, l1tex 77 % busy), and a 160-byte record straddles two lines (37 % L1 hit rate on the camera table). The
// rotation matrix and the SO(3) left Jacobian are therefore NOT stored (18 doubles) but applied from the angle-axis
// vector and four scalar coefficients:
//   R   = I + a [w]x + b [w]x^2        (ceres::AngleAxisRotatePoint; first-order branch: a = 1, b = 0)
//   J_l = I + A [w]x + B [w]x^2        (A = b, B = (th - sin th)/th^3; first-order branch: A = B = 0, the flag is A == 0)
//   w[3] | a b A B | C[3] | column scale[6]                                                      = 16 doubles
// Constant coordinates (SubsetManifold) carry column scale 0, which zeroes their Jacobian columns; th^2 = w.w is
// recomputed; the intrinsics group comes from BaConst::cam_group when there is more than one group.
constexpr int CAMD = 16;
constexpr int CD_W = 0, CD_A = 3, CD_B = 4, CD_JA = 5, CD_JB = 6, CD_C = 7, CD_SCALE = 10;

__device__ __forceinline__ void ld256(const double* p, double* o) {
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}

// Shared-memory copy of the camera table: record c keeps its 128 bytes, with its eight 16-byte pieces rotated by c.
// The lanes of a quarter-warp read the same piece k of eight different cameras in one LDS.128 phase: unrotated they would
// all sit on banks 4k..4k+3 (8-way conflict), rotated they collide only when two cameras agree mod 8.
__device__ __forceinline__ unsigned cam_smem_piece(unsigned base, int c, int k) { return base + c * (CAMD * 8) + (((k + c) & 7) << 4); }
__device__ __forceinline__ void lds128(unsigned addr, double* o) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// All threads of the CTA: record pieces go straight to their rotated place with cp.async (no registers, every copy in flight
// at once - a load / store loop here waited out one L2 latency per 8 KB: 14 % of the kernel's samples in the r01 capture).
// The caller commits, waits (cp_async_wait_all) and places a __syncthreads() before the first eval_obs.
__device__ __forceinline__ void stage_cam_table(const double* __restrict__ camd, int nc, unsigned base) {
  static_assert(CAMD == 16, "eight 16-byte pieces per record");
  const double2* src = reinterpret_cast<const double2*>(camd);
  for (int q = threadIdx.x; q < nc * 8; q += blockDim.x) cp_async16(cam_smem_piece(base, q >> 3, q & 7), src + q);
}

// y = (I + s*a [w]x + b [w]x^2) v with [w]x^2 = w w^T - th2 I; s = +1: the matrix, s = -1: its transpose.
__device__ __forceinline__ void rot_apply(const double w[3], double a, double b, double th2, const double v[3], double y[3]) {
  const double wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
  const double c0 = w[1] * v[2] - w[2] * v[1], c1 = w[2] * v[0] - w[0] * v[2], c2 = w[0] * v[1] - w[1] * v[0];
  y[0] = v[0] + a * c0 + b * (w[0] * wv - th2 * v[0]);
  y[1] = v[1] + a * c1 + b * (w[1] * wv - th2 * v[1]);
  y[2] = v[2] + a * c2 + b * (w[2] * wv - th2 * v[2]);
}

// Camera prior residual blocks (no loss): r[3] and the 3 x 6 ambient Jacobian, from the packed prior [sqrt information (9,
// row-major) | prior (3)] and the camera record (CAMD doubles: w, a, b, A, B, C).
//   position (position_error.h:44-80): r = A (prior - C),                 J = [-A | 0]
//   gravity  (gravity_error.h:44-86):  r = A (R(w) (0,0,-1) - prior),     J = [0 | A d(R g)/dw], d(R g)/dw = -[q]x J_l(w) with q = R g
//                                      (ceres' first-order branch: q = g, J_l = I), the chain eval_obs uses for the reprojection blocks
//   orientation (orientation_error.h:44-80): r = A log(exp(w) exp(prior)^-1) (Sophus SO3: unit quaternions), J = [0 | A J_l^-1(e) J_l(w)]
//                                      with e = the SO(3) logarithm: exp(w + d) = exp(J_l(w) d) exp(w) and log(exp(u) E) = e + J_l^-1(e) u
// Packed per camera: PRIOR_STRIDE doubles = 12 per kind (kind 0 position, 1 gravity, 2 orientation); has_prior bit = 1 << kind.
constexpr int PRIOR_STRIDE = 36, PRIOR_KINDS = 3;
__device__ __forceinline__ void cam_prior_position(const double* pr, const double* rec, double r[3], double J[3][6]);
__device__ __forceinline__ void cam_prior_gravity(const double* pr, const double* rec, double r[3], double J[3][6]);
__device__ __forceinline__ void cam_prior_orientation(const double* pr, const double* rec, double r[3], double J[3][6]);
// the prior block `kind` of a camera; `pr_cam` = the camera's PRIOR_STRIDE doubles
__device__ __forceinline__ void cam_prior_eval(int kind, const double* pr_cam, const double* rec, double r[3], double J[3][6]);

struct BaState {  // one set of parameter values (current x, or the candidate)
  double* cam;
  double* camd;
  double* intr;
  double* pts;
};

struct ObsSoA {  // observations in one ordering
  const int* cam;
  const int* pt;
  const double2* xy;
  const double2* si;
};

struct BaConst {  // problem structure, constant over the solve
  int nc, ng, np, no;
  const int* cam_group;     // [nc]
  const int* intr_model;    // [ng]
  const uint8_t* cam_const; // [nc] THB_CAM_CONST_* (never null on device)
  const uint16_t* intr_const; // [ng] bit k => constant
  const uint8_t* pt_const;  // [np]
  const int* intr_slot;     // [ng] index of the group among variable-intrinsics groups, or -1
  int loss_type;
  double loss_width;
};

// ceres::LossFunction::Evaluate for create_loss_function.cc:44-76 (formulas: ceres/loss_function.cc,
// external; TRUNCATED: loss_functions.cc:40-44). rho = {rho(s), rho'(s), rho''(s)}.
__device__ __forceinline__ void eval_loss(int type, double a, double s, double rho[3]) {
  const double b = a * a;
  switch (type) {
    case THB_LOSS_HUBER:
      if (s > b) {
        const double r = sqrt(s);
        rho[0] = 2.0 * a * r - b; rho[1] = fmax(2.2250738585072014e-308, a / r); rho[2] = -rho[1] / (2.0 * s);
      } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
      return;
    case THB_LOSS_SOFTLONE: {
      const double c = 1.0 / b, sum = 1.0 + s * c, tmp = sqrt(sum);
      rho[0] = 2.0 * b * (tmp - 1.0); rho[1] = fmax(2.2250738585072014e-308, 1.0 / tmp); rho[2] = -(c * rho[1]) / (2.0 * sum);
      return; }
    case THB_LOSS_CAUCHY: {
      const double c = 1.0 / b, sum = 1.0 + s * c, inv = 1.0 / sum;
      rho[0] = b * log(sum); rho[1] = fmax(2.2250738585072014e-308, inv); rho[2] = -c * (inv * inv);
      return; }
    case THB_LOSS_ARCTAN: {
      const double bb = 1.0 / b, sum = 1.0 + s * s * bb, inv = 1.0 / sum;
      rho[0] = a * atan2(s, a); rho[1] = fmax(2.2250738585072014e-308, inv); rho[2] = -2.0 * s * bb * (inv * inv);
      return; }
    case THB_LOSS_TUKEY:
      if (s <= b) {
        const double value = 1.0 - s / b, value_sq = value * value;
        rho[0] = b / 3.0 * (1.0 - value_sq * value); rho[1] = value_sq; rho[2] = -2.0 / b * value;
      } else { rho[0] = b / 3.0; rho[1] = 0.0; rho[2] = 0.0; }
      return;
    case THB_LOSS_TRUNCATED:
      rho[0] = fmin(s, b); rho[1] = s < b ? 1.0 : 0.0; rho[2] = 0.0;
      return;
    default:
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

// Householder vector of the homogeneous point (ceres SphereManifold<4>, bundle_adjuster.cc:538-545):
// H = I - beta v v^T maps x/|x| to e4, v[3] = 1.
__device__ __forceinline__ void householder4(const double x[4], double v[3], double* beta) {
  const double sigma = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  v[0] = x[0]; v[1] = x[1]; v[2] = x[2];
  *beta = 0.0;
  if (sigma <= 2.220446049250313e-16) {
    if (x[3] < 0.0) *beta = 2.0;
    return;
  }
  const double mu = sqrt(x[3] * x[3] + sigma);
  const double vp = (x[3] <= 0.0) ? (x[3] - mu) : (-sigma / (x[3] + mu));
  *beta = 2.0 * vp * vp / (sigma + vp * vp);
  const double ivp = 1.0 / vp;
  v[0] *= ivp; v[1] *= ivp; v[2] *= ivp;
}

// Residual only: ReprojectionError<Model>::operator() (reprojection_error.h:49-114).
// cam_rec != nullptr: the camera record is read from there (shared or global memory, plain loads) instead of S.camd[c] -
// the inner-iteration kernels evaluate a camera whose record changes inside the kernel.
template <int MODEL>
__device__ __forceinline__ bool eval_residual(const BaConst& K, const BaState& S, int c, int p, double2 xy,
                                              double2 si, double r[2], const double* cam_rec = nullptr) {
  double cd[12];
  if (cam_rec) {
#pragma unroll
    for (int k = 0; k < 12; ++k) cd[k] = cam_rec[k];
  } else {
    const double* rec = S.camd + (size_t)c * CAMD;
    ld256(rec, cd); ld256(rec + 4, cd + 4); ld256(rec + 8, cd + 8);
  }
  const double4 X = *reinterpret_cast<const double4*>(S.pts + (size_t)p * 4);
  const double adj[3] = {X.x - X.w * cd[CD_C], X.y - X.w * cd[CD_C + 1], X.z - X.w * cd[CD_C + 2]};
  if (adj[0] * adj[0] + adj[1] * adj[1] + adj[2] * adj[2] < 1e-8) return false;
  double pc[3];
  rot_apply(cd + CD_W, cd[CD_A], cd[CD_B], cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2], adj, pc);
  const int g = K.ng > 1 ? K.cam_group[c] : 0;
  double pix[2];
  if (!project<MODEL, double, double>(K.intr_model[g], S.intr + (size_t)g * KS, pc, pix)) return false;
  r[0] = si.x * (pix[0] - xy.x);
  r[1] = si.y * (pix[1] - xy.y);
  return true;
}

// Residual + tangent-space Jacobian blocks of one observation, robustified (ceres Corrector) and
// column-scaled (Jacobi scaling): what ProgramEvaluator hands to the linear solver.
//   r[2]; jc[2][6] (position | angle-axis); jp[2][PD]; ji[2][NK] (only NK > 0)
//   *half_rho = 0.5*rho(|r|^2) (cost contribution)
// cs/ps/is: column scales of the camera, point and intrinsics blocks (may be null => 1).
// ROBUST = false compiles the loss / Corrector code out (TRIVIAL loss, the reference default).
// CAM_SMEM (k_jacobian_sc): the camera record comes from the CTA's shared-memory copy of the table (cam_smem = its shared-space
// address, written by stage_cam_table), and the point, its column scales and its constness flag were fetched one round ahead:
// the point as two 16-byte pieces at pt_x and pt_x + 16 * pt_stride (cp.async slots of this thread), the scales in
// pt_scale[PD] (registers), the flag in pt_const_flag.
template <int MODEL, int PD, int NK, bool ROBUST = true, bool CAM_SMEM = false>
__device__ __forceinline__ bool eval_obs(const BaConst& K, const BaState& S, int c, int p, double2 xy, double2 si,
                                         const double* cs, const double* ps, const double* is, double r[2],
                                         double jc[12], double jp[2 * PD], double* ji, double* half_rho,
                                         unsigned cam_smem = 0, unsigned pt_x = 0, int pt_stride = 0, int pt_const_flag = 0,
                                         const double* pt_scale = nullptr, const double* cam_rec = nullptr) {
  constexpr int ND = 3 + NK;
  typedef Dual<ND> D;
  double cd[CAMD];
  if (CAM_SMEM) {
#pragma unroll
    for (int k = 0; k < CD_SCALE / 2; ++k) lds128(cam_smem_piece(cam_smem, c, k), cd + 2 * k);  // column scales: read where used
  } else if (cam_rec) {
#pragma unroll
    for (int k = 0; k < CAMD; ++k) cd[k] = cam_rec[k];
  } else {
    const double* rec = S.camd + (size_t)c * CAMD;
#pragma unroll
    for (int k = 0; k < CAMD / 4; ++k) ld256(rec + 4 * k, cd + 4 * k);
  }
  double X[4];
  if (CAM_SMEM) {
    lds128(pt_x, X); lds128(pt_x + 16 * pt_stride, X + 2);
  } else {
    const double4 X4 = *reinterpret_cast<const double4*>(S.pts + (size_t)p * 4);
    X[0] = X4.x; X[1] = X4.y; X[2] = X4.z; X[3] = X4.w;
  }
  const double C[3] = {cd[CD_C], cd[CD_C + 1], cd[CD_C + 2]};
  const double adj[3] = {X[0] - X[3] * C[0], X[1] - X[3] * C[1], X[2] - X[3] * C[2]};
  if (adj[0] * adj[0] + adj[1] * adj[1] + adj[2] * adj[2] < 1e-8) return false;
  const double* w = cd + CD_W;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double pc[3];
  rot_apply(w, cd[CD_A], cd[CD_B], th2, adj, pc);

  const int g = K.ng > 1 ? K.cam_group[c] : 0;
  const int model = MODEL >= 0 ? MODEL : K.intr_model[g];
  const double* Kp = S.intr + (size_t)g * KS;
  D pd[3], pix[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) pd[a] = seed<ND>(pc[a], a);
  bool ok;
  if (NK > 0) {
    D Kd[NK > 0 ? NK : 1];
#pragma unroll
    for (int k = 0; k < NK; ++k) Kd[k] = seed<ND>(k < KS ? Kp[k] : 0.0, 3 + k);
    ok = project<MODEL, D, D>(model, Kd, pd, pix);
  } else {
    ok = project<MODEL, double, D>(model, Kp, pd, pix);
  }
  if (!ok) return false;
  r[0] = si.x * (pix[0].a - xy.x);
  r[1] = si.y * (pix[1].a - xy.y);
  // A = sqrt_info * dpix/dp (2x3)
  double A[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) { A[k] = si.x * pix[0].v[k]; A[3 + k] = si.y * pix[1].v[k]; }
  // AR = A * R (2x3): d r / d adj; row a = R^T A_a
  double AR[6];
  rot_apply(w, -cd[CD_A], cd[CD_B], th2, A, AR);
  rot_apply(w, -cd[CD_A], cd[CD_B], th2, A + 3, AR + 3);
  // d r / d C = -w * AR (the columns of constant coordinates are zeroed by their column scale below)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) jc[6 * a + k] = -X[3] * AR[3 * a + k];
  // d r / d aa = A * (-[q]x M): q = R adj and M = left Jacobian of SO(3); in Ceres' small-angle
  // branch (R = I + [aa]x) q = adj and M = I.
  {
    const bool small = cd[CD_JA] == 0.0;  // first-order branch of k_cam_derive (A > 0 otherwise)
    const double q0 = small ? adj[0] : pc[0], q1 = small ? adj[1] : pc[1], q2 = small ? adj[2] : pc[2];
    // G = A * (-[q]x), with [q]x = [[0,-q2,q1],[q2,0,-q0],[-q1,q0,0]]; then row a of G * J_l = J_l^T G_a
    double G[6], GM[6];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const double a0 = A[3 * a], a1 = A[3 * a + 1], a2 = A[3 * a + 2];
      G[3 * a + 0] = -(a1 * q2 - a2 * q1);
      G[3 * a + 1] = -(a2 * q0 - a0 * q2);
      G[3 * a + 2] = -(a0 * q1 - a1 * q0);
    }
    rot_apply(w, -cd[CD_JA], cd[CD_JB], th2, G, GM);
    rot_apply(w, -cd[CD_JA], cd[CD_JB], th2, G + 3, GM + 3);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) jc[6 * a + 3 + k] = GM[3 * a + k];
  }
  // d r / d X (2x4) = [AR | -AR*C], then the tangent block
  const bool pconst = CAM_SMEM ? pt_const_flag != 0 : K.pt_const[p] != 0;
  if (PD == 3) {
    double v[3], beta;
    householder4(X, v, &beta);
    const double nx = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2] + X[3] * X[3]);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const double jw = -(AR[3 * a] * C[0] + AR[3 * a + 1] * C[1] + AR[3 * a + 2] * C[2]);
      const double jv = AR[3 * a] * v[0] + AR[3 * a + 1] * v[1] + AR[3 * a + 2] * v[2] + jw;  // J_X . v (v[3]=1)
#pragma unroll
      for (int k = 0; k < 3; ++k) jp[PD * a + k] = pconst ? 0.0 : nx * (AR[3 * a + k] - beta * v[k] * jv);
    }
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int k = 0; k < 3; ++k) jp[PD * a + k] = pconst ? 0.0 : AR[3 * a + k];
      jp[PD * a + 3] = pconst ? 0.0 : -(AR[3 * a] * C[0] + AR[3 * a + 1] * C[1] + AR[3 * a + 2] * C[2]);
    }
  }
  if (NK > 0) {
    const uint16_t icm = K.intr_const[g];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const bool kc = (icm >> k) & 1;
      ji[k] = kc ? 0.0 : si.x * pix[0].v[3 + k];
      ji[NK + k] = kc ? 0.0 : si.y * pix[1].v[3 + k];
    }
  }
  // loss: cost and Corrector (ceres/corrector.cc, external)
  const double sq = r[0] * r[0] + r[1] * r[1];
  double rs = 1.0, js = 1.0, asn = 0.0;
  if (ROBUST && K.loss_type != THB_LOSS_TRIVIAL) {
    double rho[3];
    eval_loss(K.loss_type, K.loss_width, sq, rho);
    *half_rho = 0.5 * rho[0];
    js = sqrt(rho[1]);
    if (sq == 0.0 || rho[2] <= 0.0) {
      rs = js;
    } else {
      const double Dd = 1.0 + 2.0 * sq * rho[2] / rho[1];
      const double alpha = 1.0 - sqrt(Dd);
      rs = js / (1.0 - alpha);
      asn = alpha / sq;
    }
  } else {
    *half_rho = 0.5 * sq;
  }
  // correct + scale columns
  if (CAM_SMEM && cs) {
#pragma unroll
    for (int k = CD_SCALE / 2; k < CAMD / 2; ++k) lds128(cam_smem_piece(cam_smem, c, k), cd + 2 * k);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double j0 = jc[k], j1 = jc[6 + k];
    if (asn != 0.0) { const double rtj = j0 * r[0] + j1 * r[1]; j0 -= asn * r[0] * rtj; j1 -= asn * r[1] * rtj; }
    const double s = js * (cs ? cd[CD_SCALE + k] : 1.0);
    jc[k] = j0 * s; jc[6 + k] = j1 * s;
  }
#pragma unroll
  for (int k = 0; k < PD; ++k) {
    double j0 = jp[k], j1 = jp[PD + k];
    if (asn != 0.0) { const double rtj = j0 * r[0] + j1 * r[1]; j0 -= asn * r[0] * rtj; j1 -= asn * r[1] * rtj; }
    const double s = js * (ps ? (CAM_SMEM ? pt_scale[k] : ps[(size_t)p * PD + k]) : 1.0);
    jp[k] = j0 * s; jp[PD + k] = j1 * s;
  }
  if (NK > 0) {
    const int slot = K.intr_slot[g];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      double j0 = ji[k], j1 = ji[NK + k];
      if (asn != 0.0) { const double rtj = j0 * r[0] + j1 * r[1]; j0 -= asn * r[0] * rtj; j1 -= asn * r[1] * rtj; }
      const double s = js * ((is && slot >= 0) ? is[(size_t)slot * NK + k] : 1.0);
      ji[k] = slot >= 0 ? j0 * s : 0.0; ji[NK + k] = slot >= 0 ? j1 * s : 0.0;
    }
  }
  r[0] *= rs; r[1] *= rs;
  return true;
}

// Per-camera record from the 6 extrinsics (used by k_cam_derive and by kernels that move a camera inside the kernel)
__device__ __forceinline__ void cam_derive_record(const double* __restrict__ cam6, double* __restrict__ o) {
  const double wx = cam6[3], wy = cam6[4], wz = cam6[5];
  const double th2 = wx * wx + wy * wy + wz * wz;
  o[CD_W] = wx; o[CD_W + 1] = wy; o[CD_W + 2] = wz;
  if (th2 > 2.220446049250313e-16) {
    const double th = sqrt(th2);
    double s, co;
    sincos(th, &s, &co);
    // R = I + (sin th / th) [w]x + ((1 - cos th) / th^2) [w]x^2; J_l = I + A [w]x + B [w]x^2,
    // A = (1 - cos th)/th^2, B = (th - sin th)/th^3 (series below 1e-2: the closed forms cancel)
    double A, B;
    if (th < 1e-2) {
      A = 0.5 - th2 / 24.0 + th2 * th2 / 720.0;
      B = 1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0;
    } else {
      const double sh = sin(0.5 * th);
      A = 2.0 * sh * sh / th2;
      B = (th - s) / (th2 * th);
    }
    o[CD_A] = s / th; o[CD_B] = A; o[CD_JA] = A; o[CD_JB] = B;
  } else {
    o[CD_A] = 1.0; o[CD_B] = 0.0; o[CD_JA] = 0.0; o[CD_JB] = 0.0;
  }
  o[CD_C] = cam6[0]; o[CD_C + 1] = cam6[1]; o[CD_C + 2] = cam6[2];
}

// Small SPD inverse by Cholesky, N in {3,4}; returns false if not positive definite.
__device__ __forceinline__ void cam_prior_position(const double* pr, const double* rec, double r[3], double J[3][6]) {
  const double d[3] = {pr[9] - rec[CD_C], pr[10] - rec[CD_C + 1], pr[11] - rec[CD_C + 2]};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r[k] = pr[3 * k] * d[0] + pr[3 * k + 1] * d[1] + pr[3 * k + 2] * d[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) { J[k][a] = -pr[3 * k + a]; J[k][3 + a] = 0.0; }
  }
}
__device__ __forceinline__ void cam_prior_gravity(const double* pr, const double* rec, double r[3], double J[3][6]) {
  const double* w = rec + CD_W;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double g[3] = {0.0, 0.0, -1.0};
  double q[3];
  rot_apply(w, rec[CD_A], rec[CD_B], th2, g, q);
  const bool small = rec[CD_JA] == 0.0;
  const double q0 = small ? g[0] : q[0], q1 = small ? g[1] : q[1], q2 = small ? g[2] : q[2];
  const double d[3] = {q[0] - pr[9], q[1] - pr[10], q[2] - pr[11]};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double a0 = pr[3 * k], a1 = pr[3 * k + 1], a2 = pr[3 * k + 2];
    r[k] = a0 * d[0] + a1 * d[1] + a2 * d[2];
    const double G[3] = {-(a1 * q2 - a2 * q1), -(a2 * q0 - a0 * q2), -(a0 * q1 - a1 * q0)};  // A_k (-[q]x)
    double GM[3];
    rot_apply(w, -rec[CD_JA], rec[CD_JB], th2, G, GM);                                         // ... J_l
    J[k][0] = J[k][1] = J[k][2] = 0.0;
    J[k][3] = GM[0]; J[k][4] = GM[1]; J[k][5] = GM[2];
  }
}

// Sophus::SO3::exp (so3.hpp): unit quaternion (w, x, y, z) of a rotation vector, Taylor branch below |w|^2 < 1e-20
__device__ __forceinline__ void so3_exp_quat(const double w[3], double q[4]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (th2 < 1e-10 * 1e-10) {
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    const double th = sqrt(th2), half = 0.5 * th;
    imag = sin(half) / th;
    real = cos(half);
  }
  q[0] = real; q[1] = imag * w[0]; q[2] = imag * w[1]; q[3] = imag * w[2];
}
__device__ __forceinline__ void cam_prior_orientation(const double* pr, const double* rec, double r[3], double J[3][6]) {
  const double* w = rec + CD_W;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double qa[4], qb[4];
  so3_exp_quat(w, qa);
  so3_exp_quat(pr + 9, qb);
  qb[1] = -qb[1]; qb[2] = -qb[2]; qb[3] = -qb[3];  // inverse of a unit quaternion
  double q[4] = {qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2] - qa[3] * qb[3],
                 qa[0] * qb[1] + qa[1] * qb[0] + qa[2] * qb[3] - qa[3] * qb[2],
                 qa[0] * qb[2] - qa[1] * qb[3] + qa[2] * qb[0] + qa[3] * qb[1],
                 qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1] + qa[3] * qb[0]};
  const double n2q = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2q != 1.0) { const double sc = 2.0 / (1.0 + n2q); q[0] *= sc; q[1] *= sc; q[2] *= sc; q[3] *= sc; }  // SO3::operator* renormalisation
  // SO3::log: e = 2 atan(|v| / w) v / |v|
  const double n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double f;
  if (n2 < 1e-10 * 1e-10) {
    f = 2.0 / q[0] - (2.0 / 3.0) * n2 / (q[0] * q[0] * q[0]);
  } else {
    const double n = sqrt(n2);
    const double at = q[0] < 0.0 ? atan2(-n, -q[0]) : atan2(n, q[0]);
    f = 2.0 * at / n;
  }
  const double e[3] = {f * q[1], f * q[2], f * q[3]};
  const double ph2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  // J_l^-1(e) = I - [e]x / 2 + c [e]x^2, c = 1 / ph^2 - (1 + cos ph) / (2 ph sin ph) (series below 1e-2: the closed form cancels)
  double c;
  if (ph2 < 1e-4) {
    c = 1.0 / 12.0 + ph2 / 720.0 + ph2 * ph2 / 30240.0;
  } else {
    const double ph = sqrt(ph2);
    double sn, cs;
    sincos(ph, &sn, &cs);
    c = 1.0 / ph2 - (1.0 + cs) / (2.0 * ph * sn);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double g[3] = {pr[3 * k], pr[3 * k + 1], pr[3 * k + 2]};
    r[k] = g[0] * e[0] + g[1] * e[1] + g[2] * e[2];
    double t1[3], t2[3];
    rot_apply(e, 0.5, c, ph2, g, t1);                       // (J_l^-1(e))^T A_k^T
    rot_apply(w, -rec[CD_JA], rec[CD_JB], th2, t1, t2);     // J_l(w)^T ...
    J[k][0] = J[k][1] = J[k][2] = 0.0;
    J[k][3] = t2[0]; J[k][4] = t2[1]; J[k][5] = t2[2];
  }
}
__device__ __forceinline__ void cam_prior_eval(int kind, const double* pr_cam, const double* rec, double r[3], double J[3][6]) {
  if (kind == 0) cam_prior_position(pr_cam, rec, r, J);
  else if (kind == 1) cam_prior_gravity(pr_cam + 12, rec, r, J);
  else cam_prior_orientation(pr_cam + 24, rec, r, J);
}

template <int N>
__device__ __forceinline__ bool spd_inverse(const double* A /*row-major NxN, lower used*/, double* Ainv) {
  double L[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double Li[N][N];  // inverse of L (lower)
#pragma unroll
  for (int j = 0; j < N; ++j) {
    Li[j][j] = 1.0 / L[j][j];
#pragma unroll
    for (int r = j + 1; r < N; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int k = j; k < r; ++k) acc += L[r][k] * Li[k][j];
      Li[r][j] = -acc / L[r][r];
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = i; k < N; ++k) s += Li[k][i] * Li[k][j];
      Ainv[i * N + j] = s; Ainv[j * N + i] = s;
    }
  return true;
}


// Solve the SPD system A y = b (A row-major N x N, lower triangle read) by Cholesky; false if not positive definite.
template <int N>
__device__ __host__ inline bool spd_solve(const double* A, const double* b, double* y) {
  double L[N * N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = A[i * N + j];
      for (int k = 0; k < j; ++k) v -= L[i * N + k] * L[j * N + k];
      if (i == j) { if (!(v > 0.0)) return false; L[i * N + i] = sqrt(v); }
      else L[i * N + j] = v / L[j * N + j];
    }
  double z[N];
  for (int i = 0; i < N; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[i * N + k] * z[k]; z[i] = v / L[i * N + i]; }
  for (int i = N - 1; i >= 0; --i) { double v = z[i]; for (int k = i + 1; k < N; ++k) v -= L[k * N + i] * y[k]; y[i] = v / L[i * N + i]; }
  return true;
}


}  // namespace thb
#endif  // THB_BA_DEVICE_CUH_
