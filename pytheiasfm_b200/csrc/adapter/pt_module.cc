// Host-side mirror of the reference interface for the two hot paths, in C++ behind pybind11 (the reference's host
// language): the same free functions, argument orders and return types that src/pytheia/sfm/sfm.cc binds —
//   pt.sfm.BundleAdjustReconstruction(opts, recon)                sfm.cc:1605-1621, bundle_adjustment_wrapper.cc:98-103
//   pt.sfm.BundleAdjustPartialReconstruction(opts, views, tracks, recon)
//   pt.sfm.BundleAdjustView(recon, opts, view_id) / BundleAdjustViews      bundle_adjustment_wrapper.cc:14-40
//   pt.sfm.BundleAdjustTrack(recon, opts, track_id) / BundleAdjustTracks
//   pt.sfm.EstimateRelativePose / EstimateCalibratedAbsolutePose / EstimateHomography   sfm.cc:825-850, estimators_wrapper.cc
//   pt.sfm.FivePointRelativePose / FourPointHomography / SevenPointFundamentalMatrix / PoseFromThreePoints  sfm.cc:577-597
// implemented as gather -> C-ABI (include/theia_b200.h) -> scatter over a minimal mirror of
// Reconstruction / View / Track / Camera / Feature (sfm/reconstruction.h, view.h, track.h, camera/camera.h, feature.h)
// with the reference's method names. No Eigen: vectors and matrices cross the binding as numpy arrays.
// There is no CPU implementation here: every compute call goes to libtheia_b200.so and fails loudly without a B200.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/theia_b200.h"
#include "../camera_models.cuh"  // plain C++ under g++ (THB_HD = inline): Camera::ProjectPoint uses the same camera math

namespace py = pybind11;

namespace {

using Vec = py::array_t<double, py::array::c_style | py::array::forcecast>;
typedef uint32_t ViewId;
typedef uint32_t TrackId;
typedef uint32_t GroupId;
constexpr uint32_t kInvalid = std::numeric_limits<uint32_t>::max();  // types.h:48-56

Vec MakeVec(const double* d, int n) {
  Vec v(n);
  std::memcpy(v.mutable_data(), d, sizeof(double) * n);
  return v;
}
Vec MakeMat(const double* d, int r, int c) {
  Vec v({r, c});
  std::memcpy(v.mutable_data(), d, sizeof(double) * r * c);
  return v;
}
void CopyVec(const Vec& v, double* out, int n, const char* what) {
  if (v.size() != n) throw std::invalid_argument(std::string(what) + ": expected " + std::to_string(n) + " values");
  std::memcpy(out, v.data(), sizeof(double) * n);
}
void Check(int rc) {
  if (rc != THB_OK) throw std::runtime_error(std::string("theia_b200: ") + thb_last_error() + " (code " + std::to_string(rc) + ")");
}

// ceres::AngleAxisRotatePoint / AngleAxisToRotationMatrix semantics (external)
void AngleAxisToRotation(const double* aa, double* R) {
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (th2 > std::numeric_limits<double>::epsilon()) {
    const double th = std::sqrt(th2), c = std::cos(th), s = std::sin(th), oc = 1.0 - c;
    const double x = aa[0] / th, y = aa[1] / th, z = aa[2] / th;
    R[0] = c + oc * x * x; R[1] = oc * x * y - s * z; R[2] = oc * x * z + s * y;
    R[3] = oc * x * y + s * z; R[4] = c + oc * y * y; R[5] = oc * y * z - s * x;
    R[6] = oc * x * z - s * y; R[7] = oc * y * z + s * x; R[8] = c + oc * z * z;
  } else {
    R[0] = 1; R[1] = -aa[2]; R[2] = aa[1]; R[3] = aa[2]; R[4] = 1; R[5] = -aa[0]; R[6] = -aa[1]; R[7] = aa[0]; R[8] = 1;
  }
}
void RotationToAngleAxis(const double* R, double* aa) {  // via the quaternion, like ceres::RotationMatrixToAngleAxis
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr >= 0.0) {
    double t = std::sqrt(tr + 1.0);
    q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    q[i + 1] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j + 1] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k + 1] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = std::sqrt(s2);
    const double two_theta = 2.0 * (q[0] < 0.0 ? std::atan2(-s, -q[0]) : std::atan2(s, q[0]));
    const double k = two_theta / s;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
  } else {
    aa[0] = q[1] * 2.0; aa[1] = q[2] * 2.0; aa[2] = q[3] * 2.0;
  }
}

// feature.h:47-112 (point_, covariance_; depth prior unused on this path)
struct Feature {
  double point[2] = {0, 0};
  double cov[4] = {1, 0, 0, 1};
};

// camera_intrinsics_model.h + the per-model parameter layouts; shared between views of one group
struct Intrinsics {
  int model = THB_MODEL_PINHOLE;
  double params[THB_INTR_STRIDE] = {1, 1, 0, 0, 0, 0, 0, 0, 0, 0};
  void SetType(int m) {
    if (thb::num_intrinsics(m) < 0) throw std::invalid_argument("camera model not on the hot path");
    model = m;
    std::fill(params, params + THB_INTR_STRIDE, 0.0);
    params[0] = 1.0; params[1] = 1.0;                        // f = 1, aspect = 1 (all models)
    if (m == THB_MODEL_FOV) params[4] = 0.75;                 // fov_camera_model.cc:61
    if (m == THB_MODEL_DOUBLE_SPHERE) params[6] = 0.75;       // alpha, double_sphere_camera_model.cc:57-64
    if (m == THB_MODEL_EXTENDED_UNIFIED) { params[5] = 0.5; params[6] = 1.0; }
  }
  bool has_skew() const { return model != THB_MODEL_FOV && model != THB_MODEL_DIVISION_UNDISTORTION; }
  int cx_index() const { return has_skew() ? 3 : 2; }
};

// camera/camera.h: extrinsics [C(3), angle-axis(3)] (camera.h:202-204), shared_ptr intrinsics
struct Camera {
  double ext[6] = {0, 0, 0, 0, 0, 0};
  std::shared_ptr<Intrinsics> intr = std::make_shared<Intrinsics>();
  int width = 0, height = 0;

  void SetPosition(const Vec& p) { CopyVec(p, ext, 3, "position"); }
  Vec GetPosition() const { return MakeVec(ext, 3); }
  void SetOrientationFromAngleAxis(const Vec& a) { CopyVec(a, ext + 3, 3, "angle_axis"); }
  Vec GetOrientationAsAngleAxis() const { return MakeVec(ext + 3, 3); }
  Vec GetOrientationAsRotationMatrix() const { double R[9]; AngleAxisToRotation(ext + 3, R); return MakeMat(R, 3, 3); }
  void SetOrientationFromRotationMatrix(const Vec& Rm) { double R[9]; CopyVec(Rm, R, 9, "rotation"); RotationToAngleAxis(R, ext + 3); }
  void SetFocalLength(double f) { intr->params[0] = f; }
  double FocalLength() const { return intr->params[0]; }
  void SetPrincipalPoint(double cx, double cy) { intr->params[intr->cx_index()] = cx; intr->params[intr->cx_index() + 1] = cy; }
  double PrincipalPointX() const { return intr->params[intr->cx_index()]; }
  double PrincipalPointY() const { return intr->params[intr->cx_index() + 1]; }
  void SetImageSize(int w, int h) { width = w; height = h; }
  void SetCameraIntrinsicsModelType(int m) { intr->SetType(m); }
  int GetCameraIntrinsicsModelType() const { return intr->model; }
  Vec Parameters() const { return MakeVec(intr->params, thb::num_intrinsics(intr->model)); }
  void SetParameters(const Vec& v) { CopyVec(v, intr->params, thb::num_intrinsics(intr->model), "intrinsics"); }
  void DeepCopy(const Camera& o) {  // camera.cc: copies extrinsics and a private copy of the intrinsics
    std::copy(o.ext, o.ext + 6, ext);
    intr = std::make_shared<Intrinsics>(*o.intr);
    width = o.width; height = o.height;
  }
  // Camera::ProjectPoint (camera.cc:206-216): returns (depth, pixel)
  double ProjectRaw(const double* X4, double* pixel) const {
    double R[9];
    AngleAxisToRotation(ext + 3, R);
    const double a[3] = {X4[0] - X4[3] * ext[0], X4[1] - X4[3] * ext[1], X4[2] - X4[3] * ext[2]};
    double p[3];
    for (int r = 0; r < 3; ++r) p[r] = R[3 * r] * a[0] + R[3 * r + 1] * a[1] + R[3 * r + 2] * a[2];
    if (pixel) thb::project<-1, double, double>(intr->model, intr->params, p, pixel);
    return p[2] / X4[3];
  }
  py::tuple ProjectPoint(const Vec& X) const {
    double x[4], pix[2] = {0, 0};
    CopyVec(X, x, 4, "point");
    const double depth = ProjectRaw(x, pix);
    return py::make_tuple(depth, MakeVec(pix, 2));
  }
};

// track.h
struct Track {
  double point[4] = {0, 0, 0, 1};
  bool estimated = false;
  double inverse_depth = 0.0;
  ViewId reference_view = kInvalid;
  std::vector<ViewId> views;  // insertion order; the first view added is the reference view (track.cc:78-83)
  mutable int flat_index = -1;          // slot of this track in the gather of the current BundleAdjust* call ...
  mutable uint64_t flat_epoch = 0;      // ... valid when it carries that call's epoch (saves a hash look-up per observation)
  void SetPoint(const Vec& p) { CopyVec(p, point, 4, "point"); }
  Vec Point() const { return MakeVec(point, 4); }
};

// view.h
struct View {
  std::string name;
  bool estimated = false;
  Camera camera;
  std::unordered_map<TrackId, Feature> features;
  std::vector<TrackId> track_order;
  bool has_position_prior = false;                       // view.h:101-106
  double position_prior[3] = {0, 0, 0};
  double position_prior_sqrt_info[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  bool has_gravity_prior = false;                        // view.h: gravity direction in the camera frame (IMU)
  double gravity_prior[3] = {0, 0, -1};
  double gravity_prior_sqrt_info[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  bool has_orientation_prior = false;                    // view.h:95-99: angle-axis orientation prior
  double orientation_prior[3] = {0, 0, 0};
  double orientation_prior_sqrt_info[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::vector<TrackId> TrackIds() const { return track_order; }
  const Feature* GetFeature(TrackId t) const { auto it = features.find(t); return it == features.end() ? nullptr : &it->second; }
};

// reconstruction.h:196-206 (hash maps of views and tracks, intrinsics groups)
struct Reconstruction {
  std::unordered_map<ViewId, View> views;
  std::unordered_map<TrackId, Track> tracks;
  std::unordered_map<ViewId, GroupId> view_group;
  std::unordered_map<GroupId, std::vector<ViewId>> group_views;
  std::unordered_map<std::string, ViewId> name_to_view;
  std::vector<ViewId> view_order;
  std::vector<TrackId> track_order;
  ViewId next_view = 0; TrackId next_track = 0; GroupId next_group = 0;

  ViewId AddView(const std::string& name, GroupId group, double /*timestamp*/) {  // reconstruction.cc:104-146
    if (name_to_view.count(name)) return kInvalid;
    const ViewId id = next_view++;
    View& v = views[id];
    v.name = name;
    name_to_view[name] = id;
    view_order.push_back(id);
    auto& members = group_views[group];
    if (!members.empty()) v.camera.intr = views[members.front()].camera.intr;  // shared intrinsics (reconstruction.cc:129-140)
    members.push_back(id);
    view_group[id] = group;
    next_group = std::max(next_group, group + 1);
    return id;
  }
  ViewId AddViewAutoGroup(const std::string& name, double ts) { return AddView(name, next_group, ts); }
  TrackId AddTrack() { const TrackId id = next_track++; tracks[id]; track_order.push_back(id); return id; }
  bool AddObservation(ViewId v, TrackId t, const Feature& f) {  // reconstruction.cc:221-251
    auto vi = views.find(v); auto ti = tracks.find(t);
    if (vi == views.end() || ti == tracks.end() || vi->second.features.count(t)) return false;
    vi->second.features[t] = f;
    vi->second.track_order.push_back(t);
    ti->second.views.push_back(v);
    if (ti->second.reference_view == kInvalid) ti->second.reference_view = v;
    return true;
  }
  View* MutableView(ViewId v) { auto it = views.find(v); return it == views.end() ? nullptr : &it->second; }
  Track* MutableTrack(TrackId t) { auto it = tracks.find(t); return it == tracks.end() ? nullptr : &it->second; }
  GroupId CameraIntrinsicsGroupIdFromViewId(ViewId v) const { auto it = view_group.find(v); return it == view_group.end() ? kInvalid : it->second; }
};

// bundle_adjustment.h:71-85
enum OptimizeIntrinsicsType { NONE = 0x00, FOCAL_LENGTH = 0x01, ASPECT_RATIO = 0x02, SKEW = 0x04, PRINCIPAL_POINTS = 0x08,
                              RADIAL_DISTORTION = 0x10, TANGENTIAL_DISTORTION = 0x20, ALL = 0x3f };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY = 0, DENSE_QR = 1, SPARSE_NORMAL_CHOLESKY = 2, DENSE_SCHUR = 3, SPARSE_SCHUR = 4,
                        ITERATIVE_SCHUR = 5, CGNR = 6 };  // ceres::LinearSolverType values

// GetSubsetFromOptimizeIntrinsicsType (e.g. pinhole_camera_model.cc:132-162): bit k set => parameter k CONSTANT
uint16_t ConstantIntrinsicsMask(const Intrinsics& in, int to_optimize) {
  const int K = thb::num_intrinsics(in.model);
  if (to_optimize == ALL) return 0;
  uint16_t m = 0;
  if (!(to_optimize & FOCAL_LENGTH)) m |= 1u << 0;
  if (!(to_optimize & ASPECT_RATIO)) m |= 1u << 1;
  int k = 2;
  if (in.has_skew()) { if (!(to_optimize & SKEW)) m |= 1u << 2; k = 3; }
  if (!(to_optimize & PRINCIPAL_POINTS)) m |= (1u << k) | (1u << (k + 1));
  if (!(to_optimize & RADIAL_DISTORTION)) for (int i = k + 2; i < K; ++i) m |= 1u << i;
  return m;
}

// bundle_adjustment.h:87-167
struct BundleAdjustmentOptions {
  int loss_function_type = THB_LOSS_TRIVIAL;
  double robust_loss_width = 2.0;
  int linear_solver_type = SPARSE_SCHUR;
  bool verbose = false, constant_camera_orientation = false, constant_camera_position = false;
  bool use_homogeneous_point_parametrization = true, use_inverse_depth_parametrization = false;
  int intrinsics_to_optimize = NONE;
  int num_threads = 1, max_num_iterations = 100;
  double max_solver_time_in_seconds = 3600.0;
  bool use_inner_iterations = true;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8, max_trust_region_radius = 1e12;
  bool use_position_priors = false, use_orientation_priors = false, use_depth_priors = false, orthographic_camera = false,
       use_gravity_priors = false;
};
struct BundleAdjustmentSummary {  // bundle_adjustment.h:170-178
  bool success = false;
  double initial_cost = 0, final_cost = 0, setup_time_in_seconds = 0, solve_time_in_seconds = 0;
};

struct Flat {
  std::vector<ViewId> view_ids; std::unordered_map<ViewId, int> view_index;
  std::vector<const Track*> track_ptrs;  // the Track objects behind track_ids (scatter without hash lookups)
  std::vector<TrackId> track_ids; std::unordered_map<TrackId, int> track_index;
  std::vector<Intrinsics*> groups; std::unordered_map<Intrinsics*, int> group_index;
  std::vector<double> cam_ext, intr, pts, obs_xy, obs_si;
  std::vector<uint8_t> cam_const, pt_const; std::vector<uint16_t> intr_const;
  std::vector<int32_t> cam_group, intr_model, obs_cam, obs_pt;
  std::vector<uint8_t> has_prior; std::vector<double> prior, prior_sqrt_info;   // position priors of the AddView cameras
  std::vector<uint8_t> has_gprior; std::vector<double> gprior, gprior_sqrt_info;  // gravity priors
  std::vector<uint8_t> has_oprior; std::vector<double> oprior, oprior_sqrt_info;  // orientation priors
  bool any_prior = false, any_gprior = false, any_oprior = false;
};

// SetSolverOptions, bundle_adjuster.cc:63-89
ThbBaOptions MapOptions(const BundleAdjustmentOptions& o, bool force_no_inner) {
  ThbBaOptions opt;
  thb_ba_default_options(&opt);
  opt.loss_function_type = o.loss_function_type; opt.robust_loss_width = o.robust_loss_width;
  opt.use_homogeneous_point_parametrization = o.use_homogeneous_point_parametrization ? 1 : 0;
  opt.use_inner_iterations = (!force_no_inner && o.use_inner_iterations) ? 1 : 0;
  opt.max_num_iterations = o.max_num_iterations;
  opt.function_tolerance = o.function_tolerance; opt.gradient_tolerance = o.gradient_tolerance;
  opt.parameter_tolerance = o.parameter_tolerance; opt.max_trust_region_radius = o.max_trust_region_radius;
  opt.max_solver_time_in_seconds = o.max_solver_time_in_seconds; opt.verbose = o.verbose;
  opt.linear_solver = THB_SOLVER_SCHUR_CHOLESKY;                           // every exact ceres solver type maps here
  if (o.linear_solver_type == ITERATIVE_SCHUR) opt.linear_solver = THB_SOLVER_SCHUR_PCG;  // SCHUR_JACOBI, ceres' default eta / iteration cap
  if (o.linear_solver_type == CGNR) throw std::runtime_error("CGNR is not implemented; use ITERATIVE_SCHUR or an exact Schur solver");
  return opt;
}

// What BundleAdjuster::AddView / AddTrack register (bundle_adjuster.cc:116-221), flattened.
// covariance blocks of the BundleAdjust{View,Track}(s) overloads (bundle_adjustment.cc:287-380,419-500), row-major
struct CovOut {
  std::map<ViewId, std::array<double, 36>> views;
  std::map<TrackId, std::array<double, 9>> tracks;
  bool ok = false;  // GetCovarianceFor* succeeded for every requested block
};

BundleAdjustmentSummary RunBa(const BundleAdjustmentOptions& o, const std::vector<ViewId>& views, const std::vector<TrackId>& tracks,
                              Reconstruction* r, bool force_no_inner, CovOut* cov = nullptr) {
  if (o.use_inverse_depth_parametrization) throw std::runtime_error("use_inverse_depth_parametrization is not implemented");
  if (o.use_depth_priors || o.orthographic_camera)
    throw std::runtime_error("depth prior residuals and orthographic cameras are not implemented (position, gravity and orientation priors are)");
  Flat f;
  const bool prof = getenv("THB_ADAPTER_PROF") != nullptr;   // host time stamps of gather / solve / scatter on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  double t_gather = 0.0, t_solve = 0.0;
  const std::unordered_set<ViewId> vset(views.begin(), views.end());
  auto add_view = [&](ViewId v, bool free_cam) {
    auto it = f.view_index.find(v);
    if (it != f.view_index.end()) return it->second;
    View& view = r->views.at(v);
    const int idx = (int)f.view_ids.size();
    f.view_index[v] = idx; f.view_ids.push_back(v);
    f.cam_ext.insert(f.cam_ext.end(), view.camera.ext, view.camera.ext + 6);
    uint8_t c = free_cam ? 0 : THB_CAM_CONST_ALL;
    if (o.constant_camera_position) c |= THB_CAM_CONST_POSITION;          // bundle_adjuster.cc:357-380
    if (o.constant_camera_orientation) c |= THB_CAM_CONST_ORIENTATION;
    f.cam_const.push_back(c);
    const bool prior = free_cam && o.use_position_priors && view.has_position_prior;    // AddView only (bundle_adjuster.cc:160-163)
    f.has_prior.push_back(prior ? 1 : 0);
    f.prior.insert(f.prior.end(), view.position_prior, view.position_prior + 3);
    f.prior_sqrt_info.insert(f.prior_sqrt_info.end(), view.position_prior_sqrt_info, view.position_prior_sqrt_info + 9);
    f.any_prior |= prior;
    const bool gprior = free_cam && o.use_gravity_priors && view.has_gravity_prior;      // bundle_adjuster.cc:165-168
    f.has_gprior.push_back(gprior ? 1 : 0);
    f.gprior.insert(f.gprior.end(), view.gravity_prior, view.gravity_prior + 3);
    f.gprior_sqrt_info.insert(f.gprior_sqrt_info.end(), view.gravity_prior_sqrt_info, view.gravity_prior_sqrt_info + 9);
    f.any_gprior |= gprior;
    const bool oprior = free_cam && o.use_orientation_priors && view.has_orientation_prior;  // bundle_adjuster.cc:170-172
    f.has_oprior.push_back(oprior ? 1 : 0);
    f.oprior.insert(f.oprior.end(), view.orientation_prior, view.orientation_prior + 3);
    f.oprior_sqrt_info.insert(f.oprior_sqrt_info.end(), view.orientation_prior_sqrt_info, view.orientation_prior_sqrt_info + 9);
    f.any_oprior |= oprior;
    Intrinsics* in = view.camera.intr.get();
    if (!f.group_index.count(in)) {
      f.group_index[in] = (int)f.groups.size(); f.groups.push_back(in);
      f.intr_model.push_back(in->model);
      f.intr.insert(f.intr.end(), in->params, in->params + THB_INTR_STRIDE);
      f.intr_const.push_back(0xffff);                                      // constant until an optimised view claims it
    }
    const int g = f.group_index[in];
    if (free_cam) f.intr_const[g] = ConstantIntrinsicsMask(*in, o.intrinsics_to_optimize);   // bundle_adjuster.cc:382-460
    f.cam_group.push_back(g);
    return idx;
  };
  static std::atomic<uint64_t> g_epoch{0};
  const uint64_t epoch = ++g_epoch;
  auto add_track = [&](TrackId t, const Track& tr, bool free_pt) {
    if (tr.flat_epoch == epoch) { if (free_pt) f.pt_const[tr.flat_index] = 0; return tr.flat_index; }
    const int idx = (int)f.track_ids.size();
    tr.flat_epoch = epoch; tr.flat_index = idx;
    f.track_ids.push_back(t); f.track_ptrs.push_back(&tr);
    f.pts.insert(f.pts.end(), tr.point, tr.point + 4);
    f.pt_const.push_back(free_pt ? 0 : 1);
    return idx;
  };
  auto add_obs = [&](int ci, int pi, const Feature& feat) {
    f.obs_cam.push_back(ci); f.obs_pt.push_back(pi);
    f.obs_xy.push_back(feat.point[0]); f.obs_xy.push_back(feat.point[1]);
    f.obs_si.push_back(1.0 / std::sqrt(feat.cov[0])); f.obs_si.push_back(1.0 / std::sqrt(feat.cov[3]));   // reprojection_error.h:96-103
  };
  {  // one pass over the sizes so that the flat arrays grow once
    size_t nobs = 0;
    for (ViewId v : views) { auto vi = r->views.find(v); if (vi != r->views.end()) nobs += vi->second.track_order.size(); }
    f.obs_cam.reserve(nobs); f.obs_pt.reserve(nobs); f.obs_xy.reserve(2 * nobs); f.obs_si.reserve(2 * nobs);
    f.pts.reserve(4 * r->tracks.size()); f.pt_const.reserve(r->tracks.size()); f.track_ids.reserve(r->tracks.size());
  }
  bool all_views_optimised = vset.size() >= r->views.size();
  if (all_views_optimised) for (const auto& kv : r->views) if (!vset.count(kv.first)) { all_views_optimised = false; break; }
  // AddView. The two hash lookups per observation (track, feature) are the cost of this function (1M observations: ~120 ms on one
  // thread), and they are read-only: host threads resolve them view by view into pointer lists, the index assignment - which has to
  // see the observations in view order - then only walks those lists.
  struct Resolved { View* view = nullptr; std::vector<const Track*> tracks; std::vector<double> xysi; };  // per observation: x, y, 1/sigma_x, 1/sigma_y
  std::vector<Resolved> resolved(views.size());
  for (size_t i = 0; i < views.size(); ++i) {
    auto vi = r->views.find(views[i]);
    if (vi == r->views.end()) throw std::invalid_argument("unknown view id");   // the reference CHECK-aborts (bundle_adjuster.cc:117)
    resolved[i].view = &vi->second;
  }
  {
    std::atomic<size_t> next{0};
    std::atomic<bool> missing{false};
    auto work = [&]() {
      for (size_t i = next++; i < views.size(); i = next++) {
        Resolved& rv = resolved[i];
        if (!rv.view->estimated) continue;
        const auto& order = rv.view->track_order;
        rv.tracks.resize(order.size()); rv.xysi.resize(4 * order.size());
        for (size_t k = 0; k < order.size(); ++k) {
          auto ti = r->tracks.find(order[k]);
          auto fi = rv.view->features.find(order[k]);
          if (ti == r->tracks.end() || fi == rv.view->features.end()) { missing = true; return; }
          rv.tracks[k] = &ti->second;
          const Feature& feat = fi->second;
          double* q = &rv.xysi[4 * k];
          q[0] = feat.point[0]; q[1] = feat.point[1];
          q[2] = 1.0 / std::sqrt(feat.cov[0]); q[3] = 1.0 / std::sqrt(feat.cov[3]);   // reprojection_error.h:96-103
        }
      }
    };
    size_t total = 0;
    for (const Resolved& rv : resolved) total += rv.view->track_order.size();
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = total < 50000 ? 1 : std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)16, views.size()});
    std::vector<std::thread> pool;
    for (size_t k = 1; k < nthreads; ++k) pool.emplace_back(work);
    work();
    for (std::thread& th : pool) th.join();
    if (missing) throw std::out_of_range("a view lists a track that the reconstruction or the view's features do not hold");
  }
  for (size_t i = 0; i < views.size(); ++i) {
    const Resolved& rv = resolved[i];
    if (!rv.view->estimated) continue;
    const int ci = add_view(views[i], true);
    const auto& order = rv.view->track_order;
    for (size_t k = 0; k < order.size(); ++k) {
      const Track& tr = *rv.tracks[k];
      if (!tr.estimated) continue;
      const double* q = &rv.xysi[4 * k];
      f.obs_cam.push_back(ci); f.obs_pt.push_back(add_track(order[k], tr, false));
      f.obs_xy.push_back(q[0]); f.obs_xy.push_back(q[1]);
      f.obs_si.push_back(q[2]); f.obs_si.push_back(q[3]);
    }
  }
  for (TrackId t : tracks) {                                               // AddTrack
    auto ti = r->tracks.find(t);
    if (ti == r->tracks.end()) throw std::invalid_argument("unknown track id");
    if (!ti->second.estimated) continue;
    const int pi = add_track(t, ti->second, true);
    if (all_views_optimised) continue;  // every observing view is already in (or unestimated): nothing to add (bundle_adjuster.cc:192-198)
    for (ViewId v : ti->second.views) {
      View& view = r->views.at(v);
      if (vset.count(v) || !view.estimated) continue;
      add_obs(add_view(v, false), pi, view.features.at(t));
    }
  }
  BundleAdjustmentSummary out;
  if (f.obs_cam.empty()) { out.success = true; return out; }
  ThbBaProblem p;
  std::memset(&p, 0, sizeof(p));
  p.num_cameras = (int)f.view_ids.size(); p.num_groups = (int)f.groups.size();
  p.num_points = (int)f.track_ids.size(); p.num_observations = (int)f.obs_cam.size();
  p.memory_space = THB_MEM_HOST;
  p.cam_ext = f.cam_ext.data(); p.cam_const = f.cam_const.data(); p.cam_group = f.cam_group.data();
  p.intr = f.intr.data(); p.intr_model = f.intr_model.data(); p.intr_const = f.intr_const.data();
  p.pts = f.pts.data(); p.pt_const = f.pt_const.data();
  p.obs_cam = f.obs_cam.data(); p.obs_pt = f.obs_pt.data(); p.obs_xy = f.obs_xy.data(); p.obs_sqrt_info = f.obs_si.data();
  if (f.any_prior) { p.cam_has_position_prior = f.has_prior.data(); p.cam_position_prior = f.prior.data(); p.cam_position_prior_sqrt_info = f.prior_sqrt_info.data(); }
  if (f.any_gprior) { p.cam_has_gravity_prior = f.has_gprior.data(); p.cam_gravity_prior = f.gprior.data(); p.cam_gravity_prior_sqrt_info = f.gprior_sqrt_info.data(); }
  if (f.any_oprior) { p.cam_has_orientation_prior = f.has_oprior.data(); p.cam_orientation_prior = f.oprior.data(); p.cam_orientation_prior_sqrt_info = f.oprior_sqrt_info.data(); }
  const ThbBaOptions opt = MapOptions(o, force_no_inner);
  ThbBaSummary s;
  int rc;
  t_gather = since();
  {
    py::gil_scoped_release nogil;
    rc = thb_ba_solve(&p, &opt, &s, nullptr);
  }
  t_solve = since();
  if (rc == THB_E_NUMERICAL) return out;                                   // ceres FAILURE -> summary.success = false
  Check(rc);
  out.success = s.success != 0;
  out.initial_cost = s.initial_cost; out.final_cost = s.final_cost;
  out.setup_time_in_seconds = s.setup_time_in_seconds; out.solve_time_in_seconds = s.solve_time_in_seconds;
  if (!out.success) return out;
  for (size_t i = 0; i < f.view_ids.size(); ++i) std::copy_n(&f.cam_ext[6 * i], 6, r->views.at(f.view_ids[i]).camera.ext);
  for (size_t i = 0; i < f.track_ids.size(); ++i) std::copy_n(&f.pts[4 * i], 4, const_cast<Track*>(f.track_ptrs[i])->point);
  for (size_t g = 0; g < f.groups.size(); ++g) std::copy_n(&f.intr[THB_INTR_STRIDE * g], THB_INTR_STRIDE, f.groups[g]->params);
  if (prof) fprintf(stderr, "THB_ADAPTER_PROF gather %.2f ms, thb_ba_solve %.2f ms, scatter %.2f ms (%d observations)\n", t_gather, t_solve - t_gather, since() - t_solve, p.num_observations);
  if (cov) {  // ceres::Covariance at the refined parameters (p still points at them)
    std::vector<double> cc(36 * f.view_ids.size()), pc(9 * f.track_ids.size());
    std::vector<uint8_t> cok(f.view_ids.size()), pok(f.track_ids.size());
    {
      py::gil_scoped_release nogil;
      rc = thb_ba_covariance(&p, &opt, cc.data(), cok.data(), pc.data(), pok.data(), nullptr);
    }
    Check(rc);
    cov->ok = true;
    for (ViewId v : views) {
      auto it = f.view_index.find(v);
      if (it == f.view_index.end() || !cok[it->second]) { cov->ok = false; continue; }   // "could not be found or is set to fixed"
      std::copy_n(&cc[36 * it->second], 36, cov->views[v].begin());
    }
    std::unordered_map<TrackId, int> slot;                     // the gather keeps track slots in the Track objects, not in a map
    for (size_t i = 0; i < f.track_ids.size(); ++i) slot[f.track_ids[i]] = (int)i;
    for (TrackId t : tracks) {
      auto it = slot.find(t);
      if (it == slot.end() || !pok[it->second]) { cov->ok = false; continue; }
      std::copy_n(&pc[9 * it->second], 9, cov->tracks[t].begin());
    }
  }
  return out;
}

// UpdateInverseDepth (bundle_adjustment.cc:69-83); tracks are independent: host threads above 50 000 of them
void UpdateInverseDepth(const std::vector<TrackId>& ids, Reconstruction* r) {
  auto range = [&](size_t i0, size_t i1) {
    for (size_t i = i0; i < i1; ++i) {
      auto it = r->tracks.find(ids[i]);
      if (it == r->tracks.end() || !it->second.estimated) continue;
      auto vi = r->views.find(it->second.reference_view);
      if (vi == r->views.end()) continue;
      it->second.inverse_depth = 1.0 / vi->second.camera.ProjectRaw(it->second.point, nullptr);
    }
  };
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t nthreads = ids.size() < 50000 ? 1 : std::min<size_t>((size_t)(hw ? hw : 1), (size_t)16);
  if (nthreads <= 1) { range(0, ids.size()); return; }
  std::vector<std::thread> pool;
  const size_t per = (ids.size() + nthreads - 1) / nthreads;
  for (size_t k = 1; k < nthreads; ++k) pool.emplace_back(range, std::min(ids.size(), k * per), std::min(ids.size(), (k + 1) * per));
  range(0, std::min(ids.size(), per));
  for (std::thread& th : pool) th.join();
}
std::vector<TrackId> TracksOfViews(const std::vector<ViewId>& vs, Reconstruction* r) {
  std::vector<TrackId> out;
  for (ViewId v : vs) { auto it = r->views.find(v); if (it != r->views.end()) for (TrackId t : it->second.track_order) if (r->tracks.at(t).reference_view == v) out.push_back(t); }
  return out;
}

// solvers/sample_consensus_estimator.h:58-144
struct RansacParameters {
  double error_thresh = -1, failure_probability = 0.01, min_inlier_ratio = 0;
  int min_iterations = 100, max_iterations = std::numeric_limits<int>::max();
  bool use_mle = false, use_Tdd_test = false, use_lo = false;
  int lo_start_iterations = 50;
  int64_t seed = -1;  // ADDITIVE: the reference's `rng` member is not bindable (solvers.cc:89-102); < 0 => clock-seeded like the reference
};
struct RansacSummary {
  std::vector<int> inliers;
  int num_input_data_points = 0, num_iterations = 0, num_lo_iterations = 0;
  double confidence = 0;
};
struct FeatureCorrespondence { Feature feature1, feature2; };                       // matching/feature_correspondence.h:49-72
struct FeatureCorrespondence2D3D { double feature[2] = {0, 0}; double world_point[3] = {0, 0, 0}; };  // feature_correspondence_2d_3d.h
struct RelativePose { double E[9] = {0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0}; };      // estimate_relative_pose.h:49-53
struct CalibratedAbsolutePose { double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0}; };

ThbRansacParams ToC(const RansacParameters& q, int ransac_type) {
  ThbRansacParams p;
  thb_ransac_default_params(&p);
  p.error_thresh = q.error_thresh; p.failure_probability = q.failure_probability; p.min_inlier_ratio = q.min_inlier_ratio;
  p.min_iterations = q.min_iterations; p.max_iterations = q.max_iterations; p.use_mle = q.use_mle; p.use_lo = q.use_lo;
  p.lo_start_iterations = q.lo_start_iterations; p.ransac_type = ransac_type; p.use_tdd_test = q.use_Tdd_test;
  return p;
}
uint32_t SeedOf(const RansacParameters& q) { return q.seed >= 0 ? (uint32_t)q.seed : std::random_device{}(); }

typedef int (*BatchFn)(const ThbPairBatch*, const ThbRansacParams*, ThbRelPoseResult*, uint8_t*, void*);
bool RunOne(BatchFn fn, const RansacParameters& q, int type, const std::vector<double>& data, int width, ThbRelPoseResult* res, RansacSummary* sum) {
  const int64_t n = (int64_t)(data.size() / width);
  const int64_t off[2] = {0, n};
  const uint32_t seed = SeedOf(q);
  ThbPairBatch b = {1, THB_MEM_HOST, off, data.data(), &seed};
  const ThbRansacParams p = ToC(q, type);
  std::vector<uint8_t> mask((size_t)n);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = fn(&b, &p, res, mask.data(), nullptr);
  }
  Check(rc);
  sum->inliers.clear();
  for (int64_t i = 0; i < n; ++i) if (mask[i]) sum->inliers.push_back((int)i);
  sum->num_input_data_points = res->num_input_data_points; sum->num_iterations = res->num_iterations; sum->confidence = res->confidence; sum->num_lo_iterations = res->num_lo_iterations;
  return res->success != 0;
}
std::vector<double> Flatten(const std::vector<FeatureCorrespondence>& c) {
  std::vector<double> d(c.size() * 4);
  for (size_t i = 0; i < c.size(); ++i) { d[4 * i] = c[i].feature1.point[0]; d[4 * i + 1] = c[i].feature1.point[1]; d[4 * i + 2] = c[i].feature2.point[0]; d[4 * i + 3] = c[i].feature2.point[1]; }
  return d;
}
std::vector<double> Points2(const std::vector<Vec>& pts, size_t expect_min) {
  if (pts.size() < expect_min) throw std::invalid_argument("not enough points");
  std::vector<double> d(pts.size() * 2);
  for (size_t i = 0; i < pts.size(); ++i) CopyVec(pts[i], &d[2 * i], 2, "image point");
  return d;
}

// ---- EstimateTwoViewInfo (sfm/estimate_twoview_info.cc:133-305), calibrated branch -------------------------------------
template <int N> struct Prior { bool is_set = false; double value[N] = {0.0}; };   // camera_intrinsics_prior.h:50-79
struct CameraIntrinsicsPrior {                                                      // camera_intrinsics_prior.h:83-118
  int image_width = 0, image_height = 0;
  std::string camera_intrinsics_model_type = "PINHOLE";
  Prior<1> focal_length; Prior<2> principal_point; Prior<1> aspect_ratio; Prior<1> skew;
  Prior<4> radial_distortion; Prior<2> tangential_distortion;
  Prior<3> position; Prior<3> orientation; Prior<1> latitude; Prior<1> longitude; Prior<1> altitude;
};
struct TwoViewInfo {                                                                // twoview_info.h:54-90
  double focal_length_1 = 0, focal_length_2 = 0, position_2[3] = {0, 0, 0}, rotation_2[3] = {0, 0, 0};
  int num_verified_matches = 0, num_homography_inliers = 0, visibility_score = 0;
  double scale_estimate = -1.0;
};
struct EstimateTwoViewInfoOptions {                                                 // estimate_twoview_info.h:52-80
  int ransac_type = 0;
  double max_sampson_error_pixels = 6.0, expected_ransac_confidence = 0.9999;
  int min_ransac_iterations = 10, max_ransac_iterations = 1000;
  bool use_mle = true, use_lo = false;
  int lo_start_iterations = 10;
  double min_focal_length = 1.0, max_focal_length = std::numeric_limits<double>::max();
  int64_t seed = -1;  // ADDITIVE, like RansacParameters::seed: the reference's `rng` member is not bound (sfm.cc:864-885)
};

// <Model>::SetFromCameraIntrinsicsPriors on a default-constructed model, all six models (e.g. pinhole_camera_model.cc:74-107,
// double_sphere_camera_model.cc:74-106: prior.radial_distortion = [alpha, xi] but storage [.., XI = 5, ALPHA = 6];
// fisheye: focal guess 0.4 max(w, h) instead of 1.2; FOV / division reset their distortion when the prior has none)
int ModelFromString(const std::string& s) {
  if (s == "PINHOLE") return THB_MODEL_PINHOLE;
  if (s == "FISHEYE") return THB_MODEL_FISHEYE;
  if (s == "FOV") return THB_MODEL_FOV;
  if (s == "DIVISION_UNDISTORTION") return THB_MODEL_DIVISION_UNDISTORTION;
  if (s == "DOUBLE_SPHERE") return THB_MODEL_DOUBLE_SPHERE;
  if (s == "EXTENDED_UNIFIED") return THB_MODEL_EXTENDED_UNIFIED;
  throw std::invalid_argument("camera_intrinsics_model_type '" + s + "' is not on the B200 hot path");
}
void IntrinsicsFromPrior(const CameraIntrinsicsPrior& p, Intrinsics* in) {
  const int m = ModelFromString(p.camera_intrinsics_model_type);
  if (in->model != m) in->SetType(m);
  double* K = in->params;
  const bool sized = p.image_width != 0 && p.image_height != 0;
  if (p.focal_length.is_set) K[0] = p.focal_length.value[0];
  else if (sized) K[0] = (m == THB_MODEL_FISHEYE ? 0.4 : 1.2) * (double)std::max(p.image_width, p.image_height);
  const int cx = in->cx_index();
  if (p.principal_point.is_set) { K[cx] = p.principal_point.value[0]; K[cx + 1] = p.principal_point.value[1]; }
  else if (sized) { K[cx] = p.image_width / 2.0; K[cx + 1] = p.image_height / 2.0; }
  if (p.aspect_ratio.is_set) K[1] = p.aspect_ratio.value[0];
  if (in->has_skew() && p.skew.is_set) K[2] = p.skew.value[0];
  const double* rd = p.radial_distortion.value;
  switch (m) {
    case THB_MODEL_PINHOLE: if (p.radial_distortion.is_set) { K[5] = rd[0]; K[6] = rd[1]; } break;
    case THB_MODEL_FISHEYE: if (p.radial_distortion.is_set) for (int k = 0; k < 4; ++k) K[5 + k] = rd[k]; break;
    case THB_MODEL_FOV: K[4] = p.radial_distortion.is_set ? rd[0] : 0.75; break;
    case THB_MODEL_DIVISION_UNDISTORTION: K[4] = p.radial_distortion.is_set ? rd[0] : 0.0; break;
    case THB_MODEL_DOUBLE_SPHERE: if (p.radial_distortion.is_set) { K[6] = rd[0]; K[5] = rd[1]; } break;
    case THB_MODEL_EXTENDED_UNIFIED: if (p.radial_distortion.is_set) { K[5] = rd[0]; K[6] = rd[1]; } break;
  }
}
ThbViewIntrinsics ViewIntrinsicsFromPrior(const CameraIntrinsicsPrior& p) {
  Intrinsics in;
  IntrinsicsFromPrior(p, &in);
  ThbViewIntrinsics v;
  v.model = in.model; v.image_width = p.image_width; v.image_height = p.image_height; v.focal_length_is_set = p.focal_length.is_set ? 1 : 0;
  std::copy(in.params, in.params + THB_INTR_STRIDE, v.params);
  return v;
}
// PinholeCameraModel::SetFromCameraIntrinsicsPriors (pinhole_camera_model.cc:74-107) on a default-constructed model
void PinholeFromPrior(const CameraIntrinsicsPrior& p, double K[7]) {
  const double def[7] = {1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // f, aspect, skew, cx, cy, k1, k2
  std::copy(def, def + 7, K);
  if (p.focal_length.is_set) K[0] = p.focal_length.value[0];
  else if (p.image_width != 0 && p.image_height != 0) K[0] = 1.2 * (double)std::max(p.image_width, p.image_height);
  if (p.principal_point.is_set) { K[3] = p.principal_point.value[0]; K[4] = p.principal_point.value[1]; }
  else if (p.image_width != 0 && p.image_height != 0) { K[3] = p.image_width / 2.0; K[4] = p.image_height / 2.0; }
  if (p.aspect_ratio.is_set) K[1] = p.aspect_ratio.value[0];
  if (p.skew.is_set) K[2] = p.skew.value[0];
  if (p.radial_distortion.is_set) { K[5] = p.radial_distortion.value[0]; K[6] = p.radial_distortion.value[1]; }
}
// PinholeCameraModel::PixelToCameraCoordinates + UndistortPoint (pinhole_camera_model.h:214-241, 262-300), then hnormalized()
void PinholePixelToNormalized(const double K[7], const double px[2], double out[2]) {
  const double fy = K[0] * K[1];
  const double dy = (px[1] - K[4]) / fy;
  const double dx = (px[0] - K[3] - dy * K[2]) / K[0];
  double ux = dx, uy = dy;
  for (int it = 0; it < 100; ++it) {
    const double pxv = ux, pyv = uy;
    const double r2 = ux * ux + uy * uy;
    const double d = 1.0 + r2 * (K[5] + K[6] * r2);
    ux = dx / d; uy = dy / d;
    if (std::fabs(ux - pxv) < 1e-10 && std::fabs(uy - pyv) < 1e-10) break;
  }
  out[0] = ux / 1.0; out[1] = uy / 1.0;
}
struct VerificationOptions {                                                        // two_view_match_geometric_verification.h:53-92
  EstimateTwoViewInfoOptions estimate_twoview_info_options;
  int min_num_inlier_matches = 30;
  bool guided_matching = false, bundle_adjustment = true;
  double triangulation_max_reprojection_error = 15.0, min_triangulation_angle_degrees = 4.0, final_max_reprojection_error = 5.0;
};
// One call for a whole list of pairs: thb_estimate_two_view_info_batch / thb_verify_two_view_matches_batch. Returns, per pair,
// what sfm_wrapper.cc:12-26 returns for one: (success, TwoViewInfo, inlier / verified-match indices).
std::vector<py::tuple> RunTwoViewBatch(const VerificationOptions& vo, const std::vector<CameraIntrinsicsPrior>& p1,
                                       const std::vector<CameraIntrinsicsPrior>& p2, const std::vector<std::vector<FeatureCorrespondence>>& corr,
                                       const std::vector<uint32_t>& seeds, bool verify) {
  const EstimateTwoViewInfoOptions& o = vo.estimate_twoview_info_options;
  const size_t np = corr.size();
  if (p1.size() != np || p2.size() != np || (!seeds.empty() && seeds.size() != np)) throw std::invalid_argument("one prior pair (and seed) per list of correspondences");
  if (vo.guided_matching) throw std::runtime_error("guided_matching needs descriptors: not part of the B200 hot path");
  std::vector<ThbViewIntrinsics> a1(np), a2(np);
  std::vector<int64_t> off(np + 1, 0);
  std::vector<uint32_t> sd(np);
  for (size_t i = 0; i < np; ++i) {
    if (!(p1[i].focal_length.is_set && p2[i].focal_length.is_set))
      throw std::runtime_error("EstimateTwoViewInfo: the uncalibrated branch (EstimateUncalibratedRelativePose, 8-point) is not on the B200 hot path");
    a1[i] = ViewIntrinsicsFromPrior(p1[i]); a2[i] = ViewIntrinsicsFromPrior(p2[i]);
    off[i + 1] = off[i] + (int64_t)corr[i].size();
    sd[i] = seeds.empty() ? (o.seed >= 0 ? (uint32_t)(o.seed + (int64_t)i) : std::random_device{}()) : seeds[i];
  }
  std::vector<double> px((size_t)off[np] * 4);
  for (size_t i = 0; i < np; ++i)
    for (size_t k = 0; k < corr[i].size(); ++k) {
      double* d = &px[4 * ((size_t)off[i] + k)];
      d[0] = corr[i][k].feature1.point[0]; d[1] = corr[i][k].feature1.point[1]; d[2] = corr[i][k].feature2.point[0]; d[3] = corr[i][k].feature2.point[1];
    }
  ThbTwoViewOptions t;
  thb_two_view_default_options(&t);
  t.max_sampson_error_pixels = o.max_sampson_error_pixels; t.expected_ransac_confidence = o.expected_ransac_confidence;
  t.min_ransac_iterations = o.min_ransac_iterations; t.max_ransac_iterations = o.max_ransac_iterations; t.use_mle = o.use_mle; t.use_lo = o.use_lo;
  t.lo_start_iterations = o.lo_start_iterations; t.ransac_type = o.ransac_type;
  t.min_num_inlier_matches = vo.min_num_inlier_matches; t.bundle_adjustment = vo.bundle_adjustment;
  t.triangulation_max_reprojection_error = vo.triangulation_max_reprojection_error;
  t.min_triangulation_angle_degrees = vo.min_triangulation_angle_degrees; t.final_max_reprojection_error = vo.final_max_reprojection_error;
  ThbPairBatch b = {(int32_t)np, THB_MEM_HOST, off.data(), px.data(), sd.data()};
  std::vector<ThbTwoViewInfo> info(np);
  std::vector<uint8_t> mask((size_t)off[np]);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = verify ? thb_verify_two_view_matches_batch(&b, a1.data(), a2.data(), &t, info.data(), mask.data(), nullptr)
                : thb_estimate_two_view_info_batch(&b, a1.data(), a2.data(), &t, info.data(), mask.data(), nullptr);
  }
  Check(rc);
  std::vector<py::tuple> out;
  for (size_t i = 0; i < np; ++i) {
    TwoViewInfo ti;
    std::vector<int> idx;
    if (info[i].success || info[i].num_verified_matches > 0) {
      ti.focal_length_1 = info[i].focal_length_1; ti.focal_length_2 = info[i].focal_length_2;
      std::copy_n(info[i].position_2, 3, ti.position_2); std::copy_n(info[i].rotation_2, 3, ti.rotation_2);
      ti.num_verified_matches = info[i].num_verified_matches; ti.num_homography_inliers = info[i].num_homography_inliers;
      ti.visibility_score = info[i].visibility_score;
      for (int64_t k = off[i]; k < off[i + 1]; ++k) if (mask[(size_t)k]) idx.push_back((int)(k - off[i]));
    }
    out.push_back(py::make_tuple(info[i].success != 0, ti, idx));
  }
  return out;
}
// ---- TrackEstimator (sfm/estimate_track.{h,cc}) over thb_estimate_tracks_batch: every unestimated track in ONE launch ------
struct TrackEstimatorOptions {                                                      // estimate_track.h:59-84
  int num_threads = 1;
  double max_acceptable_reprojection_error_pixels = 5.0, min_triangulation_angle_degrees = 3.0;
  bool bundle_adjustment = true;
  BundleAdjustmentOptions ba_options;
  int multithreaded_step_size = 100;
  int triangulation_method = 0;                                                     // TriangulationMethodType::MIDPOINT
};
struct TrackEstimatorSummary {                                                      // estimate_track.h:86-96
  int input_num_estimated_tracks = 0, num_triangulation_attempts = 0;
  std::unordered_set<TrackId> estimated_tracks;
  int num_bad_angles = 0, num_failed_triangulations = 0, num_bad_reprojections = 0; // ADDITIVE: the counters the reference only logs
};
struct TrackEstimator {
  TrackEstimatorOptions options;
  Reconstruction* recon;
  TrackEstimator(const TrackEstimatorOptions& o, Reconstruction* r) : options(o), recon(r) {}

  TrackEstimatorSummary EstimateAllTracks() {                                       // estimate_track.cc:124-139
    std::vector<TrackId> ids; std::unordered_set<TrackId> seen;
    for (ViewId v : recon->view_order) {
      const View& view = recon->views.at(v);
      if (!view.estimated) continue;
      for (TrackId t : view.track_order) if (seen.insert(t).second) ids.push_back(t);
    }
    return EstimateTracks(ids);
  }

  TrackEstimatorSummary EstimateTracks(const std::vector<TrackId>& track_ids) {    // estimate_track.cc:141-203 + :209-321
    const BundleAdjustmentOptions& bo = options.ba_options;
    if (options.triangulation_method != 0) throw std::runtime_error("only TriangulationMethodType.MIDPOINT is implemented");
    if (bo.use_inverse_depth_parametrization) throw std::runtime_error("use_inverse_depth_parametrization is not implemented");
    TrackEstimatorSummary sum;
    Flat f;
    std::vector<double> rays;
    for (TrackId t : track_ids) {
      auto ti = recon->tracks.find(t);
      if (ti == recon->tracks.end()) throw std::invalid_argument("unknown track id");
      if (ti->second.estimated) { ++sum.input_num_estimated_tracks; continue; }
      if (f.track_index.count(t)) continue;
      const int pi = (int)f.track_ids.size();
      f.track_index[t] = pi; f.track_ids.push_back(t);
      f.pts.insert(f.pts.end(), ti->second.point, ti->second.point + 4);
      for (ViewId v : ti->second.views) {                                           // GetObservationsFromTrackViews, :57-90
        auto vi = recon->views.find(v);
        if (vi == recon->views.end() || !vi->second.estimated) continue;
        View& view = vi->second;
        int ci;
        auto it = f.view_index.find(v);
        if (it != f.view_index.end()) ci = it->second;
        else {
          ci = (int)f.view_ids.size();
          f.view_index[v] = ci; f.view_ids.push_back(v);
          f.cam_ext.insert(f.cam_ext.end(), view.camera.ext, view.camera.ext + 6);
          Intrinsics* in = view.camera.intr.get();
          if (!f.group_index.count(in)) {
            f.group_index[in] = (int)f.groups.size(); f.groups.push_back(in);
            f.intr_model.push_back(in->model);
            f.intr.insert(f.intr.end(), in->params, in->params + THB_INTR_STRIDE);
          }
          f.cam_group.push_back(f.group_index[in]);
        }
        const Feature& feat = view.features.at(t);
        f.obs_cam.push_back(ci); f.obs_pt.push_back(pi);
        f.obs_xy.push_back(feat.point[0]); f.obs_xy.push_back(feat.point[1]);
        f.obs_si.push_back(1.0 / std::sqrt(feat.cov[0])); f.obs_si.push_back(1.0 / std::sqrt(feat.cov[3]));
        // Camera::PixelToUnitDepthRay(feature).normalized() (camera.cc:218-226)
        double u[3], R[9];
        thb::pixel_to_camera(view.camera.intr->model, view.camera.intr->params, feat.point, u);  // PixelToNormalizedCoordinates, any model
        AngleAxisToRotation(view.camera.ext + 3, R);
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = R[0 * 3 + k] * u[0] + R[1 * 3 + k] * u[1] + R[2 * 3 + k] * u[2];   // R^T u
        const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int k = 0; k < 3; ++k) rays.push_back(d[k] / nd);
      }
    }
    sum.num_triangulation_attempts = (int)f.track_ids.size();
    if (f.track_ids.empty()) return sum;
    ThbBaProblem p;
    std::memset(&p, 0, sizeof(p));
    p.num_cameras = (int)f.view_ids.size(); p.num_groups = (int)f.groups.size();
    p.num_points = (int)f.track_ids.size(); p.num_observations = (int)f.obs_cam.size();
    p.memory_space = THB_MEM_HOST;
    p.cam_ext = f.cam_ext.data(); p.cam_group = f.cam_group.data(); p.intr = f.intr.data(); p.intr_model = f.intr_model.data();
    p.pts = f.pts.data(); p.obs_cam = f.obs_cam.data(); p.obs_pt = f.obs_pt.data(); p.obs_xy = f.obs_xy.data(); p.obs_sqrt_info = f.obs_si.data();
    const ThbBaOptions opt = MapOptions(bo, true);                                  // BundleAdjustTrack: no inner iterations
    ThbTrackEstimatorOptions eo;
    eo.max_acceptable_reprojection_error_pixels = options.max_acceptable_reprojection_error_pixels;
    eo.min_triangulation_angle_degrees = options.min_triangulation_angle_degrees;
    eo.bundle_adjustment = options.bundle_adjustment ? 1 : 0; eo.reserved0 = 0;
    std::vector<int32_t> status(f.track_ids.size());
    int rc;
    {
      py::gil_scoped_release nogil;
      rc = thb_estimate_tracks_batch(&p, rays.data(), &eo, &opt, status.data(), nullptr, nullptr);
    }
    Check(rc);
    for (size_t i = 0; i < f.track_ids.size(); ++i) {
      Track& tr = recon->tracks.at(f.track_ids[i]);
      switch (status[i]) {
        case THB_TRACK_ESTIMATED: std::copy_n(&f.pts[4 * i], 4, tr.point); tr.estimated = true; sum.estimated_tracks.insert(f.track_ids[i]); break;
        case THB_TRACK_BAD_ANGLE: ++sum.num_bad_angles; break;
        case THB_TRACK_FAILED_TRIANGULATION: ++sum.num_failed_triangulations; break;
        case THB_TRACK_BAD_REPROJECTION: ++sum.num_bad_reprojections; break;
        default: break;
      }
    }
    std::vector<TrackId> done(sum.estimated_tracks.begin(), sum.estimated_tracks.end());
    UpdateInverseDepth(done, recon);                                                // bundle_adjustment.cc:69-83, 283
    return sum;
  }
};

// reconstruction_estimator_utils.cc:98-110
double ComputeResolutionScaledThreshold(double threshold_pixels, int w, int h) {
  if (w == 0 && h == 0) return threshold_pixels;
  return threshold_pixels * (double)std::max(w, h) / 1024.0;
}

}  // namespace

PYBIND11_MODULE(_pt, m) {
  m.doc() = "theia-b200: pt.sfm / pt.solvers / pt.matching names of pyTheia for the B200 hot paths";
  py::module_ sfm = m.def_submodule("sfm");
  py::module_ solvers = m.def_submodule("solvers");
  py::module_ matching = m.def_submodule("matching");

  py::enum_<OptimizeIntrinsicsType>(sfm, "OptimizeIntrinsicsType", py::arithmetic())
      .value("NONE", NONE).value("FOCAL_LENGTH", FOCAL_LENGTH).value("ASPECT_RATIO", ASPECT_RATIO).value("SKEW", SKEW)
      .value("PRINCIPAL_POINTS", PRINCIPAL_POINTS).value("RADIAL_DISTORTION", RADIAL_DISTORTION)
      .value("TANGENTIAL_DISTORTION", TANGENTIAL_DISTORTION).value("ALL", ALL);
  py::enum_<LinearSolverType>(sfm, "LinearSolverType")
      .value("DENSE_NORMAL_CHOLESKY", DENSE_NORMAL_CHOLESKY).value("DENSE_QR", DENSE_QR).value("SPARSE_NORMAL_CHOLESKY", SPARSE_NORMAL_CHOLESKY)
      .value("DENSE_SCHUR", DENSE_SCHUR).value("SPARSE_SCHUR", SPARSE_SCHUR).value("ITERATIVE_SCHUR", ITERATIVE_SCHUR).value("CGNR", CGNR);
  struct LossNs {}; struct ModelNs {}; struct RansacNs {}; struct PnPNs {};
  py::class_<LossNs>(sfm, "LossFunctionType")
      .def_property_readonly_static("TRIVIAL", [](py::object) { return THB_LOSS_TRIVIAL; }).def_property_readonly_static("HUBER", [](py::object) { return THB_LOSS_HUBER; })
      .def_property_readonly_static("SOFTLONE", [](py::object) { return THB_LOSS_SOFTLONE; }).def_property_readonly_static("CAUCHY", [](py::object) { return THB_LOSS_CAUCHY; })
      .def_property_readonly_static("ARCTAN", [](py::object) { return THB_LOSS_ARCTAN; }).def_property_readonly_static("TUKEY", [](py::object) { return THB_LOSS_TUKEY; })
      .def_property_readonly_static("TRUNCATED", [](py::object) { return THB_LOSS_TRUNCATED; });
  py::class_<ModelNs>(sfm, "CameraIntrinsicsModelType")
      .def_property_readonly_static("PINHOLE", [](py::object) { return THB_MODEL_PINHOLE; }).def_property_readonly_static("FISHEYE", [](py::object) { return THB_MODEL_FISHEYE; })
      .def_property_readonly_static("FOV", [](py::object) { return THB_MODEL_FOV; })
      .def_property_readonly_static("DIVISION_UNDISTORTION", [](py::object) { return THB_MODEL_DIVISION_UNDISTORTION; })
      .def_property_readonly_static("DOUBLE_SPHERE", [](py::object) { return THB_MODEL_DOUBLE_SPHERE; })
      .def_property_readonly_static("EXTENDED_UNIFIED", [](py::object) { return THB_MODEL_EXTENDED_UNIFIED; });
  py::class_<RansacNs>(sfm, "RansacType")
      .def_property_readonly_static("RANSAC", [](py::object) { return 0; }).def_property_readonly_static("PROSAC", [](py::object) { return 1; })
      .def_property_readonly_static("LMED", [](py::object) { return 2; }).def_property_readonly_static("EXHAUSTIVE", [](py::object) { return 3; });
  py::class_<PnPNs>(sfm, "PnPType")
      .def_property_readonly_static("KNEIP", [](py::object) { return 0; }).def_property_readonly_static("DLS", [](py::object) { return 1; })
      .def_property_readonly_static("SQPnP", [](py::object) { return 2; });

  py::class_<Feature>(sfm, "Feature")
      .def(py::init<>())
      .def(py::init([](const Vec& p) { Feature f; CopyVec(p, f.point, 2, "point"); return f; }))
      .def(py::init([](const Vec& p, const Vec& c) { Feature f; CopyVec(p, f.point, 2, "point"); CopyVec(c, f.cov, 4, "covariance"); return f; }))
      .def_property("point", [](const Feature& f) { return MakeVec(f.point, 2); }, [](Feature& f, const Vec& p) { CopyVec(p, f.point, 2, "point"); })
      .def_property("covariance", [](const Feature& f) { return MakeMat(f.cov, 2, 2); }, [](Feature& f, const Vec& c) { CopyVec(c, f.cov, 4, "covariance"); });

  py::class_<Camera>(sfm, "Camera")
      .def(py::init<>())
      .def("DeepCopy", &Camera::DeepCopy)
      .def("SetPosition", &Camera::SetPosition).def("GetPosition", &Camera::GetPosition)
      .def("SetOrientationFromAngleAxis", &Camera::SetOrientationFromAngleAxis).def("GetOrientationAsAngleAxis", &Camera::GetOrientationAsAngleAxis)
      .def("SetOrientationFromRotationMatrix", &Camera::SetOrientationFromRotationMatrix)
      .def("GetOrientationAsRotationMatrix", &Camera::GetOrientationAsRotationMatrix)
      .def("SetFocalLength", &Camera::SetFocalLength).def("FocalLength", &Camera::FocalLength)
      .def("SetPrincipalPoint", &Camera::SetPrincipalPoint).def("PrincipalPointX", &Camera::PrincipalPointX).def("PrincipalPointY", &Camera::PrincipalPointY)
      .def("SetImageSize", &Camera::SetImageSize).def("ImageWidth", [](const Camera& c) { return c.width; }).def("ImageHeight", [](const Camera& c) { return c.height; })
      .def("SetCameraIntrinsicsModelType", &Camera::SetCameraIntrinsicsModelType).def("GetCameraIntrinsicsModelType", &Camera::GetCameraIntrinsicsModelType)
      .def("Parameters", &Camera::Parameters).def("SetParameters", &Camera::SetParameters)
      .def("ProjectPoint", &Camera::ProjectPoint)
      .def("SetFromCameraIntrinsicsPriors", [](Camera& c, const CameraIntrinsicsPrior& p) {  // camera.cc:126-136
        IntrinsicsFromPrior(p, c.intr.get());
        c.width = p.image_width; c.height = p.image_height; })
      .def("PixelToNormalizedCoordinates", [](const Camera& c, const Vec& px) {
        double p[2], out[3]; CopyVec(px, p, 2, "pixel");
        thb::pixel_to_camera(c.intr->model, c.intr->params, p, out);
        return MakeVec(out, 3); })
      .def("PixelToUnitDepthRay", [](const Camera& c, const Vec& px) {  // camera.cc:218-226
        double p[2], u[3], R[9], d[3]; CopyVec(px, p, 2, "pixel");
        thb::pixel_to_camera(c.intr->model, c.intr->params, p, u);
        AngleAxisToRotation(c.ext + 3, R);
        for (int k = 0; k < 3; ++k) d[k] = R[0 * 3 + k] * u[0] + R[1 * 3 + k] * u[1] + R[2 * 3 + k] * u[2];
        return MakeVec(d, 3); });

  py::class_<Track>(sfm, "Track")
      .def("SetPoint", &Track::SetPoint).def("Point", &Track::Point)
      .def("SetIsEstimated", [](Track& t, bool e) { t.estimated = e; }).def("IsEstimated", [](const Track& t) { return t.estimated; })
      .def("NumViews", [](const Track& t) { return (int)t.views.size(); }).def("ViewIds", [](const Track& t) { return t.views; })
      .def("InverseDepth", [](const Track& t) { return t.inverse_depth; }).def("ReferenceViewId", [](const Track& t) { return t.reference_view; });

  py::class_<View>(sfm, "View")
      .def("Name", [](const View& v) { return v.name; })
      .def("SetIsEstimated", [](View& v, bool e) { v.estimated = e; }).def("IsEstimated", [](const View& v) { return v.estimated; })
      .def("Camera", [](View& v) -> Camera& { return v.camera; }, py::return_value_policy::reference_internal)
      .def("MutableCamera", [](View& v) -> Camera& { return v.camera; }, py::return_value_policy::reference_internal)
      .def("NumFeatures", [](const View& v) { return (int)v.features.size(); }).def("TrackIds", &View::TrackIds)
      .def("SetPositionPrior", [](View& v, const Vec& prior, const py::array_t<double, py::array::c_style | py::array::forcecast>& sqrt_information) {
        CopyVec(prior, v.position_prior, 3, "position prior");
        if (sqrt_information.ndim() != 2 || sqrt_information.shape(0) != 3 || sqrt_information.shape(1) != 3) throw std::invalid_argument("sqrt information must be 3 x 3");
        std::copy_n(sqrt_information.data(), 9, v.position_prior_sqrt_info);
        v.has_position_prior = true;
      })
      .def("HasPositionPrior", [](const View& v) { return v.has_position_prior; })
      .def("SetGravityPrior", [](View& v, const Vec& prior, const py::array_t<double, py::array::c_style | py::array::forcecast>& sqrt_information) {
        CopyVec(prior, v.gravity_prior, 3, "gravity prior");
        if (sqrt_information.ndim() != 2 || sqrt_information.shape(0) != 3 || sqrt_information.shape(1) != 3) throw std::invalid_argument("sqrt information must be 3 x 3");
        std::copy_n(sqrt_information.data(), 9, v.gravity_prior_sqrt_info);
        v.has_gravity_prior = true;
      })
      .def("HasGravityPrior", [](const View& v) { return v.has_gravity_prior; })
      .def("SetOrientationPrior", [](View& v, const Vec& prior, const py::array_t<double, py::array::c_style | py::array::forcecast>& sqrt_information) {
        CopyVec(prior, v.orientation_prior, 3, "orientation prior");
        if (sqrt_information.ndim() != 2 || sqrt_information.shape(0) != 3 || sqrt_information.shape(1) != 3) throw std::invalid_argument("sqrt information must be 3 x 3");
        std::copy_n(sqrt_information.data(), 9, v.orientation_prior_sqrt_info);
        v.has_orientation_prior = true;
      })
      .def("HasOrientationPrior", [](const View& v) { return v.has_orientation_prior; })
      .def("GetOrientationPrior", [](const View& v) { py::array_t<double> a(3); std::copy_n(v.orientation_prior, 3, a.mutable_data()); return a; })
      .def("GetOrientationPriorSqrtInformation", [](const View& v) { py::array_t<double> a({3, 3}); std::copy_n(v.orientation_prior_sqrt_info, 9, a.mutable_data()); return a; })
      .def("GetGravityPrior", [](const View& v) { py::array_t<double> a(3); std::copy_n(v.gravity_prior, 3, a.mutable_data()); return a; })
      .def("GetGravityPriorSqrtInformation", [](const View& v) { py::array_t<double> a({3, 3}); std::copy_n(v.gravity_prior_sqrt_info, 9, a.mutable_data()); return a; })
      .def("GetPositionPrior", [](const View& v) { py::array_t<double> a(3); std::copy_n(v.position_prior, 3, a.mutable_data()); return a; })
      .def("GetPositionPriorSqrtInformation", [](const View& v) { py::array_t<double> a({3, 3}); std::copy_n(v.position_prior_sqrt_info, 9, a.mutable_data()); return a; })
      .def("GetFeature", [](const View& v, TrackId t) { const Feature* f = v.GetFeature(t); if (!f) throw std::invalid_argument("no such feature"); return *f; });

  py::class_<Reconstruction>(sfm, "Reconstruction")
      .def(py::init<>())
      .def("AddView", &Reconstruction::AddView, py::arg("name"), py::arg("group_id"), py::arg("timestamp"))
      .def("AddView", &Reconstruction::AddViewAutoGroup, py::arg("name"), py::arg("timestamp"))
      .def("AddTrack", &Reconstruction::AddTrack)
      .def("AddObservation", &Reconstruction::AddObservation)
      .def("View", [](Reconstruction& r, ViewId v) -> View* { return r.MutableView(v); }, py::return_value_policy::reference_internal)
      .def("MutableView", &Reconstruction::MutableView, py::return_value_policy::reference_internal)
      .def("Track", [](Reconstruction& r, TrackId t) -> Track* { return r.MutableTrack(t); }, py::return_value_policy::reference_internal)
      .def("MutableTrack", &Reconstruction::MutableTrack, py::return_value_policy::reference_internal)
      .def("ViewIds", [](const Reconstruction& r) { return r.view_order; }).def("TrackIds", [](const Reconstruction& r) { return r.track_order; })
      .def("NumViews", [](const Reconstruction& r) { return (int)r.views.size(); }).def("NumTracks", [](const Reconstruction& r) { return (int)r.tracks.size(); })
      .def("CameraIntrinsicsGroupIdFromViewId", &Reconstruction::CameraIntrinsicsGroupIdFromViewId);

  py::class_<BundleAdjustmentOptions>(sfm, "BundleAdjustmentOptions")
      .def(py::init<>())
      .def_readwrite("loss_function_type", &BundleAdjustmentOptions::loss_function_type)
      .def_readwrite("robust_loss_width", &BundleAdjustmentOptions::robust_loss_width)
      .def_readwrite("linear_solver_type", &BundleAdjustmentOptions::linear_solver_type)
      .def_readwrite("verbose", &BundleAdjustmentOptions::verbose)
      .def_readwrite("constant_camera_orientation", &BundleAdjustmentOptions::constant_camera_orientation)
      .def_readwrite("constant_camera_position", &BundleAdjustmentOptions::constant_camera_position)
      .def_readwrite("use_homogeneous_point_parametrization", &BundleAdjustmentOptions::use_homogeneous_point_parametrization)
      .def_readwrite("use_inverse_depth_parametrization", &BundleAdjustmentOptions::use_inverse_depth_parametrization)
      .def_readwrite("intrinsics_to_optimize", &BundleAdjustmentOptions::intrinsics_to_optimize)
      .def_readwrite("num_threads", &BundleAdjustmentOptions::num_threads)
      .def_readwrite("max_num_iterations", &BundleAdjustmentOptions::max_num_iterations)
      .def_readwrite("max_solver_time_in_seconds", &BundleAdjustmentOptions::max_solver_time_in_seconds)
      .def_readwrite("use_inner_iterations", &BundleAdjustmentOptions::use_inner_iterations)
      .def_readwrite("function_tolerance", &BundleAdjustmentOptions::function_tolerance)
      .def_readwrite("gradient_tolerance", &BundleAdjustmentOptions::gradient_tolerance)
      .def_readwrite("parameter_tolerance", &BundleAdjustmentOptions::parameter_tolerance)
      .def_readwrite("max_trust_region_radius", &BundleAdjustmentOptions::max_trust_region_radius)
      .def_readwrite("use_position_priors", &BundleAdjustmentOptions::use_position_priors)
      .def_readwrite("use_orientation_priors", &BundleAdjustmentOptions::use_orientation_priors)
      .def_readwrite("use_depth_priors", &BundleAdjustmentOptions::use_depth_priors)
      .def_readwrite("orthographic_camera", &BundleAdjustmentOptions::orthographic_camera)
      .def_readwrite("use_gravity_priors", &BundleAdjustmentOptions::use_gravity_priors);
  py::class_<BundleAdjustmentSummary>(sfm, "BundleAdjustmentSummary")
      .def_readonly("success", &BundleAdjustmentSummary::success).def_readonly("initial_cost", &BundleAdjustmentSummary::initial_cost)
      .def_readonly("final_cost", &BundleAdjustmentSummary::final_cost)
      .def_readonly("setup_time_in_seconds", &BundleAdjustmentSummary::setup_time_in_seconds)
      .def_readonly("solve_time_in_seconds", &BundleAdjustmentSummary::solve_time_in_seconds);

  // bundle_adjustment.cc:188-217 / :111-143 / :220-258 / :261-285,389-418 and their wrappers' argument orders
  sfm.def("BundleAdjustReconstruction", [](const BundleAdjustmentOptions& o, Reconstruction& r) {
    BundleAdjustmentSummary s = RunBa(o, r.view_order, r.track_order, &r, false);
    UpdateInverseDepth(r.track_order, &r);
    return s;
  });
  sfm.def("BundleAdjustPartialReconstruction", [](const BundleAdjustmentOptions& o, const std::vector<ViewId>& v, const std::vector<TrackId>& t, Reconstruction& r) {
    BundleAdjustmentSummary s = RunBa(o, v, t, &r, false);
    UpdateInverseDepth(t, &r);
    return s;
  });
  // BundleAdjustPartialViewsConstant (bundle_adjustment.cc:146-186): the listed views free, every other view that observes a track
  // constant, all tracks free - AddView(var) + AddView(const) + SetCameraExtrinsicsConstant(const) + AddTrack(all) flattens to the
  // same blocks as AddView(var) + AddTrack(all) (AddTrack registers the remaining observers as constant cameras, :199-204).
  sfm.def("BundleAdjustPartialViewsConstant", [](const BundleAdjustmentOptions& o, const std::vector<ViewId>& var_views, const std::vector<ViewId>& const_views,
                                               Reconstruction& r) {
    (void)const_views;
    BundleAdjustmentSummary s = RunBa(o, var_views, r.track_order, &r, false);
    UpdateInverseDepth(r.track_order, &r);
    return s;
  });
  sfm.def("BundleAdjustView", [](Reconstruction& r, const BundleAdjustmentOptions& o, ViewId v) {
    BundleAdjustmentSummary s = RunBa(o, {v}, {}, &r, true);   // forces DENSE_QR, no inner iterations (:225)
    UpdateInverseDepth(TracksOfViews({v}, &r), &r);
    return s;
  });
  sfm.def("BundleAdjustViews", [](Reconstruction& r, const BundleAdjustmentOptions& o, const std::vector<ViewId>& v) {
    BundleAdjustmentSummary s = RunBa(o, v, {}, &r, true);
    UpdateInverseDepth(TracksOfViews(v, &r), &r);
    return s;
  });
  // the covariance overloads (bundle_adjustment_wrapper.cc:52-96): (summary, covariance(s), empirical variance factor)
  auto to_mat = [](const double* d, int n) {
    py::array_t<double> a({n, n});
    std::copy_n(d, n * n, a.mutable_data());
    return a;
  };
  sfm.def("BundleAdjustViewWithCov", [to_mat](Reconstruction& r, const BundleAdjustmentOptions& o, ViewId v) {
    CovOut cov;
    BundleAdjustmentSummary s = RunBa(o, {v}, {}, &r, true, &cov);
    double factor = 1.0, eye[36] = {0};
    for (int k = 0; k < 6; ++k) eye[k * 7] = 1.0;
    std::array<double, 36> m; std::copy_n(eye, 36, m.begin());
    if (s.success) {
      if (!cov.ok) s.success = false;                                                      // :441-443
      else {
        const double redundancy = (double)r.views.at(v).features.size() * 2 - 6;           // View::NumFeatures() (:445-447)
        factor = 2.0 * s.final_cost / redundancy;
        m = cov.views.at(v);
        for (double& x : m) x *= factor;
      }
    }
    UpdateInverseDepth(TracksOfViews({v}, &r), &r);
    return py::make_tuple(s, to_mat(m.data(), 6), factor);
  });
  sfm.def("BundleAdjustViewsWithCov", [to_mat](Reconstruction& r, const BundleAdjustmentOptions& o, const std::vector<ViewId>& v) {
    CovOut cov;
    BundleAdjustmentSummary s = RunBa(o, v, {}, &r, true, &cov);
    double factor = 1.0;
    py::dict mats;
    if (s.success) {
      if (!cov.ok) s.success = false;
      else {
        double nr_obs = 0;
        for (ViewId id : v) nr_obs += (double)r.views.at(id).features.size();
        factor = 2.0 * s.final_cost / (nr_obs * 2 - 6.0 * (double)v.size());                 // :478-490
        for (auto& kv : cov.views) { for (double& x : kv.second) x *= factor; mats[py::int_(kv.first)] = to_mat(kv.second.data(), 6); }
      }
    }
    UpdateInverseDepth(TracksOfViews(v, &r), &r);
    return py::make_tuple(s, mats, factor);
  });
  sfm.def("BundleAdjustTrackWithCov", [to_mat](Reconstruction& r, const BundleAdjustmentOptions& o, TrackId t) {
    BundleAdjustmentOptions oo = o; oo.use_homogeneous_point_parametrization = true;       // :296
    CovOut cov;
    BundleAdjustmentSummary s = RunBa(oo, {}, {t}, &r, true, &cov);
    double factor = 1.0;
    std::array<double, 9> m = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (s.success) {
      if (!cov.ok) s.success = false;
      else {
        const double redundancy = (double)r.tracks.at(t).views.size() * 2 - 3;              // Track::NumViews() (:312-316)
        factor = 2.0 * s.final_cost / redundancy;
        m = cov.tracks.at(t);
        for (double& x : m) x *= factor;
      }
    }
    UpdateInverseDepth({t}, &r);
    return py::make_tuple(s, to_mat(m.data(), 3), factor);
  });
  sfm.def("BundleAdjustTracksWithCov", [to_mat](Reconstruction& r, const BundleAdjustmentOptions& o, const std::vector<TrackId>& t) {
    BundleAdjustmentOptions oo = o; oo.use_homogeneous_point_parametrization = true;
    CovOut cov;
    BundleAdjustmentSummary s = RunBa(oo, {}, t, &r, true, &cov);
    double factor = 1.0;
    py::dict mats;
    if (s.success) {
      if (!cov.ok) s.success = false;
      else {
        double nr_obs = 0;
        for (TrackId id : t) nr_obs += (double)r.tracks.at(id).views.size();
        factor = 2.0 * s.final_cost / (nr_obs * 2 - 3.0 * (double)t.size());                 // :358-369
        for (auto& kv : cov.tracks) { for (double& x : kv.second) x *= factor; mats[py::int_(kv.first)] = to_mat(kv.second.data(), 3); }
      }
    }
    UpdateInverseDepth(t, &r);
    return py::make_tuple(s, mats, factor);
  });
  // SetOutlierTracksToUnestimated (sfm_wrapper.cc:46-57 -> set_outlier_tracks_to_unestimated.cc:62-137): every listed track in
  // one launch over thb_set_outlier_tracks_batch; tracks with status > 0 become unestimated. Returns the number removed.
  sfm.def("SetOutlierTracksToUnestimated", [](const std::unordered_set<TrackId>& track_ids, double max_err, double min_angle, Reconstruction& r) {
    Flat f;
    for (TrackId t : track_ids) {
      auto it = r.tracks.find(t);
      if (it == r.tracks.end() || !it->second.estimated) continue;                       // :76-78
      const int pi = (int)f.track_ids.size();
      f.track_ids.push_back(t);
      f.pts.insert(f.pts.end(), it->second.point, it->second.point + 4);
      for (ViewId v : it->second.views) {
        const View& view = r.views.at(v);
        if (!view.estimated) continue;                                                    // :86-88
        int ci;
        auto vi = f.view_index.find(v);
        if (vi != f.view_index.end()) ci = vi->second;
        else {
          ci = (int)f.view_ids.size();
          f.view_index[v] = ci; f.view_ids.push_back(v);
          f.cam_ext.insert(f.cam_ext.end(), view.camera.ext, view.camera.ext + 6);
          Intrinsics* in = view.camera.intr.get();
          if (!f.group_index.count(in)) {
            f.group_index[in] = (int)f.groups.size(); f.groups.push_back(in);
            f.intr_model.push_back(in->model);
            f.intr.insert(f.intr.end(), in->params, in->params + THB_INTR_STRIDE);
          }
          f.cam_group.push_back(f.group_index[in]);
        }
        const Feature& feat = view.features.at(t);
        f.obs_cam.push_back(ci); f.obs_pt.push_back(pi);
        f.obs_xy.push_back(feat.point[0]); f.obs_xy.push_back(feat.point[1]);
      }
    }
    if (f.track_ids.empty()) return 0;
    ThbBaProblem p;
    std::memset(&p, 0, sizeof(p));
    p.num_cameras = (int)f.view_ids.size(); p.num_groups = (int)f.groups.size(); p.num_points = (int)f.track_ids.size();
    p.num_observations = (int)f.obs_cam.size(); p.memory_space = THB_MEM_HOST;
    p.cam_ext = f.cam_ext.data(); p.cam_group = f.cam_group.data(); p.intr = f.intr.data(); p.intr_model = f.intr_model.data();
    p.pts = f.pts.data(); p.obs_cam = f.obs_cam.data(); p.obs_pt = f.obs_pt.data(); p.obs_xy = f.obs_xy.data();
    std::vector<int32_t> status(f.track_ids.size());
    int32_t removed = 0;
    int rc;
    {
      py::gil_scoped_release nogil;
      rc = thb_set_outlier_tracks_batch(&p, max_err, min_angle, status.data(), &removed, nullptr);
    }
    Check(rc);
    for (size_t i = 0; i < f.track_ids.size(); ++i) if (status[i] > 0) r.tracks.at(f.track_ids[i]).estimated = false;
    return (int)removed;
  });
  // SelectGoodTracksForBundleAdjustment (sfm_wrapper.cc:28-44 -> select_good_tracks_for_bundle_adjustment.cc:263-325) over
  // thb_select_good_tracks_batch: views and tracks flattened in ascending id order (the order the C-ABI entry fixes where the
  // reference's unordered containers leave it open). Views of view_ids that are not estimated are skipped.
  sfm.def("SelectGoodTracksForBundleAdjustment", [](const Reconstruction& r, const std::unordered_set<ViewId>& view_ids, int long_track_length_threshold,
                                                  int image_grid_cell_size_pixels, int min_num_optimized_tracks_per_view) {
    Flat f;
    std::vector<ViewId> views;
    for (const auto& kv : r.views) if (kv.second.estimated) views.push_back(kv.first);
    std::sort(views.begin(), views.end());
    std::vector<uint8_t> cam_selected;
    for (ViewId v : views) {
      const View& view = r.views.at(v);
      f.view_index[v] = (int)f.view_ids.size(); f.view_ids.push_back(v);
      cam_selected.push_back(view_ids.count(v) ? 1 : 0);
      f.cam_ext.insert(f.cam_ext.end(), view.camera.ext, view.camera.ext + 6);
      Intrinsics* in = view.camera.intr.get();
      if (!f.group_index.count(in)) {
        f.group_index[in] = (int)f.groups.size(); f.groups.push_back(in);
        f.intr_model.push_back(in->model);
        f.intr.insert(f.intr.end(), in->params, in->params + THB_INTR_STRIDE);
      }
      f.cam_group.push_back(f.group_index[in]);
    }
    std::vector<TrackId> tracks;                                     // estimated tracks seen by a selected view (:118-141)
    for (const auto& kv : r.tracks) {
      if (!kv.second.estimated) continue;
      bool seen = false;
      for (ViewId v : kv.second.views) if (view_ids.count(v) && f.view_index.count(v)) { seen = true; break; }
      if (seen) tracks.push_back(kv.first);
    }
    std::sort(tracks.begin(), tracks.end());
    for (TrackId t : tracks) {
      const Track& tr = r.tracks.at(t);
      const int pi = (int)f.track_ids.size();
      f.track_ids.push_back(t);
      f.pts.insert(f.pts.end(), tr.point, tr.point + 4);
      std::vector<ViewId> obs(tr.views.begin(), tr.views.end());
      std::sort(obs.begin(), obs.end());
      for (ViewId v : obs) {
        auto vi = f.view_index.find(v);
        if (vi == f.view_index.end()) continue;                      // not estimated (:94-96)
        const Feature& feat = r.views.at(v).features.at(t);
        f.obs_cam.push_back(vi->second); f.obs_pt.push_back(pi);
        f.obs_xy.push_back(feat.point[0]); f.obs_xy.push_back(feat.point[1]);
      }
    }
    std::unordered_set<TrackId> out;
    if (f.track_ids.empty()) return std::make_tuple(true, out);
    ThbBaProblem p;
    std::memset(&p, 0, sizeof(p));
    p.num_cameras = (int)f.view_ids.size(); p.num_groups = (int)f.groups.size(); p.num_points = (int)f.track_ids.size();
    p.num_observations = (int)f.obs_cam.size(); p.memory_space = THB_MEM_HOST;
    p.cam_ext = f.cam_ext.data(); p.cam_group = f.cam_group.data(); p.intr = f.intr.data(); p.intr_model = f.intr_model.data();
    p.pts = f.pts.data(); p.obs_cam = f.obs_cam.data(); p.obs_pt = f.obs_pt.data(); p.obs_xy = f.obs_xy.data();
    std::vector<uint8_t> selected(f.track_ids.size(), 0);
    int32_t count = 0;
    int rc;
    {
      py::gil_scoped_release nogil;
      rc = thb_select_good_tracks_batch(&p, cam_selected.data(), long_track_length_threshold, image_grid_cell_size_pixels,
                                        min_num_optimized_tracks_per_view, selected.data(), &count, nullptr);
    }
    Check(rc);
    for (size_t i = 0; i < selected.size(); ++i) if (selected[i]) out.insert(f.track_ids[i]);
    return std::make_tuple(true, out);
  });
  sfm.def("BundleAdjustTrack", [](Reconstruction& r, const BundleAdjustmentOptions& o, TrackId t) {
    BundleAdjustmentSummary s = RunBa(o, {}, {t}, &r, true);   // no inner iterations (:267)
    UpdateInverseDepth({t}, &r);
    return s;
  });
  sfm.def("BundleAdjustTracks", [](Reconstruction& r, const BundleAdjustmentOptions& o, const std::vector<TrackId>& t) {
    BundleAdjustmentSummary s = RunBa(o, {}, t, &r, true);
    UpdateInverseDepth(t, &r);
    return s;
  });

  py::class_<RansacParameters>(solvers, "RansacParameters")
      .def(py::init<>())
      .def_readwrite("error_thresh", &RansacParameters::error_thresh).def_readwrite("failure_probability", &RansacParameters::failure_probability)
      .def_readwrite("min_inlier_ratio", &RansacParameters::min_inlier_ratio).def_readwrite("min_iterations", &RansacParameters::min_iterations)
      .def_readwrite("max_iterations", &RansacParameters::max_iterations).def_readwrite("use_mle", &RansacParameters::use_mle)
      .def_readwrite("use_Tdd_test", &RansacParameters::use_Tdd_test).def_readwrite("use_lo", &RansacParameters::use_lo)
      .def_readwrite("lo_start_iterations", &RansacParameters::lo_start_iterations).def_readwrite("seed", &RansacParameters::seed);
  py::class_<RansacSummary>(solvers, "RansacSummary")
      .def_readonly("inliers", &RansacSummary::inliers).def_readonly("num_input_data_points", &RansacSummary::num_input_data_points)
      .def_readonly("num_iterations", &RansacSummary::num_iterations).def_readonly("confidence", &RansacSummary::confidence)
      .def_readonly("num_lo_iterations", &RansacSummary::num_lo_iterations);
  py::class_<FeatureCorrespondence>(matching, "FeatureCorrespondence")
      .def(py::init<>())
      .def(py::init([](const Feature& a, const Feature& b) { FeatureCorrespondence c; c.feature1 = a; c.feature2 = b; return c; }))
      .def_readwrite("feature1", &FeatureCorrespondence::feature1).def_readwrite("feature2", &FeatureCorrespondence::feature2);
  py::class_<FeatureCorrespondence2D3D>(sfm, "FeatureCorrespondence2D3D")
      .def(py::init<>())
      .def(py::init([](const Vec& f, const Vec& w) { FeatureCorrespondence2D3D c; CopyVec(f, c.feature, 2, "feature"); CopyVec(w, c.world_point, 3, "world_point"); return c; }))
      .def_property("feature", [](const FeatureCorrespondence2D3D& c) { return MakeVec(c.feature, 2); }, [](FeatureCorrespondence2D3D& c, const Vec& v) { CopyVec(v, c.feature, 2, "feature"); })
      .def_property("world_point", [](const FeatureCorrespondence2D3D& c) { return MakeVec(c.world_point, 3); }, [](FeatureCorrespondence2D3D& c, const Vec& v) { CopyVec(v, c.world_point, 3, "world_point"); });
  py::class_<RelativePose>(sfm, "RelativePose")
      .def_property_readonly("essential_matrix", [](const RelativePose& p) { return MakeMat(p.E, 3, 3); })
      .def_property_readonly("rotation", [](const RelativePose& p) { return MakeMat(p.R, 3, 3); })
      .def_property_readonly("position", [](const RelativePose& p) { return MakeVec(p.p, 3); });
  py::class_<CalibratedAbsolutePose>(sfm, "CalibratedAbsolutePose")
      .def_property_readonly("rotation", [](const CalibratedAbsolutePose& p) { return MakeMat(p.R, 3, 3); })
      .def_property_readonly("position", [](const CalibratedAbsolutePose& p) { return MakeVec(p.p, 3); });

  // estimators_wrapper.cc:42-60, 99-115, 130-146: tuple(bool, Model, RansacSummary)
  sfm.def("EstimateRelativePose", [](const RansacParameters& q, int type, const std::vector<FeatureCorrespondence>& c) {
    ThbRelPoseResult res; RansacSummary sum; RelativePose pose;
    const bool ok = RunOne(thb_ransac_relpose_batch, q, type, Flatten(c), 4, &res, &sum);
    std::copy_n(res.essential_matrix, 9, pose.E); std::copy_n(res.rotation, 9, pose.R); std::copy_n(res.position, 3, pose.p);
    return py::make_tuple(ok, pose, sum);
  });
  sfm.def("EstimateHomography", [](const RansacParameters& q, int type, const std::vector<FeatureCorrespondence>& c) {
    ThbRelPoseResult res; RansacSummary sum;
    const bool ok = RunOne(thb_ransac_homography_batch, q, type, Flatten(c), 4, &res, &sum);
    return py::make_tuple(ok, MakeMat(res.essential_matrix, 3, 3), sum);
  });
  sfm.def("EstimateCalibratedAbsolutePose", [](const RansacParameters& q, int type, int pnp_type, const std::vector<FeatureCorrespondence2D3D>& c) {
    if (pnp_type != 0) throw std::runtime_error("only PnPType.KNEIP is implemented (DLS / SQPnP are not replayable, DESIGN.md)");
    std::vector<double> d(c.size() * 5);
    for (size_t i = 0; i < c.size(); ++i) { d[5 * i] = c[i].feature[0]; d[5 * i + 1] = c[i].feature[1]; std::copy_n(c[i].world_point, 3, &d[5 * i + 2]); }
    ThbRelPoseResult res; RansacSummary sum; CalibratedAbsolutePose pose;
    const bool ok = RunOne(thb_ransac_abspose_batch, q, type, d, 5, &res, &sum);
    std::copy_n(res.rotation, 9, pose.R); std::copy_n(res.position, 3, pose.p);
    return py::make_tuple(ok, pose, sum);
  });

  // sfm.cc:156-163, 545-568, 864-885, 1456-1467
  auto add_prior = [&](auto tag, const char* name) {
    using P = decltype(tag);
    constexpr int N = (int)(sizeof(P::value) / sizeof(double));
    py::class_<P>(sfm, name).def(py::init<>()).def_readwrite("is_set", &P::is_set)
        .def_property("value", [](const P& p) { return MakeVec(p.value, N); },
                      [](P& p, const Vec& v) { CopyVec(v, p.value, N, "prior value"); p.is_set = true; });
  };
  add_prior(Prior<1>{}, "Prior1d"); add_prior(Prior<2>{}, "Prior2d"); add_prior(Prior<3>{}, "Prior3d"); add_prior(Prior<4>{}, "Prior4d");
  py::class_<CameraIntrinsicsPrior>(sfm, "CameraIntrinsicsPrior")
      .def(py::init<>())
      .def_readwrite("image_width", &CameraIntrinsicsPrior::image_width).def_readwrite("image_height", &CameraIntrinsicsPrior::image_height)
      .def_readwrite("camera_intrinsics_model_type", &CameraIntrinsicsPrior::camera_intrinsics_model_type)
      .def_readwrite("focal_length", &CameraIntrinsicsPrior::focal_length).def_readwrite("principal_point", &CameraIntrinsicsPrior::principal_point)
      .def_readwrite("aspect_ratio", &CameraIntrinsicsPrior::aspect_ratio).def_readwrite("skew", &CameraIntrinsicsPrior::skew)
      .def_readwrite("radial_distortion", &CameraIntrinsicsPrior::radial_distortion)
      .def_readwrite("tangential_distortion", &CameraIntrinsicsPrior::tangential_distortion)
      .def_readwrite("position", &CameraIntrinsicsPrior::position).def_readwrite("orientation", &CameraIntrinsicsPrior::orientation)
      .def_readwrite("latitude", &CameraIntrinsicsPrior::latitude).def_readwrite("longitude", &CameraIntrinsicsPrior::longitude)
      .def_readwrite("altitude", &CameraIntrinsicsPrior::altitude);
  py::class_<TwoViewInfo>(sfm, "TwoViewInfo")
      .def(py::init<>())
      .def_readwrite("focal_length_1", &TwoViewInfo::focal_length_1).def_readwrite("focal_length_2", &TwoViewInfo::focal_length_2)
      .def_property("position_2", [](const TwoViewInfo& t) { return MakeVec(t.position_2, 3); }, [](TwoViewInfo& t, const Vec& v) { CopyVec(v, t.position_2, 3, "position_2"); })
      .def_property("rotation_2", [](const TwoViewInfo& t) { return MakeVec(t.rotation_2, 3); }, [](TwoViewInfo& t, const Vec& v) { CopyVec(v, t.rotation_2, 3, "rotation_2"); })
      .def_readwrite("num_verified_matches", &TwoViewInfo::num_verified_matches)
      .def_readwrite("num_homography_inliers", &TwoViewInfo::num_homography_inliers)
      .def_readwrite("visibility_score", &TwoViewInfo::visibility_score).def_readwrite("scale_estimate", &TwoViewInfo::scale_estimate);
  py::class_<EstimateTwoViewInfoOptions>(sfm, "EstimateTwoViewInfoOptions")
      .def(py::init<>())
      .def_readwrite("ransac_type", &EstimateTwoViewInfoOptions::ransac_type)
      .def_readwrite("max_sampson_error_pixels", &EstimateTwoViewInfoOptions::max_sampson_error_pixels)
      .def_readwrite("expected_ransac_confidence", &EstimateTwoViewInfoOptions::expected_ransac_confidence)
      .def_readwrite("min_ransac_iterations", &EstimateTwoViewInfoOptions::min_ransac_iterations)
      .def_readwrite("max_ransac_iterations", &EstimateTwoViewInfoOptions::max_ransac_iterations)
      .def_readwrite("use_mle", &EstimateTwoViewInfoOptions::use_mle).def_readwrite("use_lo", &EstimateTwoViewInfoOptions::use_lo)
      .def_readwrite("lo_start_iterations", &EstimateTwoViewInfoOptions::lo_start_iterations)
      .def_readwrite("min_focal_length", &EstimateTwoViewInfoOptions::min_focal_length)
      .def_readwrite("max_focal_length", &EstimateTwoViewInfoOptions::max_focal_length)
      .def_readwrite("seed", &EstimateTwoViewInfoOptions::seed);
  // sfm_wrapper.cc:12-26: tuple(bool, TwoViewInfo, inlier indices). Calibrated branch (estimate_twoview_info.cc:133-191):
  // normalise by the intrinsics, resolution-scaled Sampson threshold, EstimateRelativePose on the device, fill TwoViewInfo.
  py::class_<VerificationOptions>(sfm, "TwoViewMatchGeometricVerificationOptions")  // two_view_match_geometric_verification.h:53-92
      .def(py::init<>())
      .def_readwrite("estimate_twoview_info_options", &VerificationOptions::estimate_twoview_info_options)
      .def_readwrite("min_num_inlier_matches", &VerificationOptions::min_num_inlier_matches)
      .def_readwrite("guided_matching", &VerificationOptions::guided_matching)
      .def_readwrite("bundle_adjustment", &VerificationOptions::bundle_adjustment)
      .def_readwrite("triangulation_max_reprojection_error", &VerificationOptions::triangulation_max_reprojection_error)
      .def_readwrite("min_triangulation_angle_degrees", &VerificationOptions::min_triangulation_angle_degrees)
      .def_readwrite("final_max_reprojection_error", &VerificationOptions::final_max_reprojection_error);
  sfm.def("EstimateTwoViewInfo", [](const EstimateTwoViewInfoOptions& o, const CameraIntrinsicsPrior& i1, const CameraIntrinsicsPrior& i2,
                                    const std::vector<FeatureCorrespondence>& c) -> py::tuple {
    std::vector<std::vector<FeatureCorrespondence>> one(1, c);
    VerificationOptions vo; vo.estimate_twoview_info_options = o;
    return py::tuple(RunTwoViewBatch(vo, {i1}, {i2}, one, {}, false)[0]);
  });
  // ADDITIVE: the pair loop of the pipelines (pytests/sfm_pipeline.py:232-238) as ONE call - all pairs of a view graph
  sfm.def("EstimateTwoViewInfoBatch", [](const EstimateTwoViewInfoOptions& o, const std::vector<CameraIntrinsicsPrior>& i1,
                                         const std::vector<CameraIntrinsicsPrior>& i2, const std::vector<std::vector<FeatureCorrespondence>>& c,
                                         const std::vector<uint32_t>& seeds) {
    VerificationOptions vo; vo.estimate_twoview_info_options = o;
    return RunTwoViewBatch(vo, i1, i2, c, seeds, false);
  }, py::arg("options"), py::arg("intrinsics1"), py::arg("intrinsics2"), py::arg("correspondences"), py::arg("seeds") = std::vector<uint32_t>());
  // TwoViewMatchGeometricVerification::VerifyMatches (C++ only upstream, two_view_match_geometric_verification.cc:114-183) for a
  // batch of pairs: tuple(success, TwoViewInfo, indices of the verified matches) per pair
  sfm.def("VerifyTwoViewMatchesBatch", [](const VerificationOptions& vo, const std::vector<CameraIntrinsicsPrior>& i1,
                                          const std::vector<CameraIntrinsicsPrior>& i2, const std::vector<std::vector<FeatureCorrespondence>>& c,
                                          const std::vector<uint32_t>& seeds) {
    return RunTwoViewBatch(vo, i1, i2, c, seeds, true);
  }, py::arg("options"), py::arg("intrinsics1"), py::arg("intrinsics2"), py::arg("correspondences"), py::arg("seeds") = std::vector<uint32_t>());

  // sfm.cc:1095-1135
  py::class_<TrackEstimatorOptions>(sfm, "TrackEstimatorOptions")
      .def(py::init<>())
      .def_readwrite("num_threads", &TrackEstimatorOptions::num_threads)
      .def_readwrite("max_acceptable_reprojection_error_pixels", &TrackEstimatorOptions::max_acceptable_reprojection_error_pixels)
      .def_readwrite("min_triangulation_angle_degrees", &TrackEstimatorOptions::min_triangulation_angle_degrees)
      .def_readwrite("bundle_adjustment", &TrackEstimatorOptions::bundle_adjustment)
      .def_readwrite("ba_options", &TrackEstimatorOptions::ba_options)
      .def_readwrite("multithreaded_step_size", &TrackEstimatorOptions::multithreaded_step_size)
      .def_readwrite("triangulation_method", &TrackEstimatorOptions::triangulation_method);
  py::class_<TrackEstimatorSummary>(sfm, "TrackEstimatorSummary")
      .def_readonly("input_num_estimated_tracks", &TrackEstimatorSummary::input_num_estimated_tracks)
      .def_readonly("num_triangulation_attempts", &TrackEstimatorSummary::num_triangulation_attempts)
      .def_readonly("estimated_tracks", &TrackEstimatorSummary::estimated_tracks)
      .def_readonly("num_bad_angles", &TrackEstimatorSummary::num_bad_angles)
      .def_readonly("num_failed_triangulations", &TrackEstimatorSummary::num_failed_triangulations)
      .def_readonly("num_bad_reprojections", &TrackEstimatorSummary::num_bad_reprojections);
  py::class_<TrackEstimator>(sfm, "TrackEstimator")
      .def(py::init<const TrackEstimatorOptions&, Reconstruction*>(), py::keep_alive<1, 3>())
      .def("EstimateAllTracks", &TrackEstimator::EstimateAllTracks)
      .def("EstimateTracks", [](TrackEstimator& e, const std::unordered_set<TrackId>& ids) {
        return e.EstimateTracks(std::vector<TrackId>(ids.begin(), ids.end()));
      });

  // triangulation_wrapper.h:15-17, sfm.cc:854: (success, homogeneous point). One track per call here; the batched C-ABI entry
  // (all tracks of a reconstruction in one launch) is what a TrackEstimator replacement calls.
  sfm.def("TriangulateMidpoint", [](const std::vector<Vec>& origins, const std::vector<Vec>& directions) {
    if (origins.size() < 2 || origins.size() != directions.size()) throw std::invalid_argument("TriangulateMidpoint: need >= 2 rays, as many origins as directions");
    std::vector<double> o(origins.size() * 3), d(origins.size() * 3);
    for (size_t i = 0; i < origins.size(); ++i) { CopyVec(origins[i], &o[3 * i], 3, "ray origin"); CopyVec(directions[i], &d[3 * i], 3, "ray direction"); }
    const int64_t off[2] = {0, (int64_t)origins.size()};
    double X[4] = {0, 0, 0, 0};
    uint8_t ok = 0;
    Check(thb_triangulate_midpoint_batch(o.data(), d.data(), off, 1, THB_MEM_HOST, X, &ok, nullptr));
    return py::make_tuple(ok != 0, MakeVec(X, 4));
  });

  // pose_wrapper.cc:166-173, 210-217, 369-377 and PoseFromThreePoints
  sfm.def("FivePointRelativePose", [](const std::vector<Vec>& a, const std::vector<Vec>& b) {
    if (a.size() != 5 || b.size() != 5) throw std::runtime_error("only the minimal 5-point case is implemented");
    const std::vector<double> x1 = Points2(a, 5), x2 = Points2(b, 5);
    double E[90]; int32_t n = 0;
    Check(thb_five_point_relative_pose(x1.data(), x2.data(), 1, E, &n, nullptr));
    std::vector<Vec> out;
    for (int k = 0; k < n; ++k) out.push_back(MakeMat(E + 9 * k, 3, 3));
    return py::make_tuple(n > 0, out);
  });
  sfm.def("FourPointHomography", [](const std::vector<Vec>& a, const std::vector<Vec>& b) {
    if (a.size() != 4 || b.size() != 4) throw std::runtime_error("only the minimal 4-point case is implemented");
    const std::vector<double> x1 = Points2(a, 4), x2 = Points2(b, 4);
    double c[16], H[9]; int32_t ok = 0;
    for (int i = 0; i < 4; ++i) { c[4 * i] = x1[2 * i]; c[4 * i + 1] = x1[2 * i + 1]; c[4 * i + 2] = x2[2 * i]; c[4 * i + 3] = x2[2 * i + 1]; }
    Check(thb_four_point_homography(c, 1, H, &ok, nullptr));
    return py::make_tuple(ok != 0, MakeMat(H, 3, 3));
  });
  sfm.def("SevenPointFundamentalMatrix", [](const std::vector<Vec>& a, const std::vector<Vec>& b) {
    if (a.size() != 7 || b.size() != 7) throw std::invalid_argument("exactly 7 correspondences are required");   // CHECK_EQ in the reference
    const std::vector<double> x1 = Points2(a, 7), x2 = Points2(b, 7);
    double c[28], F[27]; int32_t n = 0;
    for (int i = 0; i < 7; ++i) { c[4 * i] = x1[2 * i]; c[4 * i + 1] = x1[2 * i + 1]; c[4 * i + 2] = x2[2 * i]; c[4 * i + 3] = x2[2 * i + 1]; }
    Check(thb_seven_point_fundamental_matrix(c, 1, F, &n, nullptr));
    std::vector<Vec> out;
    for (int k = 0; k < n; ++k) out.push_back(MakeMat(F + 9 * k, 3, 3));
    return py::make_tuple(n > 0, out);
  });
  sfm.def("PoseFromThreePoints", [](const std::vector<Vec>& feats, const std::vector<Vec>& world) {
    if (feats.size() != 3 || world.size() != 3) throw std::invalid_argument("exactly 3 correspondences are required");
    const std::vector<double> f = Points2(feats, 3);
    double w[9], R[36], t[12]; int32_t n = 0;
    for (int i = 0; i < 3; ++i) CopyVec(world[i], w + 3 * i, 3, "world point");
    Check(thb_p3p(f.data(), w, 1, R, t, &n, nullptr));
    std::vector<Vec> Rs, ts;
    for (int k = 0; k < n; ++k) { Rs.push_back(MakeMat(R + 9 * k, 3, 3)); ts.push_back(MakeVec(t + 3 * k, 3)); }
    return py::make_tuple(n > 0, Rs, ts);
  });
}
