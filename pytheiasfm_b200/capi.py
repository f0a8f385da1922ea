"""ctypes mirror of include/theia_b200.h and the loader of the CUDA C-ABI library.

The product path is libtheia_b200.so (hand-written sm_100a CUDA behind `extern "C"`).
There is no CPU fallback: if the library is missing, or no B200-class device is visible,
every compute entry point fails loudly (LibraryNotBuilt / THB_E_NO_DEVICE).
"""
import ctypes as C
import os

import numpy as np

THB_OK = 0
THB_E_INVALID_ARGUMENT = -1
THB_E_UNSUPPORTED = -2
THB_E_CUDA = -3
THB_E_NO_DEVICE = -4
THB_E_NUMERICAL = -5

THB_MEM_HOST = 0
THB_MEM_DEVICE = 1

THB_INTR_STRIDE = 10
THB_MAX_ITER_LOG = 256

# sfm/camera/camera_intrinsics_model_type.h:46-56
MODEL_PINHOLE = 0
MODEL_FISHEYE = 2
MODEL_FOV = 3
MODEL_DIVISION_UNDISTORTION = 4
MODEL_DOUBLE_SPHERE = 5
MODEL_EXTENDED_UNIFIED = 6
MODEL_NUM_PARAMS = {0: 7, 2: 9, 3: 5, 4: 5, 5: 7, 6: 7}

# sfm/bundle_adjustment/create_loss_function.h:51-59
LOSS_TRIVIAL, LOSS_HUBER, LOSS_SOFTLONE, LOSS_CAUCHY, LOSS_ARCTAN, LOSS_TUKEY, LOSS_TRUNCATED = range(7)

SOLVER_SCHUR_CHOLESKY = 0
SOLVER_SCHUR_PCG = 1

CAM_CONST_POSITION = 1
CAM_CONST_ORIENTATION = 2

TERM_CONVERGENCE, TERM_NO_CONVERGENCE, TERM_FAILURE = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)


class ThbBaProblem(C.Structure):
    _fields_ = [
        ("num_cameras", C.c_int32), ("num_groups", C.c_int32), ("num_points", C.c_int32),
        ("num_observations", C.c_int32), ("memory_space", C.c_int32), ("reserved0", C.c_int32),
        ("cam_ext", C.c_void_p), ("cam_const", C.c_void_p), ("cam_group", C.c_void_p),
        ("intr", C.c_void_p), ("intr_model", C.c_void_p), ("intr_const", C.c_void_p),
        ("pts", C.c_void_p), ("pt_const", C.c_void_p),
        ("obs_cam", C.c_void_p), ("obs_pt", C.c_void_p), ("obs_xy", C.c_void_p),
        ("obs_sqrt_info", C.c_void_p),
        ("cam_has_position_prior", C.c_void_p), ("cam_position_prior", C.c_void_p), ("cam_position_prior_sqrt_info", C.c_void_p),
        ("cam_has_gravity_prior", C.c_void_p), ("cam_gravity_prior", C.c_void_p), ("cam_gravity_prior_sqrt_info", C.c_void_p),
        ("cam_has_orientation_prior", C.c_void_p), ("cam_orientation_prior", C.c_void_p), ("cam_orientation_prior_sqrt_info", C.c_void_p),
    ]


class ThbBaOptions(C.Structure):
    _fields_ = [
        ("loss_function_type", C.c_int32), ("linear_solver", C.c_int32),
        ("use_homogeneous_point_parametrization", C.c_int32), ("use_inner_iterations", C.c_int32),
        ("max_num_iterations", C.c_int32), ("jacobi_scaling", C.c_int32), ("verbose", C.c_int32),
        ("max_num_consecutive_invalid_steps", C.c_int32),
        ("robust_loss_width", C.c_double), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("max_trust_region_radius", C.c_double), ("initial_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("max_solver_time_in_seconds", C.c_double), ("pcg_eta", C.c_double),
        ("pcg_max_iterations", C.c_int32), ("reserved0", C.c_int32),
    ]


class ThbBaSummary(C.Structure):
    _fields_ = [
        ("success", C.c_int32), ("termination_type", C.c_int32), ("num_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32), ("num_jacobian_evaluations", C.c_int32),
        ("num_cost_evaluations", C.c_int32), ("num_linear_solves", C.c_int32),
        ("gpu_launches", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("setup_time_in_seconds", C.c_double), ("solve_time_in_seconds", C.c_double),
        ("ms_jacobian", C.c_double), ("ms_normal", C.c_double), ("ms_solve", C.c_double),
        ("ms_update", C.c_double),
        ("iter_log_count", C.c_int32), ("num_linear_solver_iterations", C.c_int32),
        ("iter_cost", C.c_double * THB_MAX_ITER_LOG), ("iter_radius", C.c_double * THB_MAX_ITER_LOG),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith(("iter_cost", "iter_radius", "reserved"))}
        d["iter_cost"] = list(self.iter_cost[: self.iter_log_count])
        d["iter_radius"] = list(self.iter_radius[: self.iter_log_count])
        return d


class ThbRansacParams(C.Structure):
    _fields_ = [
        ("error_thresh", C.c_double), ("failure_probability", C.c_double), ("min_inlier_ratio", C.c_double),
        ("min_iterations", C.c_int32), ("max_iterations", C.c_int32), ("use_mle", C.c_int32), ("use_lo", C.c_int32),
        ("lo_start_iterations", C.c_int32), ("ransac_type", C.c_int32), ("use_tdd_test", C.c_int32), ("reserved0", C.c_int32),
    ]


class ThbRansacStats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("iterations", C.c_uint64), ("samples_solved", C.c_uint64), ("models_scored", C.c_uint64),
                ("data_scored", C.c_uint64), ("reserved0", C.c_uint64), ("cycles_draw", C.c_uint64), ("cycles_solve", C.c_uint64),
                ("cycles_score", C.c_uint64), ("cycles_scan", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved0"}


class ThbPairBatch(C.Structure):
    _fields_ = [("num_pairs", C.c_int32), ("memory_space", C.c_int32), ("pair_offset", C.c_void_p),
                ("corr", C.c_void_p), ("seed", C.c_void_p)]


class ThbRelPoseResult(C.Structure):
    _fields_ = [
        ("success", C.c_int32), ("num_inliers", C.c_int32), ("num_iterations", C.c_int32),
        ("num_input_data_points", C.c_int32), ("confidence", C.c_double), ("best_cost", C.c_double),
        ("essential_matrix", C.c_double * 9), ("rotation", C.c_double * 9), ("position", C.c_double * 3),
        ("num_lo_iterations", C.c_int32), ("reserved0", C.c_int32),
    ]


class ThbViewIntrinsics(C.Structure):
    _fields_ = [("model", C.c_int32), ("image_width", C.c_int32), ("image_height", C.c_int32), ("focal_length_is_set", C.c_int32),
                ("params", C.c_double * THB_INTR_STRIDE)]


class ThbTwoViewOptions(C.Structure):
    _fields_ = [("max_sampson_error_pixels", C.c_double), ("expected_ransac_confidence", C.c_double),
                ("min_ransac_iterations", C.c_int32), ("max_ransac_iterations", C.c_int32), ("use_mle", C.c_int32), ("use_lo", C.c_int32),
                ("lo_start_iterations", C.c_int32), ("ransac_type", C.c_int32), ("min_num_inlier_matches", C.c_int32),
                ("bundle_adjustment", C.c_int32), ("triangulation_max_reprojection_error", C.c_double),
                ("min_triangulation_angle_degrees", C.c_double), ("final_max_reprojection_error", C.c_double)]


VIEW_INTRINSICS_DTYPE = np.dtype([("model", np.int32), ("image_width", np.int32), ("image_height", np.int32), ("focal_length_is_set", np.int32),
                                  ("params", np.float64, (THB_INTR_STRIDE,))])
TWO_VIEW_INFO_DTYPE = np.dtype([("success", np.int32), ("num_verified_matches", np.int32), ("num_homography_inliers", np.int32),
                                ("visibility_score", np.int32), ("num_ransac_iterations", np.int32), ("num_triangulated", np.int32),
                                ("ba_iterations", np.int32), ("reserved0", np.int32), ("focal_length_1", np.float64), ("focal_length_2", np.float64),
                                ("position_2", np.float64, (3,)), ("rotation_2", np.float64, (3,)), ("ba_initial_cost", np.float64),
                                ("ba_final_cost", np.float64)])
assert VIEW_INTRINSICS_DTYPE.itemsize == C.sizeof(ThbViewIntrinsics)

TRACK_SKIPPED, TRACK_ESTIMATED, TRACK_BAD_ANGLE, TRACK_FAILED_TRIANGULATION, TRACK_BA_FAILED, TRACK_BAD_REPROJECTION = -1, 0, 1, 2, 3, 4


class ThbTrackEstimatorOptions(C.Structure):
    """TrackEstimator::Options (estimate_track.h:59-84) with its defaults."""
    _fields_ = [("max_acceptable_reprojection_error_pixels", C.c_double), ("min_triangulation_angle_degrees", C.c_double),
                ("bundle_adjustment", C.c_int32), ("reserved0", C.c_int32)]

    def __init__(self):
        super().__init__(5.0, 3.0, 1, 0)


TRACK_BA_DTYPE = np.dtype([("initial_cost", np.float64), ("final_cost", np.float64), ("num_iterations", np.int32),
                           ("termination_type", np.int32)])
RELPOSE_DTYPE = np.dtype([("success", np.int32), ("num_inliers", np.int32), ("num_iterations", np.int32),
                          ("num_input_data_points", np.int32), ("confidence", np.float64), ("best_cost", np.float64),
                          ("essential_matrix", np.float64, (3, 3)), ("rotation", np.float64, (3, 3)),
                          ("position", np.float64, (3,)), ("num_lo_iterations", np.int32), ("reserved0", np.int32)])
assert RELPOSE_DTYPE.itemsize == C.sizeof(ThbRelPoseResult)


class HostPairBatch:
    """Correspondences of a batch of image pairs in the C-ABI layout (host memory)."""

    def __init__(self, corr_list, seeds, width=4):
        self.width = width  # 4: FeatureCorrespondence (x1,y1,x2,y2); 5: FeatureCorrespondence2D3D (x,y,X,Y,Z)
        self.num_pairs = len(corr_list)
        sizes = [len(c) for c in corr_list]
        self.pair_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.corr = (np.ascontiguousarray(np.concatenate(corr_list, 0), dtype=np.float64) if self.num_pairs and self.pair_offset[-1] > 0
                     else np.zeros((0, width)))
        self.seed = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert self.corr.shape == (self.pair_offset[-1], width) and self.seed.shape == (self.num_pairs,)

    def struct(self):
        b = ThbPairBatch()
        b.num_pairs = self.num_pairs
        b.memory_space = THB_MEM_HOST
        b.pair_offset = _ptr(self.pair_offset)
        b.corr = _ptr(self.corr)
        b.seed = _ptr(self.seed)
        return b


class LibraryNotBuilt(RuntimeError):
    pass


class ThbError(RuntimeError):
    def __init__(self, code, text=""):
        super().__init__("theia_b200 error %d %s" % (code, text))
        self.code = code


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libtheia_b200.so")
_lib = None


def load_library():
    """Load libtheia_b200.so (built in-tree by __graft_entry__.build()). No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuilt(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback for this path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.thb_version.restype = C.c_int
    lib.thb_last_error.restype = C.c_char_p
    lib.thb_device_count.restype = C.c_int
    lib.thb_ba_default_options.argtypes = [C.POINTER(ThbBaOptions)]
    lib.thb_ba_default_options.restype = None
    lib.thb_ba_solve.argtypes = [C.POINTER(ThbBaProblem), C.POINTER(ThbBaOptions), C.POINTER(ThbBaSummary), C.c_void_p]
    lib.thb_ba_solve.restype = C.c_int
    lib.thb_ba_create.argtypes = [C.POINTER(ThbBaProblem), C.POINTER(ThbBaOptions), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.thb_ba_create.restype = C.c_int
    lib.thb_ba_iterate.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
    lib.thb_ba_iterate.restype = C.c_int
    lib.thb_ba_finish.argtypes = [C.c_void_p, C.POINTER(ThbBaSummary)]
    lib.thb_ba_finish.restype = C.c_int
    lib.thb_ba_evaluate.argtypes = [C.POINTER(ThbBaProblem)] + [C.c_void_p] * 6
    lib.thb_ba_evaluate.restype = C.c_int
    lib.thb_ba_time_jacobian.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    lib.thb_ba_time_jacobian.restype = C.c_int
    lib.thb_ransac_default_params.argtypes = [C.POINTER(ThbRansacParams)]
    lib.thb_ransac_default_params.restype = None
    lib.thb_ransac_relpose_batch.argtypes = [C.POINTER(ThbPairBatch), C.POINTER(ThbRansacParams), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_ransac_relpose_batch.restype = C.c_int
    for name in ("thb_ransac_abspose_batch", "thb_ransac_homography_batch"):
        getattr(lib, name).argtypes = [C.POINTER(ThbPairBatch), C.POINTER(ThbRansacParams), C.c_void_p, C.c_void_p, C.c_void_p]
        getattr(lib, name).restype = C.c_int
    lib.thb_ransac_last_stats.argtypes = [C.POINTER(ThbRansacStats)]
    lib.thb_ransac_last_stats.restype = C.c_int
    lib.thb_pack_inlier_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.thb_pack_inlier_masks.restype = C.c_int
    lib.thb_fp64_peak_tflops.argtypes = [C.c_int32, C.POINTER(C.c_double), C.c_void_p]
    lib.thb_fp64_peak_tflops.restype = C.c_int
    lib.thb_set_outlier_tracks_batch.argtypes = [C.POINTER(ThbBaProblem), C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    lib.thb_set_outlier_tracks_batch.restype = C.c_int
    lib.thb_ba_covariance.argtypes = [C.POINTER(ThbBaProblem), C.POINTER(ThbBaOptions), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_ba_covariance.restype = C.c_int
    lib.thb_select_good_tracks_batch.argtypes = [C.POINTER(ThbBaProblem), C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                 C.POINTER(C.c_int32), C.c_void_p]
    lib.thb_select_good_tracks_batch.restype = C.c_int
    lib.thb_two_view_default_options.argtypes = [C.POINTER(ThbTwoViewOptions)]
    lib.thb_two_view_default_options.restype = None
    for name in ("thb_estimate_two_view_info_batch", "thb_verify_two_view_matches_batch"):
        getattr(lib, name).argtypes = [C.POINTER(ThbPairBatch), C.c_void_p, C.c_void_p, C.POINTER(ThbTwoViewOptions), C.c_void_p, C.c_void_p, C.c_void_p]
        getattr(lib, name).restype = C.c_int
    lib.thb_p3p.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_p3p.restype = C.c_int
    lib.thb_ba_tracks_batch.argtypes = [C.POINTER(ThbBaProblem), C.POINTER(ThbBaOptions), C.c_void_p, C.c_void_p]
    lib.thb_ba_tracks_batch.restype = C.c_int
    lib.thb_estimate_tracks_batch.argtypes = [C.POINTER(ThbBaProblem), C.c_void_p, C.POINTER(ThbTrackEstimatorOptions), C.POINTER(ThbBaOptions),
                                              C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_estimate_tracks_batch.restype = C.c_int
    lib.thb_triangulate_midpoint_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_triangulate_midpoint_batch.restype = C.c_int
    lib.thb_four_point_homography.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_four_point_homography.restype = C.c_int
    lib.thb_seven_point_fundamental_matrix.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_seven_point_fundamental_matrix.restype = C.c_int
    lib.thb_five_point_relative_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.thb_five_point_relative_pose.restype = C.c_int
    lib.thb_dense_spd_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.thb_dense_spd_solve.restype = C.c_int
    lib.thb_dense_spd_time.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
    lib.thb_dense_spd_time.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != THB_OK:
        text = ""
        if _lib is not None:
            try:
                text = _lib.thb_last_error().decode()
            except Exception:  # noqa: BLE001
                pass
        raise ThbError(rc, text)


def default_options(lib=None):
    """BundleAdjustmentOptions defaults (bundle_adjustment.h:87-167) with inner iterations off."""
    o = ThbBaOptions()
    (lib or load_library()).thb_ba_default_options(C.byref(o))
    return o


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostBaProblem:
    """Owns contiguous numpy arrays in the C-ABI layout and the ThbBaProblem that views them."""

    FIELDS = [
        ("cam_ext", np.float64), ("cam_const", np.uint8), ("cam_group", np.int32),
        ("intr", np.float64), ("intr_model", np.int32), ("intr_const", np.uint16),
        ("pts", np.float64), ("pt_const", np.uint8),
        ("obs_cam", np.int32), ("obs_pt", np.int32), ("obs_xy", np.float64), ("obs_sqrt_info", np.float64),
        ("cam_has_position_prior", np.uint8), ("cam_position_prior", np.float64), ("cam_position_prior_sqrt_info", np.float64),
        ("cam_has_gravity_prior", np.uint8), ("cam_gravity_prior", np.float64), ("cam_gravity_prior_sqrt_info", np.float64),
        ("cam_has_orientation_prior", np.uint8), ("cam_orientation_prior", np.float64), ("cam_orientation_prior_sqrt_info", np.float64),
    ]

    def __init__(self, arrays):
        self.a = {}
        for name, dt in self.FIELDS:
            v = arrays.get(name)
            self.a[name] = None if v is None else np.ascontiguousarray(v, dtype=dt).copy()
        a = self.a
        self.num_cameras = a["cam_ext"].shape[0]
        self.num_groups = a["intr"].shape[0]
        self.num_points = a["pts"].shape[0]
        self.num_observations = a["obs_cam"].shape[0]
        assert a["cam_ext"].shape == (self.num_cameras, 6)
        assert a["intr"].shape == (self.num_groups, THB_INTR_STRIDE)
        assert a["pts"].shape == (self.num_points, 4)
        assert a["obs_xy"].shape == (self.num_observations, 2)

    def struct(self):
        p = ThbBaProblem()
        p.num_cameras, p.num_groups = self.num_cameras, self.num_groups
        p.num_points, p.num_observations = self.num_points, self.num_observations
        p.memory_space = THB_MEM_HOST
        for name, _ in self.FIELDS:
            setattr(p, name, _ptr(self.a[name]))
        return p

    def copy(self):
        return HostBaProblem(self.a)
