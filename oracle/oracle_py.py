"""ctypes loader of oracle/_build/liboracle.so (ORACLE — test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from pytheiasfm_b200 import capi

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.oracle_ba_default_options.argtypes = [C.POINTER(capi.ThbBaOptions)]
        lib.oracle_ba_solve.argtypes = [C.POINTER(capi.ThbBaProblem), C.POINTER(capi.ThbBaOptions), C.POINTER(capi.ThbBaSummary)]
        lib.oracle_ba_evaluate.argtypes = [C.POINTER(capi.ThbBaProblem)] + [C.c_void_p] * 5
        lib.oracle_ba_cost.argtypes = [C.POINTER(capi.ThbBaProblem), C.POINTER(capi.ThbBaOptions), C.POINTER(C.c_double)]
        lib.oracle_sphere_plus.argtypes = [C.c_void_p] * 3
        lib.oracle_sphere_plus_jacobian.argtypes = [C.c_void_p] * 2
        lib.oracle_ransac_default_params.argtypes = [C.POINTER(capi.ThbRansacParams)]
        lib.oracle_five_point.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_ransac_relpose_batch.argtypes = [C.POINTER(capi.ThbPairBatch), C.POINTER(capi.ThbRansacParams), C.c_void_p, C.c_void_p, C.c_int32]
        for name in ("oracle_ransac_abspose_batch", "oracle_ransac_homography_batch"):
            getattr(lib, name).argtypes = [C.POINTER(capi.ThbPairBatch), C.POINTER(capi.ThbRansacParams), C.c_void_p, C.c_void_p, C.c_int32]
        lib.oracle_p3p.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_triangulate_midpoint_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_four_point_homography.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_seven_point_fundamental.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_poly_roots.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_jacobi_svd3.argtypes = [C.c_void_p] * 4
        lib.oracle_eigen10.argtypes = [C.c_void_p] * 4
        lib.oracle_fullpivlu_kernel_5x9.argtypes = [C.c_void_p] * 2
        lib.oracle_fullpivlu_solve10.argtypes = [C.c_void_p] * 3
        lib.oracle_sampson.argtypes = [C.c_void_p] * 2
        lib.oracle_sampson.restype = C.c_double
        lib.oracle_best_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def set_num_threads(n):
    """Sets the OpenMP team size of the oracle (0: leave as is); returns the team size a parallel region actually gets."""
    lib = load()
    lib.oracle_set_num_threads.argtypes = [C.c_int]
    lib.oracle_set_num_threads.restype = C.c_int
    return int(lib.oracle_set_num_threads(int(n)))


def default_options():
    o = capi.ThbBaOptions()
    load().oracle_ba_default_options(C.byref(o))
    return o


def ba_solve(prob, opts):
    """Runs the oracle LM on a HostBaProblem IN PLACE; returns the summary dict."""
    s = capi.ThbBaSummary()
    p = prob.struct()
    rc = load().oracle_ba_solve(C.byref(p), C.byref(opts), C.byref(s))
    d = s.as_dict()
    d["rc"] = rc
    return d


def ba_evaluate(prob):
    n = prob.num_observations
    r = np.zeros((n, 2)); jc = np.zeros((n, 2, 6)); ji = np.zeros((n, 2, capi.THB_INTR_STRIDE)); jp = np.zeros((n, 2, 4))
    ok = np.zeros(n, np.uint8)
    p = prob.struct()
    rc = load().oracle_ba_evaluate(C.byref(p), *[a.ctypes.data_as(C.c_void_p) for a in (r, jc, ji, jp, ok)])
    assert rc == 0, rc
    return r, jc, ji, jp, ok


def ba_cost(prob, opts):
    c = C.c_double(0.0)
    p = prob.struct()
    rc = load().oracle_ba_cost(C.byref(p), C.byref(opts), C.byref(c))
    return rc, c.value


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def ransac_default_params():
    p = capi.ThbRansacParams()
    load().oracle_ransac_default_params(C.byref(p))
    return p


def five_point(x1, x2):
    """x1, x2: [count, 5, 2] -> (E [count, 10, 3, 3], num_solutions [count])"""
    x1 = np.ascontiguousarray(x1, np.float64); x2 = np.ascontiguousarray(x2, np.float64)
    count = x1.shape[0]
    E = np.zeros((count, 10, 3, 3)); n = np.zeros(count, np.int32)
    load().oracle_five_point(_vp(x1), _vp(x2), count, _vp(E), _vp(n))
    return E, n


def ransac_batch(kind, batch, params, threads=0):
    """kind in {relpose, abspose, homography}. Returns (rc, results structured array [num_pairs], inlier mask [total])."""
    res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE)
    mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
    b = batch.struct()
    rc = getattr(load(), "oracle_ransac_%s_batch" % kind)(C.byref(b), C.byref(params), _vp(res), _vp(mask), threads)
    return rc, res, mask


def ransac_relpose_batch(batch, params, threads=0):
    return ransac_batch("relpose", batch, params, threads)


def two_view_batch(batch, intr1, intr2, options, verify):
    """EstimateTwoViewInfo (verify = False) / VerifyMatches (verify = True) for every pair; returns (rc, info, mask)."""
    lib = load()
    lib.oracle_two_view_batch.argtypes = [C.POINTER(capi.ThbPairBatch), C.c_void_p, C.c_void_p, C.POINTER(capi.ThbTwoViewOptions), C.c_void_p, C.c_void_p, C.c_int32]
    info = np.zeros(batch.num_pairs, capi.TWO_VIEW_INFO_DTYPE)
    mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
    b = batch.struct()
    rc = lib.oracle_two_view_batch(C.byref(b), _vp(intr1), _vp(intr2), C.byref(options), _vp(info), _vp(mask), int(verify))
    return rc, info, mask


def set_outlier_tracks(prob, max_err, min_angle_deg):
    lib = load()
    lib.oracle_set_outlier_tracks.argtypes = [C.POINTER(capi.ThbBaProblem), C.c_double, C.c_double, C.c_void_p]
    status = np.zeros(prob.num_points, np.int32)
    p = prob.struct()
    removed = lib.oracle_set_outlier_tracks(C.byref(p), max_err, min_angle_deg, _vp(status))
    return removed, status


def ba_covariance(prob, options):
    """(rc, cam_cov [nc,6,6], cam_ok [nc], pt_cov [np,3,3], pt_ok [np]); see thb_ba_covariance."""
    lib = load()
    lib.oracle_ba_covariance.argtypes = [C.POINTER(capi.ThbBaProblem), C.POINTER(capi.ThbBaOptions)] + [C.c_void_p] * 4
    cc = np.zeros((prob.num_cameras, 6, 6)); co = np.zeros(prob.num_cameras, np.uint8)
    pc = np.zeros((prob.num_points, 3, 3)); po = np.zeros(prob.num_points, np.uint8)
    p = prob.struct()
    rc = lib.oracle_ba_covariance(C.byref(p), C.byref(options), _vp(cc), _vp(co), _vp(pc), _vp(po))
    return rc, cc, co, pc, po


def select_good_tracks(prob, long_thr, cell_size, min_per_view, cam_selected=None, selected=None):
    lib = load()
    lib.oracle_select_good_tracks.argtypes = [C.POINTER(capi.ThbBaProblem), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    sel = np.zeros(prob.num_points, np.uint8) if selected is None else np.ascontiguousarray(selected, np.uint8).copy()
    cs = None if cam_selected is None else np.ascontiguousarray(cam_selected, np.uint8)
    p = prob.struct()
    n = lib.oracle_select_good_tracks(C.byref(p), None if cs is None else _vp(cs), long_thr, cell_size, min_per_view, _vp(sel))
    return n, sel


def p3p(feat, world):
    """feat [count,3,2], world [count,3,3] -> (R [count,4,3,3], t [count,4,3], n [count])"""
    feat = np.ascontiguousarray(feat, np.float64); world = np.ascontiguousarray(world, np.float64)
    count = feat.shape[0]
    R = np.zeros((count, 4, 3, 3)); t = np.zeros((count, 4, 3)); n = np.zeros(count, np.int32)
    load().oracle_p3p(_vp(feat), _vp(world), count, _vp(R), _vp(t), _vp(n))
    return R, t, n


def triangulate_midpoint_batch(origins, directions, ray_offset):
    """origins, directions [total,3]; ray_offset [num_tracks+1] int64 -> (points [num_tracks,4], ok [num_tracks])"""
    origins = np.ascontiguousarray(origins, np.float64); directions = np.ascontiguousarray(directions, np.float64)
    ray_offset = np.ascontiguousarray(ray_offset, np.int64)
    nt = len(ray_offset) - 1
    out = np.zeros((nt, 4)); ok = np.zeros(nt, np.uint8)
    load().oracle_triangulate_midpoint_batch(_vp(origins), _vp(directions), _vp(ray_offset), nt, _vp(out), _vp(ok))
    return out, ok


def four_point_homography(corr):
    corr = np.ascontiguousarray(corr, np.float64)
    count = corr.shape[0]
    H = np.zeros((count, 3, 3)); ok = np.zeros(count, np.int32)
    load().oracle_four_point_homography(_vp(corr), count, _vp(H), _vp(ok))
    return H, ok


def seven_point_fundamental(corr):
    corr = np.ascontiguousarray(corr, np.float64)
    count = corr.shape[0]
    F = np.zeros((count, 3, 3, 3)); n = np.zeros(count, np.int32)
    load().oracle_seven_point_fundamental(_vp(corr), count, _vp(F), _vp(n))
    return F, n


def poly_roots(poly):
    poly = np.ascontiguousarray(poly, np.float64)
    re = np.zeros(4); im = np.zeros(4)
    n = load().oracle_poly_roots(_vp(poly), len(poly), _vp(re), _vp(im))
    return re[:n] + 1j * im[:n]
