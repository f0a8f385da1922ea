"""Golden vectors under tests/golden/ (restatement-derived, see tests/golden/make_golden.py): the oracle must reproduce them
(CPU, guards the checker against drift) and the CUDA path must match them through the C-ABI (-m gpu) without running the
oracle at test time. Tolerances: BA final cost <= 1e-6 relative and K1 <= 1e-12 relative (north_star); RANSAC / five-point
bit-exact (integer and IEEE-reproducible work)."""
import ctypes as C
import os

import numpy as np
import pytest

from pytheiasfm_b200 import capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def _problem(g):
    return capi.HostBaProblem({k[3:]: v for k, v in g.items() if k.startswith("in_")})


def _batch(g):
    b = capi.HostPairBatch([g["corr"][g["pair_offset"][i]:g["pair_offset"][i + 1]] for i in range(len(g["seed"]))], g["seed"])
    return b


def _results(g):
    # the committed records are the r01 ThbRelPoseResult (200 bytes); later revisions append fields (num_lo_iterations)
    raw = np.ascontiguousarray(g["results"])
    fields = [(n, capi.RELPOSE_DTYPE.fields[n][0]) for n in capi.RELPOSE_DTYPE.names]
    while sum(dt.itemsize for _, dt in fields) > raw.shape[1]:
        fields.pop()
    return np.frombuffer(raw.tobytes(), dtype=np.dtype(fields))


def _c4_params(p):
    from pytheiasfm_b200 import synthetic
    return synthetic.c4_params(p)


# ---------------------------------------------------------------- CPU: the oracle against its own committed outputs
def test_oracle_reproduces_ba_goldens(oracle):
    g = _load("ba_c1_solve.npz")
    prob = _problem(g)
    s = oracle.ba_solve(prob, oracle.default_options())
    assert s["num_iterations"] == int(g["num_iterations"])
    assert abs(s["final_cost"] - float(g["final_cost"])) <= 1e-12 * float(g["final_cost"])
    np.testing.assert_allclose(prob.a["cam_ext"], g["out_cam_ext"], rtol=0, atol=1e-10)
    g = _load("ba_k1_six_models.npz")
    r, jc, ji, jp, ok = oracle.ba_evaluate(_problem(g))
    np.testing.assert_array_equal(ok, g["ok"])
    for a, b in ((r, g["residuals"]), (jc, g["jac_cam"]), (ji, g["jac_intr"]), (jp, g["jac_pt"])):
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)


def test_oracle_reproduces_ransac_goldens(oracle):
    g = _load("ransac_relpose_6pairs.npz")
    rc, res, mask = oracle.ransac_relpose_batch(_batch(g), _c4_params(oracle.ransac_default_params()))
    assert rc == 0
    np.testing.assert_array_equal(mask, g["inlier_mask"])
    gold = _results(g)
    for f in ("success", "num_inliers", "num_iterations"):
        np.testing.assert_array_equal(res[f], gold[f])
    assert np.array_equal(res["essential_matrix"], gold["essential_matrix"])
    g = _load("five_point_32samples.npz")
    E, n = oracle.five_point(g["x1"], g["x2"])
    np.testing.assert_array_equal(n, g["num_solutions"])
    assert np.array_equal(E, g["E"])


# ---------------------------------------------------------------- GPU: the CUDA path against the committed outputs
@pytest.mark.gpu
def test_cuda_ba_matches_goldens(lib):
    g = _load("ba_c1_solve.npz")
    prob = _problem(g)
    s = capi.ThbBaSummary(); o = capi.default_options(lib); p = prob.struct()
    capi.check(lib.thb_ba_solve(C.byref(p), C.byref(o), C.byref(s), None))
    assert s.success == 1 and s.num_iterations == int(g["num_iterations"])
    assert abs(s.final_cost - float(g["final_cost"])) <= 1e-6 * float(g["final_cost"])          # north_star tolerance
    np.testing.assert_allclose(np.array(s.iter_cost[: s.iter_log_count]), g["iter_cost"], rtol=1e-6)
    np.testing.assert_allclose(prob.a["cam_ext"], g["out_cam_ext"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(prob.a["pts"], g["out_pts"], rtol=0, atol=1e-6)

    g = _load("ba_k1_six_models.npz")
    prob = _problem(g)
    n = prob.num_observations
    r = np.zeros((n, 2)); jc = np.zeros((n, 2, 6)); ji = np.zeros((n, 2, capi.THB_INTR_STRIDE)); jp = np.zeros((n, 2, 4))
    ok = np.zeros(n, np.uint8)
    p = prob.struct()
    capi.check(lib.thb_ba_evaluate(C.byref(p), *[a.ctypes.data_as(C.c_void_p) for a in (r, jc, ji, jp, ok)], None))
    np.testing.assert_array_equal(ok, g["ok"])
    good = ok.astype(bool)
    for a, b in ((r, g["residuals"]), (jc, g["jac_cam"]), (ji, g["jac_intr"]), (jp, g["jac_pt"])):
        scale = np.abs(b[good]).max()
        assert np.abs(a[good] - b[good]).max() <= 1e-12 * scale                                  # K1 tolerance


@pytest.mark.gpu
def test_cuda_ransac_matches_goldens_bit_for_bit(lib):
    g = _load("ransac_relpose_6pairs.npz")
    batch = _batch(g)
    res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE); mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
    b = batch.struct(); params = _c4_params(capi.ThbRansacParams())
    capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
    gold = _results(g)
    np.testing.assert_array_equal(mask, g["inlier_mask"])                                        # identical inlier sets
    for f in ("success", "num_inliers", "num_iterations"):
        np.testing.assert_array_equal(res[f], gold[f])
    assert np.array_equal(res["essential_matrix"], gold["essential_matrix"])
    assert np.array_equal(res["rotation"], gold["rotation"]) and np.array_equal(res["position"], gold["position"])

    g = _load("five_point_32samples.npz")
    x1 = np.ascontiguousarray(g["x1"]); x2 = np.ascontiguousarray(g["x2"])
    E = np.zeros((32, 10, 3, 3)); ns = np.zeros(32, np.int32)
    capi.check(lib.thb_five_point_relative_pose(x1.ctypes.data_as(C.c_void_p), x2.ctypes.data_as(C.c_void_p), 32, E.ctypes.data_as(C.c_void_p),
                                                ns.ctypes.data_as(C.c_void_p), None))
    np.testing.assert_array_equal(ns, g["num_solutions"])
    assert np.array_equal(E, g["E"])


# ---------------------------------------------------------------- track stage (triangulation + per-track BA)
def _track_results(g):
    return np.frombuffer(np.ascontiguousarray(g["track_results"]).tobytes(), dtype=capi.TRACK_BA_DTYPE)


def test_oracle_reproduces_track_goldens(oracle):
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import single_track_problem
    g = _load("track_stage_80tracks.npz")
    pts, ok = oracle.triangulate_midpoint_batch(g["tri_origins"], g["tri_directions"], g["tri_ray_offset"])
    np.testing.assert_array_equal(ok, g["tri_ok"])
    np.testing.assert_array_equal(pts, g["tri_points"])
    assert ok[5] == 0 and ok.sum() == 79
    prob = _problem(g)
    o = oracle.default_options()
    o.use_inner_iterations = 0
    want = _track_results(g)
    for t in (0, 17, 63):
        sub = single_track_problem(prob, t)
        s = oracle.ba_solve(sub, o)
        assert s["num_iterations"] == want[t]["num_iterations"]
        assert abs(s["final_cost"] - want[t]["final_cost"]) <= 1e-12 * want[t]["final_cost"]
        np.testing.assert_allclose(sub.a["pts"][t], g["out_pts"][t], rtol=1e-12)


@pytest.mark.gpu
def test_cuda_track_stage_matches_goldens(lib):
    g = _load("track_stage_80tracks.npz")
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    org = np.ascontiguousarray(g["tri_origins"]); d = np.ascontiguousarray(g["tri_directions"]); off = np.ascontiguousarray(g["tri_ray_offset"])
    pts = np.zeros((80, 4)); ok = np.zeros(80, np.uint8)
    capi.check(lib.thb_triangulate_midpoint_batch(vp(org), vp(d), vp(off), 80, capi.THB_MEM_HOST, vp(pts), vp(ok), None))
    np.testing.assert_array_equal(ok, g["tri_ok"])
    np.testing.assert_allclose(pts, g["tri_points"], rtol=1e-12, atol=1e-13)
    prob = _problem(g)
    o = capi.default_options(lib)
    o.use_inner_iterations = 0
    res = np.zeros(80, capi.TRACK_BA_DTYPE)
    p = prob.struct()
    capi.check(lib.thb_ba_tracks_batch(C.byref(p), C.byref(o), vp(res), None))
    want = _track_results(g)
    np.testing.assert_array_equal(res["num_iterations"], want["num_iterations"])
    np.testing.assert_array_equal(res["termination_type"], want["termination_type"])
    np.testing.assert_allclose(res["final_cost"], want["final_cost"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(prob.a["pts"], g["out_pts"], rtol=1e-6, atol=1e-9)
