"""CPU tests of the oracle (the checker): derivative exactness, manifold identities, LM behaviour,
an independent scipy cross-check of the final cost, and the reference's own BA property tests
(bundle_adjustment_test.cc:76-258) restated on the oracle."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi, synthetic

ALL_MODELS = [capi.MODEL_PINHOLE, capi.MODEL_FISHEYE, capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION,
              capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED]


def _fd(oracle, prob, name, idx, eps):
    a = prob.copy(); a.a[name][idx] += eps
    b = prob.copy(); b.a[name][idx] -= eps
    return (oracle.ba_evaluate(a)[0] - oracle.ba_evaluate(b)[0]) / (2 * eps)


@pytest.mark.parametrize("model", ALL_MODELS)
def test_jet_jacobians_match_central_differences(oracle, model):
    prob, _ = synthetic.make_ba_problem(6, 60, 3, models=(model,), seed=10 + model)
    if model == capi.MODEL_DIVISION_UNDISTORTION:
        prob.a["intr"][0, 4] = -5e-7
    r, jc, ji, jp, ok = oracle.ba_evaluate(prob)
    assert ok.all()
    K = capi.MODEL_NUM_PARAMS[model]
    for k in range(6):
        d = _fd(oracle, prob, "cam_ext", (slice(None), k), 1e-6)
        np.testing.assert_allclose(jc[:, :, k], d, rtol=2e-5, atol=2e-5 * np.abs(jc).max())
    for k in range(4):
        d = _fd(oracle, prob, "pts", (slice(None), k), 1e-6)
        np.testing.assert_allclose(jp[:, :, k], d, rtol=2e-5, atol=2e-5 * np.abs(jp).max())
    for k in range(K):
        h = 1e-6 * max(1.0, abs(prob.a["intr"][0, k]))
        if model == capi.MODEL_DIVISION_UNDISTORTION and k == 4:
            h = 1e-10
        d = _fd(oracle, prob, "intr", (0, k), h)
        np.testing.assert_allclose(ji[:, :, k], d, rtol=5e-5, atol=5e-5 * np.abs(ji[:, :, k]).max() + 1e-9)
    assert np.all(ji[:, :, K:] == 0)


def test_sphere_manifold_identities(oracle):
    lib = oracle.load()
    rng = np.random.default_rng(0)
    for _ in range(50):
        x = rng.normal(size=4) * rng.uniform(0.1, 5)
        d = rng.normal(size=3) * 0.3
        out = np.zeros(4); J = np.zeros((4, 3))
        lib.oracle_sphere_plus(x.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        lib.oracle_sphere_plus_jacobian(x.ctypes.data_as(C.c_void_p), J.ctypes.data_as(C.c_void_p))
        assert abs(np.linalg.norm(out) - np.linalg.norm(x)) < 1e-12 * np.linalg.norm(x)   # stays on the sphere
        np.testing.assert_allclose(J.T @ x, 0, atol=1e-12 * np.linalg.norm(x) ** 2)         # tangent to it
        eps = 1e-6                                                                          # J = dPlus/ddelta at 0
        for k in range(3):
            e = np.zeros(3); e[k] = eps
            p = np.zeros(4); m = np.zeros(4)
            lib.oracle_sphere_plus(x.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
            lib.oracle_sphere_plus(x.ctypes.data_as(C.c_void_p), (-e).ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p))
            np.testing.assert_allclose((p - m) / (2 * eps), J[:, k], atol=1e-8 * np.linalg.norm(x))
        z = np.zeros(3)
        lib.oracle_sphere_plus(x.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(out, x)


def test_c1_converges_to_noise_floor(oracle):
    prob, gt = synthetic.config_c1()
    o = oracle.default_options()
    s = oracle.ba_solve(prob, o)
    assert s["rc"] == 0 and s["success"] == 1 and s["termination_type"] == capi.TERM_CONVERGENCE
    dof = 2 * prob.num_observations - (6 * prob.num_cameras + 3 * prob.num_points - 7)
    expected = 0.5 * 0.25 * dof
    assert abs(s["final_cost"] - expected) / expected < 0.15
    costs = s["iter_cost"]
    assert all(costs[i + 1] <= costs[i] + 1e-9 for i in range(len(costs) - 1))


def test_final_cost_matches_scipy_least_squares(oracle):
    """Independent cross-check: a dense trust-region solver on the same residuals (Euclidean xyz,
    w fixed at 1) must reach the same minimum as the oracle's LM + Schur + sphere manifold."""
    from scipy.optimize import least_squares
    prob, _ = synthetic.make_ba_problem(4, 24, 3, seed=5, pixel_sigma=0.5)
    base = prob.copy()
    nc, npnt = base.num_cameras, base.num_points

    def fun(x):
        q = base.copy()
        q.a["cam_ext"][:] = x[: nc * 6].reshape(nc, 6)
        q.a["pts"][:, :3] = x[nc * 6:].reshape(npnt, 3)
        return oracle.ba_evaluate(q)[0].ravel()

    x0 = np.concatenate([base.a["cam_ext"].ravel(), base.a["pts"][:, :3].ravel()])
    sol = least_squares(fun, x0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=400)
    o = oracle.default_options()
    o.function_tolerance = 1e-14; o.parameter_tolerance = 1e-14; o.gradient_tolerance = 1e-12; o.max_num_iterations = 200
    s = oracle.ba_solve(prob, o)
    assert s["success"] == 1
    assert abs(s["final_cost"] - sol.cost) <= 1e-7 * sol.cost


@pytest.mark.parametrize("loss", [capi.LOSS_HUBER, capi.LOSS_SOFTLONE, capi.LOSS_CAUCHY, capi.LOSS_ARCTAN, capi.LOSS_TUKEY, capi.LOSS_TRUNCATED])
def test_robust_losses_reduce_cost(oracle, loss):
    prob, _ = synthetic.make_ba_problem(8, 200, 4, seed=7)
    rng = np.random.default_rng(1)
    bad = rng.choice(prob.num_observations, 40, replace=False)
    prob.a["obs_xy"][bad] += rng.normal(0, 30, (40, 2))  # outliers
    o = oracle.default_options(); o.loss_function_type = loss; o.robust_loss_width = 2.0
    rc, c0 = oracle.ba_cost(prob, o)
    s = oracle.ba_solve(prob, o)
    assert s["success"] == 1 and s["final_cost"] < c0
    assert abs(s["initial_cost"] - c0) <= 1e-9 * c0


def test_bundle_adjust_view_property(oracle):
    """bundle_adjustment_test.cc:76-115,212-216: 1 camera, 100 points, points constant, start at the
    noise-free truth: 2*final_cost/n < 1e-15. With 0.1 px noise (:218-222): < 0.1."""
    prob, gt = synthetic.make_ba_problem(1, 100, 1, seed=52, pixel_sigma=0.0, pt_sigma=0.0, pos_sigma=0.0, rot_sigma=0.0)
    prob.a["pt_const"][:] = 1
    s = oracle.ba_solve(prob, oracle.default_options())
    assert s["success"] == 1
    assert 2 * s["final_cost"] / prob.num_observations < 1e-15
    prob, gt = synthetic.make_ba_problem(1, 100, 1, seed=52, pixel_sigma=0.1, pt_sigma=0.0, pos_sigma=0.0, rot_sigma=0.0)
    prob.a["pt_const"][:] = 1
    s = oracle.ba_solve(prob, oracle.default_options())
    assert s["success"] == 1 and 2 * s["final_cost"] / prob.num_observations < 0.1


def test_bundle_adjust_view_recovers_pose(oracle):
    """pytests/sfm/bundle_adjuster_test.py:6-15: position recovered to 1e-4 after 1e-3 position noise."""
    prob, gt = synthetic.make_ba_problem(1, 100, 1, seed=52, pixel_sigma=0.0, pt_sigma=0.0, pos_sigma=1e-3, rot_sigma=1e-2)
    prob.a["pt_const"][:] = 1
    s = oracle.ba_solve(prob, oracle.default_options())
    assert s["success"] == 1
    assert np.linalg.norm(prob.a["cam_ext"][0, :3] - gt["cam_ext"][0, :3]) < 1e-4


@pytest.mark.parametrize("homogeneous", [1, 0])
def test_bundle_adjust_tracks_property(oracle, homogeneous):
    """bundle_adjustment_test.cc:117-207,224-258: fixed cameras, 100 points observed 3 times, XYZW /
    XYZW_MANIFOLD, start at the truth: 2*final_cost/n < 1e-15 noise-free, < 0.5 with 0.5 px noise.
    Plus a perturbed start (ours): converges back to a ~zero cost."""
    for pix, pt_sigma, bound in [(0.0, 0.0, 1e-15), (0.5, 0.0, 0.5), (0.0, 0.05, 1e-9)]:
        prob, gt = synthetic.make_ba_problem(4, 100, 3, seed=53, pixel_sigma=pix, pos_sigma=0.0, rot_sigma=0.0, pt_sigma=pt_sigma)
        prob.a["cam_const"][:] = capi.CAM_CONST_POSITION | capi.CAM_CONST_ORIENTATION
        o = oracle.default_options(); o.use_homogeneous_point_parametrization = homogeneous
        s = oracle.ba_solve(prob, o)
        assert s["success"] == 1
        assert 2 * s["final_cost"] / prob.num_observations < bound


def test_rejects_unsupported_and_invalid(oracle):
    prob, _ = synthetic.make_ba_problem(3, 20, 3, seed=1)
    bad = prob.copy(); bad.a["obs_cam"][0] = 99
    assert oracle.ba_solve(bad, oracle.default_options())["rc"] == capi.THB_E_INVALID_ARGUMENT


def test_intrinsics_refinement_recovers_focal_and_respects_subset(oracle):
    """bundle_adjuster.cc:382-441: FOCAL_LENGTH|RADIAL_DISTORTION free on a shared block, the rest constant."""
    prob, gt = synthetic.make_ba_problem(10, 300, 4, seed=81, pixel_sigma=0.0, intr_const_mask=[0b0011110])
    prob.a["intr"][0, 0] *= 1.01
    before = prob.a["intr"].copy()
    s = oracle.ba_solve(prob, oracle.default_options())
    assert s["success"] == 1 and s["final_cost"] < 1e-8 * s["initial_cost"]
    np.testing.assert_array_equal(prob.a["intr"][0, 1:5], before[0, 1:5])
    assert abs(prob.a["intr"][0, 0] - gt["intr"][0, 0]) < 1e-3 * gt["intr"][0, 0]


def test_intrinsics_bounds_are_enforced(oracle):
    """bundle_adjuster.cc:407-427: DS alpha in [0,1], EUCM beta >= 0.1; an infeasible start is projected first."""
    prob, gt = synthetic.config_c3(scale=0.04, seed=5)
    prob.a["intr"][0, 6] = 1.02
    prob.a["intr"][1, 6] = 0.05
    s = oracle.ba_solve(prob, oracle.default_options())
    assert s["success"] == 1
    assert 0.0 <= prob.a["intr"][0, 6] <= 1.0 and prob.a["intr"][1, 6] >= 0.1


def test_inner_iterations_coordinate_descent(oracle):
    """use_inner_iterations (reference default, bundle_adjustment.h:144; ordering reversed, bundle_adjuster.cc:329-334): every
    candidate is followed by block coordinate descent (cameras, intrinsics, points), so the first accepted step lands
    lower than the plain trust-region step, the solve needs no more iterations, and both end at the same minimum."""
    for make in (lambda: synthetic.config_c1()[0], lambda: synthetic.config_c3(scale=0.04)[0]):
        runs = {}
        for inner in (0, 1):
            prob = make()
            o = oracle.default_options(); o.use_inner_iterations = inner
            runs[inner] = oracle.ba_solve(prob, o)
            assert runs[inner]["rc"] == 0 and runs[inner]["success"] == 1
        assert runs[1]["iter_cost"][1] < runs[0]["iter_cost"][1]
        assert runs[1]["num_iterations"] <= runs[0]["num_iterations"]
        assert abs(runs[1]["final_cost"] - runs[0]["final_cost"]) <= 1e-4 * runs[0]["final_cost"]
        assert np.all(np.diff(runs[1]["iter_cost"]) <= 0)
    # a single parameter block: the preprocessor switches inner iterations off (same result as without)
    prob, _ = synthetic.make_ba_problem(3, 20, 3, seed=1)
    prob.a["cam_const"][:] = 3
    prob.a["pt_const"][1:] = 1
    a, b = prob.copy(), prob.copy()
    o = oracle.default_options()
    ra = oracle.ba_solve(a, o)
    o.use_inner_iterations = 1
    rb = oracle.ba_solve(b, o)
    # (the oracle's OpenMP cost reduction sums in thread order: equal up to the last ulps, not bit for bit)
    assert len(ra["iter_cost"]) == len(rb["iter_cost"])
    np.testing.assert_allclose(ra["iter_cost"], rb["iter_cost"], rtol=1e-12)


def test_iterative_schur_reaches_the_exact_minimum(oracle):
    """ITERATIVE_SCHUR + SCHUR_JACOBI restated (ceres conjugate_gradients_solver.h): inexact Newton steps, so more LM iterations
    than the exact solver, the same minimum; a tighter forcing term eta follows the exact trajectory more closely."""
    for make in (lambda: synthetic.config_c1()[0], lambda: synthetic.config_c2(scale=0.02)[0]):
        exact = oracle.ba_solve(make(), oracle.default_options())
        o = oracle.default_options(); o.linear_solver = capi.SOLVER_SCHUR_PCG
        it = oracle.ba_solve(make(), o)
        assert it["rc"] == 0 and it["success"] == 1 and it["termination_type"] == capi.TERM_CONVERGENCE
        assert it["num_linear_solver_iterations"] > 0 and exact["num_linear_solver_iterations"] == 0
        assert abs(it["final_cost"] - exact["final_cost"]) <= 1e-4 * exact["final_cost"]
        assert np.all(np.diff(it["iter_cost"]) <= 0)
        o.pcg_eta = 1e-8
        tight = oracle.ba_solve(make(), o)
        assert tight["num_linear_solver_iterations"] > it["num_linear_solver_iterations"]
        n = min(len(tight["iter_cost"]), len(exact["iter_cost"]))
        np.testing.assert_allclose(tight["iter_cost"][:n], exact["iter_cost"][:n], rtol=1e-3)


def test_select_good_tracks_oracle_properties(oracle):
    """select_good_tracks_for_bundle_adjustment.cc:263-325 restated: a subset, every view reaches its quota, a finer grid chooses
    more tracks, a caller-provided starting set is kept, views outside the subset impose no quota."""
    prob, _ = synthetic.make_ba_problem(14, 1500, 5, seed=3)
    n, sel = oracle.select_good_tracks(prob, 10, 100, 50)
    assert 0 < n < prob.num_points and n == int((sel != 0).sum())
    for c in range(prob.num_cameras):
        pts = prob.a["obs_pt"][prob.a["obs_cam"] == c]
        assert (sel[pts] != 0).sum() >= min(50, len(pts))
    n_fine, _ = oracle.select_good_tracks(prob, 10, 25, 1)
    n_coarse, _ = oracle.select_good_tracks(prob, 10, 400, 1)
    assert n_fine > n_coarse >= 1
    start = np.zeros(prob.num_points, np.uint8); start[5] = 1
    cam_sel = np.zeros(prob.num_cameras, np.uint8); cam_sel[0] = 1
    n_sub, sel_sub = oracle.select_good_tracks(prob, 10, 100, 50, cam_selected=cam_sel, selected=start)
    assert sel_sub[5] == 1 and n_sub < n
    pts0 = prob.a["obs_pt"][prob.a["obs_cam"] == 0]
    assert set(np.nonzero(sel_sub)[0]) - {5} <= set(pts0.tolist())


def test_position_priors_fix_the_gauge(oracle):
    """use_position_priors (bundle_adjuster.cc:160-163, position_error.h:44-80): priors shifted by a common offset move the whole
    scene along its free translation; the reprojection part of the final cost stays at the no-prior minimum; a prior on a
    camera with a constant position only adds a constant to the cost."""
    prob, gt = synthetic.config_c1()
    base = oracle.ba_solve(prob.copy(), oracle.default_options())
    a = dict(prob.a)
    nc = prob.num_cameras
    a["cam_has_position_prior"] = np.ones(nc, np.uint8)
    a["cam_position_prior"] = gt["cam_ext"][:, :3] + np.array([0.4, 0.0, -0.3])
    a["cam_position_prior_sqrt_info"] = np.tile((np.eye(3) * 50.0).reshape(1, 9), (nc, 1))
    pp = capi.HostBaProblem(a)
    s = oracle.ba_solve(pp, oracle.default_options())
    assert s["rc"] == 0 and s["success"] == 1
    assert s["initial_cost"] > base["initial_cost"]
    assert abs(s["final_cost"] - base["final_cost"]) < 0.05 * base["final_cost"]
    np.testing.assert_allclose(pp.a["cam_ext"][:, :3], a["cam_position_prior"], atol=0.05)
    q = dict(a); q["cam_const"] = np.full(nc, capi.CAM_CONST_POSITION, np.uint8)
    c0 = oracle.ba_solve(capi.HostBaProblem({k: v for k, v in q.items() if "prior" not in k}), oracle.default_options())
    c1 = oracle.ba_solve(capi.HostBaProblem(q), oracle.default_options())
    prior_cost = 0.5 * np.sum((50.0 * (a["cam_position_prior"] - prob.a["cam_ext"][:, :3])) ** 2)
    np.testing.assert_allclose(c1["final_cost"] - c0["final_cost"], prior_cost, rtol=1e-9)
    assert c1["num_iterations"] == c0["num_iterations"]


def test_orientation_priors_against_scipy(oracle):
    """use_orientation_priors (bundle_adjuster.cc:170-172, orientation_error.h:44-80): residual sqrt_info * log(exp(w) exp(prior)^-1).
    The restated Sophus exp / product / log gives the rotation vector scipy computes for R(w) R(prior)^T (initial cost difference
    to the problem without priors), and strong priors pull the orientations onto them."""
    from scipy.spatial.transform import Rotation
    prob, gt = synthetic.config_c1()
    rng = np.random.default_rng(12)
    nc = prob.num_cameras
    a = dict(prob.a)
    prior = gt["cam_ext"][:, 3:] + 0.03 * rng.normal(size=(nc, 3))
    info = np.stack([(40.0 * (np.eye(3) + 0.1 * rng.normal(size=(3, 3)))).reshape(9) for _ in range(nc)])
    a["cam_has_orientation_prior"] = np.ones(nc, np.uint8)
    a["cam_orientation_prior"] = prior
    a["cam_orientation_prior_sqrt_info"] = info
    o = oracle.default_options()
    base = oracle.ba_solve(prob.copy(), o)
    pp = capi.HostBaProblem(a)
    w0 = pp.a["cam_ext"][:, 3:].copy()
    s = oracle.ba_solve(pp, o)
    assert s["rc"] == 0 and s["success"] == 1
    e = (Rotation.from_rotvec(w0) * Rotation.from_rotvec(prior).inv()).as_rotvec()
    expected = 0.5 * sum(np.sum((info[c].reshape(3, 3) @ e[c]) ** 2) for c in range(nc))
    np.testing.assert_allclose(s["initial_cost"] - base["initial_cost"], expected, rtol=1e-9)
    assert np.all(np.diff(s["iter_cost"]) <= 0)
    a["cam_orientation_prior_sqrt_info"] = info * 1e4          # priors that outweigh the reprojection errors
    strong = capi.HostBaProblem(a)
    s2 = oracle.ba_solve(strong, o)
    assert s2["rc"] == 0 and s2["success"] == 1
    after = (Rotation.from_rotvec(strong.a["cam_ext"][:, 3:]) * Rotation.from_rotvec(prior).inv()).magnitude()
    before = (Rotation.from_rotvec(w0) * Rotation.from_rotvec(prior).inv()).magnitude()
    assert after.max() < 0.05 * before.max()

