"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C-ABI vs the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi, synthetic

pytestmark = pytest.mark.gpu

ALL_MODELS = [capi.MODEL_PINHOLE, capi.MODEL_FISHEYE, capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION,
              capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED]


def gpu_evaluate(lib, prob):
    n = prob.num_observations
    r = np.zeros((n, 2)); jc = np.zeros((n, 2, 6)); ji = np.zeros((n, 2, capi.THB_INTR_STRIDE)); jp = np.zeros((n, 2, 4))
    ok = np.zeros(n, np.uint8)
    p = prob.struct()
    capi.check(lib.thb_ba_evaluate(C.byref(p), *[a.ctypes.data_as(C.c_void_p) for a in (r, jc, ji, jp, ok)], None))
    return r, jc, ji, jp, ok


def gpu_solve(lib, prob, opts):
    s = capi.ThbBaSummary()
    p = prob.struct()
    capi.check(lib.thb_ba_solve(C.byref(p), C.byref(opts), C.byref(s), None))
    return s.as_dict()


@pytest.mark.parametrize("model", ALL_MODELS)
def test_k1_residual_jacobian_vs_dual_number_oracle(lib, oracle, model):
    """K1 against the oracle's Jet (ceres-autodiff-equivalent) Jacobian: <= 1e-12 relative."""
    prob, _ = synthetic.make_ba_problem(8, 300, 4, models=(model,), seed=20 + model, w_scale=True)
    if model == capi.MODEL_DIVISION_UNDISTORTION:
        prob.a["intr"][0, 4] = -5e-7
    g = gpu_evaluate(lib, prob)
    o = oracle.ba_evaluate(prob)
    assert g[4].all() and o[4].all()
    for name, a, b in zip(("r", "jc", "ji", "jp"), g[:4], o[:4]):
        scale = np.abs(b).max()
        assert np.abs(a - b).max() <= 1e-12 * scale, (name, np.abs(a - b).max() / scale)


def test_k1_mixed_models_and_failure_flags(lib, oracle):
    prob, _ = synthetic.make_ba_problem(10, 200, 4, models=(capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED, capi.MODEL_FISHEYE), seed=31)
    # put one point behind a double-sphere camera's valid cone and one on top of a camera centre
    c0 = int(prob.a["obs_cam"][0]); p0 = int(prob.a["obs_pt"][0])
    prob.a["pts"][p0, :3] = prob.a["cam_ext"][c0, :3]; prob.a["pts"][p0, 3] = 1.0
    g = gpu_evaluate(lib, prob)
    o = oracle.ba_evaluate(prob)
    np.testing.assert_array_equal(g[4], o[4])
    assert g[4][0] == 0
    good = o[4].astype(bool)
    for a, b in zip(g[:4], o[:4]):
        assert np.abs(a[good] - b[good]).max() <= 1e-12 * np.abs(b[good]).max()


def _compare_solves(lib, oracle, prob, opts, rel=1e-6, check_log=True):
    pg, po = prob.copy(), prob.copy()
    g = gpu_solve(lib, pg, opts)
    o = oracle.ba_solve(po, opts)
    assert o["rc"] == 0
    assert g["success"] == o["success"] == 1
    assert abs(g["initial_cost"] - o["initial_cost"]) <= 1e-10 * o["initial_cost"]
    assert abs(g["final_cost"] - o["final_cost"]) <= rel * max(o["final_cost"], 1e-30), (g["final_cost"], o["final_cost"])
    assert g["termination_type"] == o["termination_type"]
    if check_log:
        assert g["num_iterations"] == o["num_iterations"]
        np.testing.assert_allclose(g["iter_cost"], o["iter_cost"], rtol=1e-6)
        # accept/reject (hence the radius) is rounding-driven once the cost has stopped moving: compare the
        # radius only while an iteration still changes the cost by more than 1e-7 relative
        oc = np.array(o["iter_cost"])
        live = np.concatenate([[True], np.abs(np.diff(oc)) > 1e-7 * oc[1:]])
        n_live = int(np.argmin(live)) if not live.all() else len(live)
        np.testing.assert_allclose(g["iter_radius"][:n_live], o["iter_radius"][:n_live], rtol=1e-4)
    assert g["gpu_launches"] > 0
    return g, o, pg, po


@pytest.mark.parametrize("model", ALL_MODELS)
@pytest.mark.parametrize("homogeneous", [True, False])
def test_k1_shared_camera_kernel_matches_gather_kernel(lib, oracle, monkeypatch, model, homogeneous):
    """The two K1 kernels (camera records gathered through L1 / read from the CTA's shared-memory copy of the table) run
    the same arithmetic: the same iteration logs and refined parameters, and parity with the oracle."""
    prob, _ = synthetic.make_ba_problem(12, 600, 5, models=(model,), seed=70 + model)
    opts = capi.default_options(lib)
    opts.use_homogeneous_point_parametrization = int(homogeneous)
    opts.loss_function_type = capi.LOSS_HUBER if model % 2 else capi.LOSS_TRIVIAL
    opts.robust_loss_width = 2.0
    runs = {}
    for mode in ("gather", "shared"):
        monkeypatch.setenv("THB_K1_MODE", mode)
        pg = prob.copy()
        runs[mode] = (gpu_solve(lib, pg, opts), pg)
    a, b = runs["gather"], runs["shared"]
    # per-observation values are identical; the cost is summed by one atomic per warp, in arrival order
    np.testing.assert_allclose(a[0]["iter_cost"], b[0]["iter_cost"], rtol=1e-11)
    assert a[0]["num_iterations"] == b[0]["num_iterations"]
    # (the two instantiations need not contract the same multiply-adds, and the Euclidean 4-vector points have a gauge
    # direction along which last-bit differences drift)
    for k in ("cam_ext", "pts", "intr"):
        np.testing.assert_allclose(a[1].a[k], b[1].a[k], rtol=1e-6, atol=1e-9)
    _compare_solves(lib, oracle, prob, opts)     # still under THB_K1_MODE=shared


def test_c1_full_ba_matches_oracle(lib, oracle):
    """BASELINE configs[0]: 10 cams / 500 pts / 2k obs, defaults (inner iterations off)."""
    prob, _ = synthetic.config_c1()
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib))
    np.testing.assert_allclose(pg.a["cam_ext"], po.a["cam_ext"], atol=1e-6)
    np.testing.assert_allclose(pg.a["pts"], po.a["pts"], atol=1e-5)


def test_c1_forced_iterations_trajectory(lib, oracle):
    """Tolerance tests disabled, K forced iterations (the benchmark's setting): same cost at every iteration."""
    prob, _ = synthetic.config_c1(seed=4)
    o = capi.default_options(lib)
    o.function_tolerance = -1.0; o.gradient_tolerance = -1.0; o.parameter_tolerance = -1.0; o.max_num_iterations = 12
    _compare_solves(lib, oracle, prob, o)


@pytest.mark.parametrize("model", ALL_MODELS)
def test_every_camera_model_full_ba(lib, oracle, model):
    prob, _ = synthetic.make_ba_problem(12, 400, 4, models=(model,), seed=40 + model)
    if model == capi.MODEL_DIVISION_UNDISTORTION:
        prob.a["intr"][0, 4] = -5e-7
    _compare_solves(lib, oracle, prob, capi.default_options(lib))


@pytest.mark.parametrize("loss", [capi.LOSS_HUBER, capi.LOSS_SOFTLONE, capi.LOSS_CAUCHY, capi.LOSS_ARCTAN, capi.LOSS_TUKEY, capi.LOSS_TRUNCATED])
def test_robust_losses(lib, oracle, loss):
    prob, _ = synthetic.make_ba_problem(10, 300, 4, seed=60 + loss)
    rng = np.random.default_rng(2)
    bad = rng.choice(prob.num_observations, 50, replace=False)
    prob.a["obs_xy"][bad] += rng.normal(0, 25, (50, 2))
    o = capi.default_options(lib); o.loss_function_type = loss; o.robust_loss_width = 2.0
    _compare_solves(lib, oracle, prob, o, check_log=loss not in (capi.LOSS_TUKEY, capi.LOSS_TRUNCATED))


def test_bundle_adjust_view_semantics(lib, oracle):
    """BundleAdjustView (bundle_adjustment.cc:220-236): one camera free, every point constant."""
    prob, gt = synthetic.make_ba_problem(1, 100, 1, seed=52, pixel_sigma=0.0, pt_sigma=0.0, pos_sigma=1e-3, rot_sigma=1e-2)
    prob.a["pt_const"][:] = 1
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib), rel=1e-3, check_log=False)
    assert np.linalg.norm(pg.a["cam_ext"][0, :3] - gt["cam_ext"][0, :3]) < 1e-4


@pytest.mark.parametrize("homogeneous", [1, 0])
def test_bundle_adjust_tracks_semantics(lib, oracle, homogeneous):
    """BundleAdjustTracks (bundle_adjustment.cc:389-418): every camera constant, points free."""
    prob, _ = synthetic.make_ba_problem(4, 100, 3, seed=53, pixel_sigma=0.5, pos_sigma=0.0, rot_sigma=0.0, pt_sigma=0.05)
    prob.a["cam_const"][:] = capi.CAM_CONST_POSITION | capi.CAM_CONST_ORIENTATION
    o = capi.default_options(lib); o.use_homogeneous_point_parametrization = homogeneous
    _compare_solves(lib, oracle, prob, o)


@pytest.mark.parametrize("mask", [capi.CAM_CONST_POSITION, capi.CAM_CONST_ORIENTATION])
def test_constant_position_or_orientation(lib, oracle, mask):
    """BundleAdjustmentOptions::constant_camera_{position,orientation} (bundle_adjuster.cc:357-380)."""
    prob, _ = synthetic.make_ba_problem(8, 200, 4, seed=71)
    before = prob.a["cam_ext"].copy()
    prob.a["cam_const"][:] = mask
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib))
    sl = slice(0, 3) if mask == capi.CAM_CONST_POSITION else slice(3, 6)
    np.testing.assert_array_equal(pg.a["cam_ext"][:, sl], before[:, sl])


def test_partial_reconstruction_semantics(lib, oracle):
    """BundleAdjustPartialReconstruction (bundle_adjustment.cc:111-143): some views constant, some tracks constant."""
    prob, _ = synthetic.make_ba_problem(12, 300, 4, seed=72)
    prob.a["cam_const"][::3] = 3
    prob.a["pt_const"][::5] = 1
    # a constant camera observing a constant point would be a fixed-cost block: drop those observations
    keep = ~((prob.a["cam_const"][prob.a["obs_cam"]] == 3) & (prob.a["pt_const"][prob.a["obs_pt"]] == 1))
    arrays = dict(prob.a)
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        arrays[k] = prob.a[k][keep]
    prob = capi.HostBaProblem(arrays)
    _compare_solves(lib, oracle, prob, capi.default_options(lib))


FREE_F_AND_DISTORTION = {  # GetSubsetFromOptimizeIntrinsicsType(FOCAL_LENGTH | RADIAL_DISTORTION): bit k set => constant
    capi.MODEL_PINHOLE: 0b0011110, capi.MODEL_DOUBLE_SPHERE: 0b0011110, capi.MODEL_EXTENDED_UNIFIED: 0b0011110,
    capi.MODEL_FISHEYE: 0b000011110, capi.MODEL_FOV: 0b01110, capi.MODEL_DIVISION_UNDISTORTION: 0b01110}


def _perturb_intrinsics(prob, rel_f=0.01):
    prob.a["intr"][:, 0] *= 1.0 + rel_f


@pytest.mark.parametrize("model", ALL_MODELS)
def test_intrinsics_refinement_shared_group(lib, oracle, model):
    """One shared intrinsics block (reconstruction.cc:129-140) refined with the cameras: Schur group 1
    (bundle_adjuster.cc:547-577), SubsetManifold on the constant coordinates (:429-441), bounds (:396-427)."""
    prob, gt = synthetic.make_ba_problem(12, 400, 4, models=(model,), seed=80 + model, intr_const_mask=[FREE_F_AND_DISTORTION[model]])
    if model == capi.MODEL_DIVISION_UNDISTORTION:
        prob.a["intr"][0, 4] = -5e-7
    _perturb_intrinsics(prob)
    before = prob.a["intr"].copy()
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib))
    const = [k for k in range(capi.MODEL_NUM_PARAMS[model]) if (FREE_F_AND_DISTORTION[model] >> k) & 1]
    np.testing.assert_array_equal(pg.a["intr"][0, const], before[0, const])
    assert pg.a["intr"][0, 0] != before[0, 0]
    np.testing.assert_allclose(pg.a["intr"], po.a["intr"], rtol=1e-6, atol=1e-9)


def test_c3_scaled_mixed_models_intrinsics_refinement(lib, oracle):
    """BASELINE configs[2] scaled: DoubleSphere + ExtendedUnified groups, FOCAL_LENGTH|RADIAL_DISTORTION free."""
    prob, gt = synthetic.config_c3(scale=0.06)
    _perturb_intrinsics(prob, 0.02)
    o = capi.default_options(lib)
    g, oo, pg, po = _compare_solves(lib, oracle, prob, o)
    np.testing.assert_allclose(pg.a["intr"], po.a["intr"], rtol=1e-6, atol=1e-9)
    # forced iterations (benchmark setting)
    o.function_tolerance = -1.0; o.gradient_tolerance = -1.0; o.parameter_tolerance = -1.0; o.max_num_iterations = 8
    _compare_solves(lib, oracle, prob, o)


def test_c3_full_size_properties(lib):
    """BASELINE configs[2] at full size (500 cams / 50k pts / 400k obs, DoubleSphere + ExtendedUnified groups, focal length
    and distortion refined with bounds): size-independent properties - success, monotone cost, chi^2 noise floor, focal
    lengths back near the generating values, idempotence at the optimum."""
    prob, gt = synthetic.config_c3()
    assert prob.num_cameras == 500 and prob.num_points == 50000 and prob.num_observations == 400000
    truth = prob.a["intr"].copy()
    _perturb_intrinsics(prob, 0.02)
    o = capi.default_options(lib)
    o.max_num_iterations = 50
    g = gpu_solve(lib, prob, o)
    assert g["success"] == 1 and g["num_iterations"] >= 3
    costs = np.array(g["iter_cost"])
    assert np.all(np.diff(costs) <= 1e-9 * costs[:-1])
    dof = 2 * prob.num_observations - (6 * prob.num_cameras + 3 * prob.num_points + 2 * 3 - 7)
    assert abs(g["final_cost"] / (0.5 * 0.25 * dof) - 1.0) < 0.03      # sigma = 0.5 px
    # the pinhole-like EUCM focal length is pulled back; the DoubleSphere one stays within the perturbation (f trades
    # against xi / alpha at the noise floor, and Ceres' function tolerance stops the solve there)
    np.testing.assert_allclose(prob.a["intr"][1, 0], truth[1, 0], rtol=2e-3)
    np.testing.assert_allclose(prob.a["intr"][:, 0], truth[:, 0], rtol=2.5e-2)
    g2 = gpu_solve(lib, prob, capi.default_options(lib))
    assert g2["num_iterations"] <= 2 and abs(g2["final_cost"] - g["final_cost"]) <= 1e-6 * g["final_cost"]


def test_intrinsics_all_free_euclidean_points_and_huber(lib, oracle):
    prob, gt = synthetic.make_ba_problem(10, 300, 5, models=(capi.MODEL_PINHOLE, capi.MODEL_FISHEYE), seed=91, intr_const_mask=[0, 0])
    _perturb_intrinsics(prob, -0.015)
    o = capi.default_options(lib); o.use_homogeneous_point_parametrization = 0
    o.loss_function_type = capi.LOSS_HUBER; o.robust_loss_width = 3.0
    _compare_solves(lib, oracle, prob, o)


def test_intrinsics_bounds_projection(lib, oracle):
    """Infeasible start (EUCM beta < 0.1, DS alpha > 1): IterationZero projects on the box (bundle_adjuster.cc:407-427)."""
    prob, gt = synthetic.config_c3(scale=0.04, seed=5)
    prob.a["intr"][0, 6] = 1.02   # DS alpha  in [0, 1]
    prob.a["intr"][1, 6] = 0.05   # EUCM beta >= 0.1
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib), rel=1e-5)
    assert 0.0 <= pg.a["intr"][0, 6] <= 1.0 and pg.a["intr"][1, 6] >= 0.1


def test_intrinsics_constant_group_next_to_free_group(lib, oracle):
    """Partial BA: the group no optimised view belongs to stays constant (bundle_adjuster.cc:442-459)."""
    prob, gt = synthetic.make_ba_problem(12, 300, 4, models=(capi.MODEL_PINHOLE, capi.MODEL_PINHOLE), seed=93, intr_const_mask=[0b0011110, 0b1111111])
    _perturb_intrinsics(prob)
    prob.a["cam_const"][6:] = 3
    before = prob.a["intr"].copy()
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib))
    np.testing.assert_array_equal(pg.a["intr"][1], before[1])


def test_too_many_free_intrinsics_groups_is_an_error(lib):
    models = (capi.MODEL_PINHOLE,) * 9
    prob, _ = synthetic.make_ba_problem(18, 200, 4, models=models, seed=94, intr_const_mask=[0b0011110] * 9)
    s = capi.ThbBaSummary(); p = prob.struct(); o = capi.default_options(lib)
    assert lib.thb_ba_solve(C.byref(p), C.byref(o), C.byref(s), None) == capi.THB_E_UNSUPPORTED


def test_device_resident_session_matches_one_shot(lib):
    """thb_ba_create/iterate/finish with THB_MEM_DEVICE pointers == thb_ba_solve with host pointers."""
    import torch
    prob, _ = synthetic.config_c1(seed=9)
    opts = capi.default_options(lib)
    opts.function_tolerance = -1.0; opts.gradient_tolerance = -1.0; opts.parameter_tolerance = -1.0; opts.max_num_iterations = 8
    ref = prob.copy()
    g = gpu_solve(lib, ref, opts)
    dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in prob.a.items()}
    p = prob.struct()
    p.memory_space = capi.THB_MEM_DEVICE
    for k, v in dev.items():
        setattr(p, k, None if v is None else v.data_ptr())
    sess = C.c_void_p()
    stream = torch.cuda.current_stream().cuda_stream
    capi.check(lib.thb_ba_create(C.byref(p), C.byref(opts), C.c_void_p(stream), C.byref(sess)))
    ran = C.c_int32(0)
    capi.check(lib.thb_ba_iterate(sess, 3, C.byref(ran)))
    assert ran.value == 3
    capi.check(lib.thb_ba_iterate(sess, 100, C.byref(ran)))
    assert ran.value == 5
    s = capi.ThbBaSummary()
    capi.check(lib.thb_ba_finish(sess, C.byref(s)))
    assert s.num_iterations == 8
    assert abs(s.final_cost - g["final_cost"]) <= 1e-10 * g["final_cost"]
    np.testing.assert_allclose(dev["cam_ext"].cpu().numpy(), ref.a["cam_ext"], rtol=0, atol=1e-8)  # S is summed with atomics: not bit-reproducible


def test_c2_scaled_parity_and_full_size_properties(lib, oracle):
    """configs[1] at 1/20 scale against the oracle; at full size (1k cams / 100k pts / 1M obs) through
    size-independent properties: monotone cost, noise-floor final cost, idempotence at the optimum."""
    prob, _ = synthetic.config_c2(scale=0.05)
    o = capi.default_options(lib)
    o.function_tolerance = -1.0; o.gradient_tolerance = -1.0; o.parameter_tolerance = -1.0; o.max_num_iterations = 8
    _compare_solves(lib, oracle, prob, o)

    prob, _ = synthetic.config_c2()
    assert prob.num_observations == 1000000 and prob.num_cameras == 1000 and prob.num_points == 100000
    o.max_num_iterations = 20
    g = gpu_solve(lib, prob, o)
    assert g["success"] == 1 and g["num_iterations"] == 20
    costs = np.array(g["iter_cost"])
    assert np.all(np.diff(costs) <= 1e-9 * costs[:-1])
    dof = 2 * prob.num_observations - (6 * prob.num_cameras + 3 * prob.num_points - 7)
    assert abs(g["final_cost"] / (0.5 * 0.25 * dof) - 1.0) < 0.02      # chi^2 noise floor, sigma = 0.5 px
    o2 = capi.default_options(lib)
    g2 = gpu_solve(lib, prob, o2)                                       # restart at the optimum: nothing to gain
    assert g2["num_iterations"] <= 2 and abs(g2["final_cost"] - g["final_cost"]) <= 1e-6 * g["final_cost"]


def test_invalid_arguments_are_error_codes(lib):
    prob, _ = synthetic.make_ba_problem(3, 20, 3, seed=1)
    s = capi.ThbBaSummary()
    o = capi.default_options(lib)
    bad = prob.copy(); bad.a["obs_pt"][3] = 10**6
    p = bad.struct()
    assert lib.thb_ba_solve(C.byref(p), C.byref(o), C.byref(s), None) == capi.THB_E_INVALID_ARGUMENT
    o.linear_solver = 7
    p = prob.struct()
    assert lib.thb_ba_solve(C.byref(p), C.byref(o), C.byref(s), None) == capi.THB_E_INVALID_ARGUMENT
    o.linear_solver = capi.SOLVER_SCHUR_PCG
    o.pcg_eta = 0.0
    assert lib.thb_ba_solve(C.byref(p), C.byref(o), C.byref(s), None) == capi.THB_E_INVALID_ARGUMENT
    assert lib.thb_ba_solve(None, None, None, None) == capi.THB_E_INVALID_ARGUMENT


@pytest.mark.parametrize("n", [1, 5, 63, 64, 65, 127, 128, 129, 200, 777, 1500])
def test_k4_dense_cholesky_solve_vs_numpy(lib, n):
    """K4 in isolation: every padding / tile-edge case of the blocked factorisation against numpy."""
    rng = np.random.default_rng(n)
    M = rng.normal(size=(n, n + 3))
    A = M @ M.T + 1e-3 * np.eye(n)
    b = rng.normal(size=n)
    x = np.zeros(n)
    Al = np.tril(A) + np.triu(np.full((n, n), 7.7), 1)      # the strictly upper triangle must never be read
    capi.check(lib.thb_dense_spd_solve(Al.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), n, x.ctypes.data_as(C.c_void_p), None))
    ref = np.linalg.solve(A, b)
    assert np.linalg.norm(A @ x - b) <= 1e-10 * (np.linalg.norm(A) * np.linalg.norm(ref) + np.linalg.norm(b))
    np.testing.assert_allclose(x, ref, rtol=1e-6, atol=1e-9 * np.abs(ref).max())


@pytest.mark.parametrize("n", [300, 2500, 6000])
def test_k4_full_size_residual(lib, n):
    """K4 at the reduced-camera-system size of C2 (6000): device-built SPD system, max-norm residual of the solve."""
    ms, res = C.c_double(0.0), C.c_double(1.0)
    capi.check(lib.thb_dense_spd_time(n, 1, C.byref(ms), C.byref(res), None))
    assert res.value <= 1e-11, res.value


def test_k4_reports_indefinite_matrix(lib):
    n = 150
    A = np.eye(n); A[100, 100] = -1.0
    b = np.ones(n); x = np.zeros(n)
    rc = lib.thb_dense_spd_solve(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), n, x.ctypes.data_as(C.c_void_p), None)
    assert rc == capi.THB_E_NUMERICAL


def test_c2_full_size_matches_oracle(lib, oracle):
    """BASELINE configs[1] at FULL size (1k cams / 100k pts / 1M obs), the benchmark's own solve (reference default
    tolerances): termination, iteration count, per-iteration costs and final cost against the oracle, <= 1e-6 relative
    (north_star's tolerance), refined parameters to 1e-6."""
    prob, _ = synthetic.config_c2()
    assert prob.num_observations == 1000000 and prob.num_cameras == 1000 and prob.num_points == 100000
    oracle.set_num_threads(__import__("os").cpu_count() or 1)
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib))
    assert g["num_iterations"] == o["num_iterations"] >= 3
    np.testing.assert_allclose(pg.a["cam_ext"], po.a["cam_ext"], rtol=0, atol=1e-6)


def test_c3_full_size_matches_oracle(lib, oracle):
    """BASELINE configs[2] at FULL size (500 cams / 50k pts / 400k obs, DoubleSphere + ExtendedUnified groups, focal length
    and distortion refined under bounds) against the oracle: same iteration count, costs <= 1e-6 relative, same intrinsics."""
    prob, _ = synthetic.config_c3()
    assert prob.num_cameras == 500 and prob.num_points == 50000 and prob.num_observations == 400000
    _perturb_intrinsics(prob, 0.02)
    oracle.set_num_threads(__import__("os").cpu_count() or 1)
    g, o, pg, po = _compare_solves(lib, oracle, prob, capi.default_options(lib))
    assert g["num_iterations"] == o["num_iterations"] >= 3
    np.testing.assert_allclose(pg.a["intr"], po.a["intr"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("case", ["c1", "c1_huber_euclidean", "const_position", "c3_intrinsics", "fisheye_partial"])
def test_inner_iterations_match_oracle(lib, oracle, case):
    """use_inner_iterations = true (the reference default): per-camera / per-intrinsics-block / per-point coordinate descent
    after every candidate (k_inner_cam, k_inner_intr + host loop, k_track_ba) against the oracle's CoordinateDescentMinimizer
    restatement: same iteration counts and termination, costs <= 1e-6 relative."""
    o = capi.default_options(lib)
    o.use_inner_iterations = 1
    if case == "c1":
        prob, _ = synthetic.config_c1()
    elif case == "c1_huber_euclidean":
        prob, _ = synthetic.config_c1(seed=4)
        o.loss_function_type = capi.LOSS_HUBER; o.robust_loss_width = 1.5; o.use_homogeneous_point_parametrization = 0
    elif case == "const_position":
        prob, _ = synthetic.make_ba_problem(12, 400, 5, seed=17)
        prob.a["cam_const"][::2] = capi.CAM_CONST_POSITION
        prob.a["cam_const"][1] = capi.CAM_CONST_ORIENTATION
        prob.a["pt_const"][::7] = 1
    elif case == "c3_intrinsics":
        prob, _ = synthetic.config_c3(scale=0.06)
        _perturb_intrinsics(prob, 0.02)
    else:
        prob, _ = synthetic.make_ba_problem(9, 300, 4, models=(capi.MODEL_FISHEYE,), seed=23)
        prob.a["cam_const"][5:] = 3
    g, oo, pg, po = _compare_solves(lib, oracle, prob, o)
    assert g["num_iterations"] == oo["num_iterations"]
    np.testing.assert_allclose(pg.a["cam_ext"], po.a["cam_ext"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(pg.a["intr"], po.a["intr"], rtol=1e-6, atol=1e-9)
    # and it is a different path from the plain trust-region solve: the first step lands lower
    o.use_inner_iterations = 0
    plain = gpu_solve(lib, prob.copy(), o)
    assert g["iter_cost"][1] < plain["iter_cost"][1]


def test_inner_iterations_c2_scaled(lib, oracle):
    prob, _ = synthetic.config_c2(scale=0.1)
    o = capi.default_options(lib)
    o.use_inner_iterations = 1
    _compare_solves(lib, oracle, prob, o)


@pytest.mark.parametrize("case", ["c1", "c1_huber", "c3_intrinsics", "c2_scaled", "const_position", "inner"])
def test_schur_pcg_matches_oracle(lib, oracle, case):
    """THB_SOLVER_SCHUR_PCG = ceres ITERATIVE_SCHUR + SCHUR_JACOBI (bundle_adjustment.h:96-99): the inexact Newton steps of
    the device CG (one persistent kernel over the explicit S) against the oracle's restatement of ceres' CG loop: same number
    of CG iterations, same LM trajectory and final cost (<= 1e-6), and the same minimum as the exact solver."""
    if case == "c1":
        prob, _ = synthetic.config_c1()
    elif case == "c1_huber":
        prob, _ = synthetic.config_c1()
    elif case == "c3_intrinsics":
        prob, _ = synthetic.config_c3(scale=0.04)
        _perturb_intrinsics(prob, 0.02)
    elif case == "c2_scaled":
        prob, _ = synthetic.config_c2(scale=0.05)
    elif case == "const_position":
        prob, _ = synthetic.config_c1()
        prob.a["cam_const"][::2] = capi.CAM_CONST_POSITION
        prob.a["cam_const"][3] = capi.CAM_CONST_POSITION | capi.CAM_CONST_ORIENTATION
    else:
        prob, _ = synthetic.config_c1()
    o = capi.default_options(lib)
    o.linear_solver = capi.SOLVER_SCHUR_PCG
    if case == "c1_huber":
        o.loss_function_type = capi.LOSS_HUBER; o.robust_loss_width = 2.0
    if case == "inner":
        o.use_inner_iterations = 1
    g, orc, pg, po = _compare_solves(lib, oracle, prob, o)
    assert g["num_linear_solver_iterations"] == orc["num_linear_solver_iterations"] > 0
    exact = capi.default_options(lib)
    exact.loss_function_type = o.loss_function_type; exact.robust_loss_width = o.robust_loss_width
    e = gpu_solve(lib, prob.copy(), exact)
    assert abs(g["final_cost"] - e["final_cost"]) <= 1e-3 * e["final_cost"]
    assert g["num_linear_solver_iterations"] < 200 * g["num_linear_solves"]


def test_schur_pcg_c2_full_size(lib, oracle):
    """C2 at full size through the iterative solver: parity with the oracle, and the linear solve is an order of magnitude
    cheaper than the dense factorisation (HBM-bound passes over S instead of n^3 / 3 flops)."""
    prob, _ = synthetic.config_c2()
    oracle.set_num_threads(__import__("os").cpu_count() or 1)
    o = capi.default_options(lib)
    o.linear_solver = capi.SOLVER_SCHUR_PCG
    g, orc, pg, po = _compare_solves(lib, oracle, prob, o)
    assert g["num_linear_solver_iterations"] == orc["num_linear_solver_iterations"]
    e = gpu_solve(lib, prob.copy(), capi.default_options(lib))
    assert abs(g["final_cost"] - e["final_cost"]) <= 1e-4 * e["final_cost"]
    assert g["ms_solve"] / g["num_linear_solves"] < 0.5 * e["ms_solve"] / e["num_linear_solves"]


def _gpu_covariance(lib, prob, opts):
    cc = np.zeros((prob.num_cameras, 6, 6)); co = np.full(prob.num_cameras, 9, np.uint8)
    pc = np.zeros((prob.num_points, 3, 3)); po = np.full(prob.num_points, 9, np.uint8)
    p = prob.struct()
    rc = lib.thb_ba_covariance(C.byref(p), C.byref(opts), cc.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p),
                               pc.ctypes.data_as(C.c_void_p), po.ctypes.data_as(C.c_void_p), None)
    return rc, cc, co, pc, po


@pytest.mark.parametrize("loss", [capi.LOSS_TRIVIAL, capi.LOSS_HUBER])
def test_covariance_of_views_and_tracks_matches_oracle(lib, oracle, loss):
    """ceres::Covariance behind BundleAdjustView(s) / BundleAdjustTrack(s) with covariance (bundle_adjuster.cc:660-773): tangent-space
    (J^T J)^-1 of every free block with the loss applied; against the oracle's Jets and, for the trivial loss, against numpy on the
    ambient Jacobians of thb_ba_evaluate."""
    prob, gt = synthetic.make_ba_problem(9, 400, 5, models=(capi.MODEL_PINHOLE, capi.MODEL_FISHEYE), seed=21)
    o = capi.default_options(lib); o.loss_function_type = loss; o.robust_loss_width = 1.0
    # (A) the AddView problem: every point constant, one camera fully constant, one with a constant position
    a = prob.copy()
    a.a["pt_const"][:] = 1
    a.a["cam_const"][2] = capi.CAM_CONST_POSITION
    keep = a.a["obs_cam"] != 4                                 # camera 4 is unobserved: no covariance
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        a.a[k] = None if a.a[k] is None else a.a[k][keep]
    a = capi.HostBaProblem(a.a)
    rc, cc, co, pc, po = _gpu_covariance(lib, a, o)
    orc, occ, oco, opc, opo = oracle.ba_covariance(a, o)
    assert rc == orc == 0
    np.testing.assert_array_equal(co, oco)
    assert co[4] == 0 and co[2] == 1 and co.sum() == a.num_cameras - 1
    np.testing.assert_allclose(cc, occ, rtol=1e-9, atol=1e-18)
    assert np.all(cc[2, :3, :] == 0) and np.all(cc[2, :, :3] == 0) and np.all(np.diag(cc[2])[3:] > 0)
    if loss == capi.LOSS_TRIVIAL:
        r, jc, ji, jp, ok = gpu_evaluate(lib, a)
        for c in (0, 7):
            J = jc[a.a["obs_cam"] == c].reshape(-1, 6)
            np.testing.assert_allclose(cc[c], np.linalg.inv(J.T @ J), rtol=1e-8)
    # (B) the AddTrack problem: every camera constant
    b = prob.copy()
    b.a["cam_const"][:] = 3
    keep = b.a["obs_pt"] != 0                                  # point 0 is unobserved: no covariance
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        b.a[k] = None if b.a[k] is None else b.a[k][keep]
    b = capi.HostBaProblem(b.a)
    rc, cc, co, pc, po = _gpu_covariance(lib, b, o)
    orc, occ, oco, opc, opo = oracle.ba_covariance(b, o)
    assert rc == orc == 0
    np.testing.assert_array_equal(po, opo)
    assert po[0] == 0 and po[1] == 1
    np.testing.assert_allclose(pc, opc, rtol=1e-9, atol=1e-18)
    assert np.all(np.linalg.eigvalsh(pc[po == 1]) > 0)
    # everything free couples the blocks; Euclidean points have a singular covariance (bundle_adjustment.cc:344-345)
    assert _gpu_covariance(lib, prob, o)[0] == capi.THB_E_UNSUPPORTED
    assert oracle.ba_covariance(prob, o)[0] == capi.THB_E_UNSUPPORTED
    o.use_homogeneous_point_parametrization = 0
    assert _gpu_covariance(lib, b, o)[0] == capi.THB_E_INVALID_ARGUMENT


def _with_position_priors(prob, gt, bias, weight, every=1, seed=0):
    rng = np.random.default_rng(seed)
    a = dict(prob.a)
    nc = prob.num_cameras
    has = np.zeros(nc, np.uint8); has[::every] = 1
    a["cam_has_position_prior"] = has
    a["cam_position_prior"] = gt["cam_ext"][:, :3] + bias + rng.normal(0, 0.01, (nc, 3))
    info = np.zeros((nc, 9))
    for c in range(nc):
        M = rng.normal(size=(3, 3)) * 0.2 + np.eye(3)            # a general (non-symmetric) square-root information matrix
        info[c] = (weight * M).reshape(9)
    a["cam_position_prior_sqrt_info"] = info
    return capi.HostBaProblem(a)


def _with_gravity_priors(prob, gt, weight, tilt=0.0, every=1, seed=0):
    rng = np.random.default_rng(seed)
    a = dict(prob.a)
    nc = prob.num_cameras
    has = np.zeros(nc, np.uint8); has[::every] = 1
    g = np.zeros((nc, 3)); info = np.zeros((nc, 9))
    for c in range(nc):
        R = synthetic.rotmat_from_rotvec(gt["cam_ext"][c, 3:] + tilt * rng.normal(size=3))
        g[c] = R @ np.array([0.0, 0.0, -1.0])
        info[c] = (weight * (np.eye(3) + 0.1 * rng.normal(size=(3, 3)))).reshape(9)
    a["cam_has_gravity_prior"] = has; a["cam_gravity_prior"] = g; a["cam_gravity_prior_sqrt_info"] = info
    return capi.HostBaProblem(a)


@pytest.mark.parametrize("case", ["gravity", "gravity_inner_position", "gravity_const_orientation", "gravity_covariance"])
def test_gravity_priors_match_oracle(lib, oracle, case):
    """BundleAdjustmentOptions::use_gravity_priors (bundle_adjuster.cc:165-168, gravity_error.h:44-86): residual
    sqrt_info * (R(aa) (0,0,-1) - prior), nonlinear in the orientation, differentiated on the device through the SO(3) left
    Jacobian and in the oracle with Jets. Alone, together with position priors and inner iterations, next to constant
    orientations, and in the covariance."""
    prob, gt = synthetic.config_c1()
    o = capi.default_options(lib)
    if case == "gravity_const_orientation":
        prob.a["cam_const"][2::4] = capi.CAM_CONST_ORIENTATION
    pp = _with_gravity_priors(prob, gt, weight=200.0, tilt=0.02, every=1 if case != "gravity_const_orientation" else 2, seed=6)
    if case == "gravity_inner_position":
        o.use_inner_iterations = 1
        pp = _with_position_priors(pp, gt, bias=0.2, weight=20.0, every=3, seed=7)
    if case == "gravity_covariance":
        pp.a["pt_const"][:] = 1
        pp = capi.HostBaProblem(pp.a)
        rc, cc, co, pc, po = _gpu_covariance(lib, pp, o)
        orc, occ, oco, opc, opo = oracle.ba_covariance(pp, o)
        assert rc == orc == 0
        np.testing.assert_array_equal(co, oco)
        np.testing.assert_allclose(cc, occ, rtol=1e-8, atol=1e-18)
        return
    g, orc, pg, po = _compare_solves(lib, oracle, pp, o)
    np.testing.assert_allclose(pg.a["cam_ext"], po.a["cam_ext"], rtol=0, atol=1e-6)
    plain = gpu_solve(lib, capi.HostBaProblem({k: v for k, v in pp.a.items() if "prior" not in k}), o)
    assert g["initial_cost"] > plain["initial_cost"] and g["final_cost"] > plain["final_cost"]


def _with_orientation_priors(prob, gt, weight, tilt=0.0, every=1, seed=0):
    rng = np.random.default_rng(seed)
    a = dict(prob.a)
    nc = prob.num_cameras
    has = np.zeros(nc, np.uint8); has[::every] = 1
    a["cam_has_orientation_prior"] = has
    a["cam_orientation_prior"] = gt["cam_ext"][:, 3:] + tilt * rng.normal(size=(nc, 3))
    a["cam_orientation_prior_sqrt_info"] = np.stack([(weight * (np.eye(3) + 0.1 * rng.normal(size=(3, 3)))).reshape(9) for _ in range(nc)])
    return capi.HostBaProblem(a)


@pytest.mark.parametrize("case", ["orientation", "orientation_inner_all_priors", "orientation_const", "orientation_pcg", "orientation_covariance"])
def test_orientation_priors_match_oracle(lib, oracle, case):
    """BundleAdjustmentOptions::use_orientation_priors (bundle_adjuster.cc:170-172, orientation_error.h:44-80): residual
    sqrt_info * log(exp(w) exp(prior)^-1); on the device through unit quaternions with the analytic Jacobian
    A J_l^-1(e) J_l(w), in the oracle through Jets over the restated Sophus code. Alone, with all three prior kinds and inner
    iterations, next to constant orientations, with the iterative solver, and in the covariance."""
    prob, gt = synthetic.config_c1()
    o = capi.default_options(lib)
    if case == "orientation_const":
        prob.a["cam_const"][1::3] = capi.CAM_CONST_ORIENTATION
    pp = _with_orientation_priors(prob, gt, weight=60.0, tilt=0.03, every=2 if case == "orientation_const" else 1, seed=16)
    if case == "orientation_inner_all_priors":
        o.use_inner_iterations = 1
        pp = _with_position_priors(pp, gt, bias=0.2, weight=20.0, every=3, seed=7)
        pp = _with_gravity_priors(pp, gt, weight=100.0, tilt=0.02, every=2, seed=8)
    if case == "orientation_pcg":
        o.linear_solver = capi.SOLVER_SCHUR_PCG
    if case == "orientation_covariance":
        pp.a["pt_const"][:] = 1
        pp = capi.HostBaProblem(pp.a)
        rc, cc, co, pc, po = _gpu_covariance(lib, pp, o)
        orc, occ, oco, opc, opo = oracle.ba_covariance(pp, o)
        assert rc == orc == 0
        np.testing.assert_array_equal(co, oco)
        np.testing.assert_allclose(cc, occ, rtol=1e-8, atol=1e-18)
        return
    g, orc, pg, po = _compare_solves(lib, oracle, pp, o)
    np.testing.assert_allclose(pg.a["cam_ext"], po.a["cam_ext"], rtol=0, atol=1e-6)
    plain = gpu_solve(lib, capi.HostBaProblem({k: v for k, v in pp.a.items() if "prior" not in k}), o)
    assert g["initial_cost"] > plain["initial_cost"] and g["final_cost"] > plain["final_cost"]


@pytest.mark.parametrize("case", ["c1", "inner", "pcg", "const_position_huber", "views_covariance"])
def test_position_priors_match_oracle(lib, oracle, case):
    """BundleAdjustmentOptions::use_position_priors (bundle_adjuster.cc:160-163, position_error.h:44-80): 3 residuals
    sqrt_info * (prior - position) per camera, no loss. Same LM trajectory and final cost as the oracle with the exact and the
    iterative solver, with inner iterations, next to constant positions (the prior is then a constant of the cost), and the
    cameras are pulled to the (biased) priors: the free gauge of the scene is fixed by them."""
    prob, gt = synthetic.config_c1()
    o = capi.default_options(lib)
    every = 1
    if case == "inner":
        o.use_inner_iterations = 1
    elif case == "pcg":
        o.linear_solver = capi.SOLVER_SCHUR_PCG
    elif case == "const_position_huber":
        prob.a["cam_const"][1::3] = capi.CAM_CONST_POSITION
        o.loss_function_type = capi.LOSS_HUBER; o.robust_loss_width = 2.0
        every = 2
    pp = _with_position_priors(prob, gt, bias=0.5, weight=30.0, every=every, seed=4)
    if case == "views_covariance":                          # the AddView problem: priors enter the covariance as well
        pp.a["pt_const"][:] = 1
        pp = capi.HostBaProblem(pp.a)
        rc, cc, co, pc, po = _gpu_covariance(lib, pp, o)
        orc, occ, oco, opc, opo = oracle.ba_covariance(pp, o)
        assert rc == orc == 0
        np.testing.assert_array_equal(co, oco)
        np.testing.assert_allclose(cc, occ, rtol=1e-9, atol=1e-18)
        no_prior = capi.HostBaProblem({k: v for k, v in pp.a.items() if "prior" not in k})
        rc2, cc2, _, _, _ = _gpu_covariance(lib, no_prior, o)
        assert np.all(np.diagonal(cc, axis1=1, axis2=2)[:, :3] < np.diagonal(cc2, axis1=1, axis2=2)[:, :3])   # the prior adds information
        return
    g, orc, pg, po = _compare_solves(lib, oracle, pp, o)
    free = np.nonzero((pp.a["cam_const"] & capi.CAM_CONST_POSITION) == 0)[0] if pp.a["cam_const"] is not None else np.arange(pp.num_cameras)
    with_prior = [c for c in free if pp.a["cam_has_position_prior"][c]]
    if case != "const_position_huber":
        d = pg.a["cam_ext"][with_prior, :3] - pp.a["cam_position_prior"][with_prior]
        assert np.abs(d).max() < 0.05                        # the cameras sit on their priors (0.5 away from the ground truth)
    np.testing.assert_allclose(pg.a["cam_ext"], po.a["cam_ext"], rtol=0, atol=1e-6)
    plain = gpu_solve(lib, capi.HostBaProblem({k: v for k, v in pp.a.items() if "prior" not in k}), o)
    assert g["initial_cost"] > plain["initial_cost"]
