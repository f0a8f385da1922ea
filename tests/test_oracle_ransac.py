"""CPU tests of the RANSAC oracle: the restated Eigen decompositions against numpy, the reference's own
known-answer tests for the five-point solver (five_point_relative_pose_test.cc:64-200), the relative-pose
RANSAC property tests (estimate_relative_pose_test.cc:65-279 style) and the libstdc++ RNG replay."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi, synthetic


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_jacobi_svd3_matches_numpy(oracle):
    lib = oracle.load()
    rng = np.random.default_rng(0)
    for k in range(200):
        A = rng.normal(size=(3, 3))
        if k % 3 == 0:  # essential-matrix-like: two equal singular values and a zero
            U0, _, V0 = np.linalg.svd(A)
            A = U0 @ np.diag([1.0, 1.0, 0.0]) @ V0
        U = np.zeros((3, 3)); S = np.zeros(3); V = np.zeros((3, 3))
        lib.oracle_jacobi_svd3(_vp(A), _vp(U), _vp(S), _vp(V))
        np.testing.assert_allclose(S, np.linalg.svd(A, compute_uv=False), atol=1e-13)
        np.testing.assert_allclose(U @ np.diag(S) @ V.T, A, atol=1e-13)
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-13)
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-13)
        assert S[0] >= S[1] >= S[2] >= 0


def test_eigensolver10_matches_numpy(oracle):
    lib = oracle.load()
    rng = np.random.default_rng(1)
    for _ in range(100):
        A = rng.normal(size=(10, 10))
        re = np.zeros(10); im = np.zeros(10); vec = np.zeros((10, 10))
        assert lib.oracle_eigen10(_vp(A), _vp(re), _vp(im), _vp(vec)) == 0
        ref = np.linalg.eigvals(A)
        got = re + 1j * im
        assert np.abs(np.sort_complex(got) - np.sort_complex(ref)).max() < 1e-10
        for j in range(10):
            if im[j] == 0.0:  # real eigenvalue: unit eigenvector
                v = vec[:, j]
                assert abs(np.linalg.norm(v) - 1) < 1e-12
                assert np.linalg.norm(A @ v - re[j] * v) < 1e-9 * np.linalg.norm(A)
        # complex eigenvalues come in adjacent conjugate pairs, positive imaginary part first (Eigen's order)
        j = 0
        while j < 10:
            if im[j] != 0.0:
                assert im[j] > 0 and im[j + 1] == -im[j] and re[j] == re[j + 1]
                j += 2
            else:
                j += 1


def test_fullpivlu_kernel_and_solve(oracle):
    lib = oracle.load()
    rng = np.random.default_rng(2)
    for _ in range(100):
        A = rng.normal(size=(5, 9))
        ker = np.zeros((9, 4))
        assert lib.oracle_fullpivlu_kernel_5x9(_vp(A), _vp(ker)) == 4
        assert np.abs(A @ ker).max() < 1e-12 * np.abs(A).max() * np.abs(ker).max() * 10
        assert np.linalg.matrix_rank(ker) == 4
        M = rng.normal(size=(10, 10)); B = rng.normal(size=(10, 10)); X = np.zeros((10, 10))
        lib.oracle_fullpivlu_solve10(_vp(M), _vp(B), _vp(X))
        np.testing.assert_allclose(M @ X, B, atol=1e-9)
    A = rng.normal(size=(5, 9)); A[4] = A[0] + A[1]  # rank deficient: the reference returns false
    assert lib.oracle_fullpivlu_kernel_5x9(_vp(A), _vp(np.zeros((9, 5)))) == 5


FIVE_PT_CASES = [
    # (points, rotation angle about z [deg], translation, noise, tolerance)  five_point_relative_pose_test.cc:116-187
    ([(-1, 3, 3), (1, -1, 2), (3, 1, 2.5), (-1, 1, 2), (2, 1, 3)], 13.0, (1, 1, 1), 0.0, 1e-4),
    ([(-1, 3, 3), (1, -1, 2), (3, 1, 2.5), (-1, 1, 2), (2, 1, 3)], 13.0, (1, 1, 1), 1.0 / 512, 1e-2),
    ([(-1, 3, 3), (1, -1, 2), (3, 1, 2), (-1, 1, 2), (2, 1, 3)], 13.0, (0, 0, 1), 1.0 / 512, 0.15),
    ([(-1, 3, 3), (1, -1, 2), (3, 1, 2), (-1, 1, 2), (2, 1, 3)], 0.0, (1, 1, 1), 1.0 / 512, 0.01),
]


def five_point_case(points, deg, t, noise, seed=52):
    X = np.array(points, float)
    a = np.deg2rad(deg)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    t = np.array(t, float)
    x1 = X[:, :2] / X[:, 2:3]
    P = X @ R.T + t
    x2 = P[:, :2] / P[:, 2:3]
    if noise:
        rng = np.random.default_rng(seed)
        x1 = x1 + rng.normal(0, noise, x1.shape); x2 = x2 + rng.normal(0, noise, x2.shape)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return x1, x2, tx @ R


def equal_up_to_scale(a, b, tol):
    a = a.ravel() / np.linalg.norm(a); b = b.ravel() / np.linalg.norm(b)
    return min(np.abs(a - b).max(), np.abs(a + b).max()) < tol


def sampson(F, x, y):
    xh = np.append(x, 1.0); yh = np.append(y, 1.0)
    ex = F @ xh
    den = (yh @ F[:, 0]) ** 2 + (yh @ F[:, 1]) ** 2 + ex[0] ** 2 + ex[1] ** 2
    return (yh @ ex) ** 2 / den


@pytest.mark.parametrize("case", FIVE_PT_CASES)
def test_five_point_reference_kats(oracle, case):
    pts, deg, t, noise, tol = case
    x1, x2, E_gt = five_point_case(pts, deg, t, noise)
    E, n = oracle.five_point(x1[None], x2[None])
    assert 1 <= n[0] <= 10
    matched = False
    for k in range(n[0]):
        for i in range(5):  # every solution satisfies the epipolar constraints of the minimal sample
            assert sampson(E[0, k], x1[i], x2[i]) < 1e-8
        matched |= equal_up_to_scale(E[0, k], E_gt, tol)
    assert matched


def test_five_point_solutions_are_essential_matrices(oracle):
    rng = np.random.default_rng(3)
    batch, gts = synthetic.make_pair_batch(20, n=5, inlier_ratio=1.0, noise=0.0, seed=4)
    x = batch.corr.reshape(20, 5, 4)
    E, n = oracle.five_point(x[:, :, :2], x[:, :, 2:])
    assert (n >= 1).all()
    for p in range(20):
        R, c, _ = gts[p]
        t = -R @ c
        E_gt = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
        ok = False
        for k in range(n[p]):
            Ek = E[p, k] / np.linalg.norm(E[p, k])
            s = np.linalg.svd(Ek, compute_uv=False)
            assert abs(s[0] - s[1]) < 1e-6 and s[2] < 1e-6  # 2 EE^T E - tr(EE^T) E = 0 and det E = 0
            ok |= equal_up_to_scale(Ek, E_gt, 1e-6)
        assert ok


def test_sampson_and_cheirality_formulas(oracle):
    lib = oracle.load()
    rng = np.random.default_rng(5)
    for _ in range(50):
        F = rng.normal(size=(3, 3)); c = rng.normal(size=4)
        assert abs(lib.oracle_sampson(_vp(F), _vp(c)) - sampson(F, c[:2], c[2:])) < 1e-12 * max(1.0, sampson(F, c[:2], c[2:]))
    batch, gts = synthetic.make_pair_batch(10, n=5, inlier_ratio=1.0, noise=0.0, seed=6)
    for p in range(10):
        R, c, _ = gts[p]
        t = -R @ c
        E = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
        Rb = np.zeros((3, 3)); pb = np.zeros(3)
        corr = np.ascontiguousarray(batch.corr[5 * p: 5 * p + 5])
        cnt = lib.oracle_best_pose(_vp(np.ascontiguousarray(E)), _vp(corr), 5, _vp(Rb), _vp(pb))
        assert cnt == 5
        np.testing.assert_allclose(Rb, R, atol=1e-9)
        np.testing.assert_allclose(pb, c / np.linalg.norm(c), atol=1e-9)


def test_relative_pose_ransac_recovers_pose(oracle):
    """estimate_relative_pose_test.cc style: clean data -> exact pose; noisy 60 %-inlier data -> within tolerance."""
    batch, gts = synthetic.make_pair_batch(6, n=400, inlier_ratio=0.6, noise=1e-3, seed=7)
    params = synthetic.c4_params(oracle.ransac_default_params())
    rc, res, mask = oracle.ransac_relpose_batch(batch, params)
    assert rc == 0
    for p in range(6):
        R, c, flags = gts[p]
        assert res["success"][p] == 1 and 10 <= res["num_iterations"][p] <= 1000
        ang = np.rad2deg(np.arccos(np.clip((np.trace(res["rotation"][p] @ R.T) - 1) / 2, -1, 1)))
        assert ang < 1.0
        assert np.rad2deg(np.arccos(np.clip(res["position"][p] @ c, -1, 1))) < 5.0
        m = mask[batch.pair_offset[p]: batch.pair_offset[p + 1]].astype(bool)
        assert m.sum() == res["num_inliers"][p]
        assert (m & flags).sum() >= 0.85 * flags.sum() and (m & ~flags).sum() <= 0.05 * (~flags).sum()
        assert 0.99 < res["confidence"][p] <= 1.0


def test_ransac_is_deterministic_in_the_seed_and_rejects_bad_params(oracle):
    batch, _ = synthetic.make_pair_batch(3, n=200, seed=8)
    params = synthetic.c4_params(oracle.ransac_default_params())
    a = oracle.ransac_relpose_batch(batch, params, threads=1)
    b = oracle.ransac_relpose_batch(batch, params, threads=4)
    assert a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes()
    batch.seed[:] += 1
    c = oracle.ransac_relpose_batch(batch, params)
    assert c[1].tobytes() != a[1].tobytes()
    params.error_thresh = -1.0
    assert oracle.ransac_relpose_batch(batch, params)[0] == capi.THB_E_INVALID_ARGUMENT
    params = synthetic.c4_params(oracle.ransac_default_params()); params.use_lo = 1
    assert oracle.ransac_batch("abspose", synthetic.make_abspose_batch(2, n=50, seed=1)[0], params)[0] == capi.THB_E_UNSUPPORTED
    lo = oracle.ransac_relpose_batch(batch, params)          # LO-RANSAC for the relative pose: RefineModel runs at least once per pair
    assert lo[0] == 0 and (lo[1]["num_lo_iterations"] >= 1).all()


def test_default_ransac_parameters(oracle):
    p = oracle.ransac_default_params()  # sample_consensus_estimator.h:59-68
    assert p.error_thresh == -1 and p.failure_probability == 0.01 and p.min_inlier_ratio == 0
    assert p.min_iterations == 100 and p.max_iterations == 2**31 - 1 and not p.use_mle and not p.use_lo and p.lo_start_iterations == 50


def test_polynomial_roots_match_numpy(oracle):
    """math/find_polynomial_roots_companion_matrix.cc via the restated balance + EigenSolver; degree 1-4, leading zeros."""
    rng = np.random.default_rng(20)
    for _ in range(200):
        deg = int(rng.integers(1, 5))
        poly = rng.normal(size=deg + 1) * 10 ** rng.uniform(-2, 2, deg + 1)
        if rng.random() < 0.2:
            poly = np.concatenate([[0.0], poly])[: 5]
        got = oracle.poly_roots(poly)
        ref = np.roots(poly)
        assert len(got) == len(ref)
        assert np.abs(np.sort_complex(got) - np.sort_complex(ref)).max() <= 1e-8 * max(1.0, np.abs(ref).max())


def _rot(axis, deg):
    a = np.deg2rad(deg); axis = np.array(axis, float)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


P3P_KAT_POINTS = np.array([[-0.3001, -0.5840, 1.2271], [-1.4487, 0.6965, 0.3889], [-0.7815, 0.7642, 0.1257]])


def p3p_kat(noise, seed=55):
    """perspective_three_point_test.cc:56-112"""
    Rg = _rot([1, 0, 0], 15.0) @ _rot([0, 1, 0], -10.0)
    tg = np.array([0.3, -1.7, 1.15])
    pc = P3P_KAT_POINTS @ Rg.T + tg
    x = pc[:, :2] / pc[:, 2:]
    if noise:
        x = x + np.random.default_rng(seed).normal(0, noise, x.shape)
    return x, Rg, tg


def check_p3p_solutions(R, t, n, x, Rg, tg):
    matched = False
    for k in range(n):
        ang = np.rad2deg(np.arccos(np.clip((np.trace(R[k] @ Rg.T) - 1) / 2, -1, 1)))
        if ang < 1.0 and np.linalg.norm((-Rg @ tg) - (-R[k] @ t[k])) < 0.1:
            matched = True
            pc = P3P_KAT_POINTS @ R[k].T + t[k]
            assert (np.linalg.norm(x - pc[:, :2] / pc[:, 2:], axis=1) * 800.0 < 2.0).all()
    assert matched


@pytest.mark.parametrize("noise", [0.0, 1.0 / 800.0])
def test_p3p_reference_kat(oracle, noise):
    x, Rg, tg = p3p_kat(noise)
    R, t, n = oracle.p3p(x[None], P3P_KAT_POINTS[None])
    assert n[0] == 4                     # the reference keeps the real part of all four roots
    check_p3p_solutions(R[0], t[0], n[0], x, Rg, tg)
    Rc, tc, nc = oracle.p3p(x[None], np.array([[[0, 0, 1.0], [1, 1, 2.0], [2, 2, 3.0]]]))
    assert nc[0] == 0                    # collinear world points: no solution (:196-204)


def test_four_point_homography_properties(oracle):
    """four_point_homography_test.cc:127-205 style: exact data -> transfer error ~ 0, H up to scale."""
    rng = np.random.default_rng(21)
    for _ in range(50):
        H = np.eye(3) + rng.normal(0, 0.3, (3, 3)); H /= H[2, 2]
        x1 = rng.uniform(-1, 1, (4, 2)); y = np.c_[x1, np.ones(4)] @ H.T; x2 = y[:, :2] / y[:, 2:]
        Ho, ok = oracle.four_point_homography(np.c_[x1, x2][None])
        assert ok[0] == 1
        yy = np.c_[x1, np.ones(4)] @ Ho[0].T
        assert np.abs(yy[:, :2] / yy[:, 2:] - x2).max() < 1e-7
        assert equal_up_to_scale(Ho[0], H, 1e-6)


def test_seven_point_reproduces_reference_including_its_quirk(oracle):
    """seven_point_fundamental_matrix.cc:72-152: every returned F satisfies the 7 epipolar constraints (any member of
    the 2-dim null space does); because the cubic's coefficients are stored in reversed order the
    rank-2 constraint det F = 0 is NOT met in general — the oracle reproduces the reference as written."""
    rng = np.random.default_rng(22)
    dets = []
    for _ in range(20):
        c, _, _, _ = synthetic.make_pair(rng, 7, 1.0, 0.0)
        F, n = oracle.seven_point_fundamental(c[None])
        assert 1 <= n[0] <= 3
        for k in range(n[0]):
            for i in range(7):
                assert sampson(F[0, k], c[i, :2], c[i, 2:]) < 1e-12
            dets.append(abs(np.linalg.det(F[0, k] / np.linalg.norm(F[0, k]))))
    assert max(dets) > 1e-6


def test_absolute_pose_and_homography_ransac(oracle):
    batch, gts = synthetic.make_abspose_batch(5, n=300, seed=23)
    p = synthetic.c4_params(oracle.ransac_default_params()); p.error_thresh = (3e-3) ** 2
    rc, res, mask = oracle.ransac_batch("abspose", batch, p)
    assert rc == 0
    for i in range(5):
        R, c, flags = gts[i]
        assert res["success"][i] == 1
        ang = np.rad2deg(np.arccos(np.clip((np.trace(res["rotation"][i] @ R.T) - 1) / 2, -1, 1)))
        assert ang < 1.0 and np.linalg.norm(res["position"][i] - c) < 0.1
        m = mask[batch.pair_offset[i]: batch.pair_offset[i + 1]].astype(bool)
        assert (m & flags).sum() >= 0.85 * flags.sum() and (m & ~flags).sum() <= 0.05 * (~flags).sum()
    batch, gts = synthetic.make_homography_batch(5, n=300, seed=24)
    rc, res, mask = oracle.ransac_batch("homography", batch, p)
    assert rc == 0
    for i in range(5):
        H, flags = gts[i]
        assert res["success"][i] == 1 and equal_up_to_scale(res["essential_matrix"][i], H, 2e-2)
        m = mask[batch.pair_offset[i]: batch.pair_offset[i + 1]].astype(bool)
        assert (m & flags).sum() >= 0.75 * flags.sum() and (m & ~flags).sum() <= 0.05 * (~flags).sum()


def test_prosac_needs_fewer_iterations_on_quality_sorted_data(oracle):
    """RansacType::PROSAC (solvers/prosac_sampler.cc:53-131): with the inliers sorted to the front the progressive sampler finds the
    model in far fewer iterations than the uniform sampler, and the model has the same support."""
    batch, _ = synthetic.make_pair_batch(12, n=500, seed=8, base_seed=55)
    params = synthetic.c4_params(oracle.ransac_default_params())
    rc, res, mask = oracle.ransac_relpose_batch(batch, params)
    lists = []
    for p in range(batch.num_pairs):
        a, b = int(batch.pair_offset[p]), int(batch.pair_offset[p + 1])
        lists.append(batch.corr[a:b][np.argsort(-mask[a:b].astype(np.int32), kind="stable")])
    sorted_batch = capi.HostPairBatch(lists, batch.seed)
    params.ransac_type = 1
    rc2, res2, mask2 = oracle.ransac_relpose_batch(sorted_batch, params)
    assert rc == rc2 == 0 and (res2["success"] == 1).all()
    assert res2["num_iterations"].mean() < 0.8 * res["num_iterations"].mean()   # the confidence bound still asks for ~150 iterations
    assert (res2["num_inliers"] >= 0.9 * res["num_inliers"]).all()
    params.ransac_type = 3   # EXHAUSTIVE
    assert oracle.ransac_relpose_batch(sorted_batch, params)[0] == capi.THB_E_UNSUPPORTED


def lmed_reference(residuals, min_sample_size):
    """numpy restatement of LmedQualityMeasurement::ComputeCost (solvers/lmed_quality_measurement.h:58-118), written from the sorted
    array instead of nth_element: upper median for an even count, mean of the two middle values for an odd one."""
    sq = np.sort(residuals * residuals)
    n = len(sq)
    median = sq[n // 2] if n % 2 == 0 else 0.5 * (sq[n // 2 - 1] + sq[n // 2])
    thr = 2.5 * 1.4826 * (1 + 5.0 / (n - min_sample_size)) * np.sqrt(median)
    return median, (residuals * residuals) < thr * thr


def in_front(c, R, pos):
    """IsTriangulatedPointInFrontOfCameras (triangulation.cc:216-232): RelativePoseEstimator::Error returns the largest double for a
    correspondence that triangulates behind a camera."""
    d1 = np.array([c[0], c[1], 1.0]); d2 = R.T @ np.array([c[2], c[3], 1.0])
    return (d2 @ d2) * (d1 @ pos) - (d1 @ d2) * (d2 @ pos) > 0 and (d1 @ d2) * (d1 @ pos) - (d1 @ d1) * (d2 @ pos) > 0


@pytest.mark.parametrize("n", [400, 401])
def test_lmed_cost_and_inliers(oracle, n):
    """RansacType::LMED (solvers/lmed.h:65-72): the reported cost is the median of the squared residuals of the returned model and
    the inliers are those below the robust threshold derived from it, for even and odd data counts; with 70 % inliers the minimiser
    of the median is a model supported by the true inliers."""
    batch, gts = synthetic.make_pair_batch(6, n=n, inlier_ratio=0.7, seed=21, base_seed=400)
    params = synthetic.c4_params(oracle.ransac_default_params())
    params.ransac_type = 2
    rc, res, mask = oracle.ransac_relpose_batch(batch, params)
    assert rc == 0 and (res["success"] == 1).all()
    for p in range(batch.num_pairs):
        a, b = int(batch.pair_offset[p]), int(batch.pair_offset[p + 1])
        c = batch.corr[a:b]
        E = res["essential_matrix"][p].reshape(3, 3)
        R = res["rotation"][p].reshape(3, 3); pos = res["position"][p]
        r = np.array([sampson(E, c[i, :2], c[i, 2:]) if in_front(c[i], R, pos) else np.finfo(np.float64).max for i in range(b - a)])
        with np.errstate(over="ignore"):
            median, inl = lmed_reference(r, 5)
        assert abs(res["best_cost"][p] - median) <= 1e-9 * median
        flips = np.nonzero(inl != mask[a:b].astype(bool))[0]     # numpy's Sampson error differs from the oracle's by ulps
        assert len(flips) <= 1
        assert res["num_inliers"][p] == mask[a:b].sum()
        truth = gts[p][2]
        assert (mask[a:b].astype(bool) & truth).sum() >= 0.8 * truth.sum()    # the threshold is a heuristic on the median, not the noise level
    params.use_lo = 1
    assert oracle.ransac_relpose_batch(batch, params)[0] == capi.THB_E_UNSUPPORTED
