"""GPU parity tests of hot path 2 (run with -m gpu on a B200): RANSAC relative pose through the C-ABI vs the oracle."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi, synthetic
from test_oracle_ransac import (FIVE_PT_CASES, five_point_case, equal_up_to_scale, sampson, p3p_kat, check_p3p_solutions,
                                P3P_KAT_POINTS)

pytestmark = pytest.mark.gpu


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def gpu_five_point(lib, x1, x2):
    x1 = np.ascontiguousarray(x1, np.float64); x2 = np.ascontiguousarray(x2, np.float64)
    count = x1.shape[0]
    E = np.zeros((count, 10, 3, 3)); n = np.zeros(count, np.int32)
    capi.check(lib.thb_five_point_relative_pose(_vp(x1), _vp(x2), count, _vp(E), _vp(n), None))
    return E, n


def gpu_ransac(lib, batch, params, want_mask=True, kind="relpose"):
    res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE)
    mask = np.zeros(int(batch.pair_offset[-1]), np.uint8) if want_mask else None
    b = batch.struct()
    fn = getattr(lib, "thb_ransac_%s_batch" % kind)
    capi.check(fn(C.byref(b), C.byref(params), _vp(res), None if mask is None else _vp(mask), None))
    return res, mask


@pytest.mark.parametrize("case", FIVE_PT_CASES)
def test_five_point_reference_kats_on_device(lib, case):
    pts, deg, t, noise, tol = case
    x1, x2, E_gt = five_point_case(pts, deg, t, noise)
    E, n = gpu_five_point(lib, x1[None], x2[None])
    assert 1 <= n[0] <= 10
    matched = False
    for k in range(n[0]):
        for i in range(5):
            assert sampson(E[0, k], x1[i], x2[i]) < 1e-8
        matched |= equal_up_to_scale(E[0, k], E_gt, tol)
    assert matched


def test_five_point_matches_oracle_solution_by_solution(lib, oracle):
    """Same number of solutions, same ORDER (it decides RANSAC tie-breaks), same values: the device solver and the
    oracle run the same IEEE operations (-fmad=false / -ffp-contract=off), so the match is bit-exact."""
    rng = np.random.default_rng(11)
    batch, _ = synthetic.make_pair_batch(400, n=5, inlier_ratio=1.0, noise=1e-3, seed=12)
    x = batch.corr.reshape(400, 5, 4).copy()
    x[::7] += rng.normal(0, 0.2, x[::7].shape)          # some samples contaminated by outliers
    Eg, ng = gpu_five_point(lib, x[:, :, :2], x[:, :, 2:])
    Eo, no = oracle.five_point(x[:, :, :2], x[:, :, 2:])
    np.testing.assert_array_equal(ng, no)
    assert np.array_equal(Eg, Eo), np.abs(Eg - Eo).max()


def test_ransac_batch_identical_inlier_sets(lib, oracle):
    """north_star: identical inlier sets given a fixed RANSAC seed. Also the iteration count (i.e. the RNG replay and
    the adaptive bound), the winning model and the confidence."""
    batch, gts = synthetic.make_pair_batch(48, n=500, inlier_ratio=0.6, noise=1e-3, seed=13)
    params = synthetic.c4_params(capi.ThbRansacParams())
    res, mask = gpu_ransac(lib, batch, params)
    rc, ores, omask = oracle.ransac_relpose_batch(batch, synthetic.c4_params(oracle.ransac_default_params()))
    assert rc == 0
    np.testing.assert_array_equal(res["success"], ores["success"])
    np.testing.assert_array_equal(res["num_iterations"], ores["num_iterations"])
    np.testing.assert_array_equal(res["num_inliers"], ores["num_inliers"])
    np.testing.assert_array_equal(mask, omask)
    np.testing.assert_array_equal(res["essential_matrix"], ores["essential_matrix"])
    np.testing.assert_array_equal(res["rotation"], ores["rotation"])
    np.testing.assert_array_equal(res["position"], ores["position"])
    np.testing.assert_allclose(res["best_cost"], ores["best_cost"], rtol=1e-12)
    np.testing.assert_allclose(res["confidence"], ores["confidence"], rtol=1e-9)
    assert (res["num_iterations"] >= 10).all() and (res["num_iterations"] < 1000).any()


@pytest.mark.parametrize("use_mle,min_ratio", [(0, 0.0), (1, 0.3), (0, 0.5)])
def test_ransac_variants_and_ragged_batch(lib, oracle, use_mle, min_ratio):
    rng = np.random.default_rng(14)
    corrs, seeds = [], []
    for i, n in enumerate([5, 6, 37, 200, 1000, 2000, 3, 0, 64, 7000]):  # ragged, incl. too-small and > shared-memory pairs
        if n >= 5:
            c, _, _, _ = synthetic.make_pair(rng, n, 0.7 if n > 10 else 1.0, 1e-3)
        else:
            c = rng.uniform(-1, 1, (n, 4))
        corrs.append(c); seeds.append(77 + i)
    batch = capi.HostPairBatch(corrs, seeds)
    def mk(p):
        p = synthetic.c4_params(p); p.use_mle = use_mle; p.min_inlier_ratio = min_ratio; p.max_iterations = 300
        return p
    res, mask = gpu_ransac(lib, batch, mk(capi.ThbRansacParams()))
    rc, ores, omask = oracle.ransac_relpose_batch(batch, mk(oracle.ransac_default_params()))
    assert rc == 0
    for f in ("success", "num_iterations", "num_inliers", "num_input_data_points", "essential_matrix", "rotation", "position"):
        np.testing.assert_array_equal(res[f], ores[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    assert list(res["success"]) == [1, 1, 1, 1, 1, 1, 0, 0, 1, 1]


def test_ransac_device_resident_batch_and_error_codes(lib):
    import torch
    batch, _ = synthetic.make_pair_batch(8, n=300, seed=15)
    params = synthetic.c4_params(capi.ThbRansacParams())
    res, mask = gpu_ransac(lib, batch, params)
    d_off = torch.from_numpy(batch.pair_offset).cuda(); d_corr = torch.from_numpy(batch.corr).cuda()
    d_seed = torch.from_numpy(batch.seed.astype(np.int64)).to(torch.int64).cuda().to(torch.int32)  # same bits as uint32 for small seeds
    d_res = torch.zeros(8 * capi.RELPOSE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    d_mask = torch.zeros(int(batch.pair_offset[-1]), dtype=torch.uint8, device="cuda")
    b = capi.ThbPairBatch(); b.num_pairs = 8; b.memory_space = capi.THB_MEM_DEVICE
    b.pair_offset = d_off.data_ptr(); b.corr = d_corr.data_ptr(); b.seed = d_seed.data_ptr()
    capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), C.c_void_p(d_res.data_ptr()), C.c_void_p(d_mask.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got = np.frombuffer(d_res.cpu().numpy().tobytes(), capi.RELPOSE_DTYPE)
    assert got.tobytes() == res.tobytes()
    np.testing.assert_array_equal(d_mask.cpu().numpy(), mask)
    bad = synthetic.c4_params(capi.ThbRansacParams()); bad.error_thresh = 0.0
    bb = batch.struct()
    assert lib.thb_ransac_relpose_batch(C.byref(bb), C.byref(bad), _vp(res), None, None) == capi.THB_E_INVALID_ARGUMENT
    bad = synthetic.c4_params(capi.ThbRansacParams()); bad.use_lo = 1
    assert lib.thb_ransac_abspose_batch(C.byref(bb), C.byref(bad), _vp(res), None, None) == capi.THB_E_UNSUPPORTED


def test_c4_full_size_properties(lib):
    """configs[3] shape (2000 correspondences, 60 % inliers) at 512 pairs: every pair recovers its pose, the
    inlier mask is consistent with the reported model (recomputed in numpy), and reruns are bit-identical."""
    batch, gts = synthetic.make_pair_batch(512, n=2000, inlier_ratio=0.6, noise=1e-3, seed=16)
    params = synthetic.c4_params(capi.ThbRansacParams())
    res, mask = gpu_ransac(lib, batch, params)
    res2, mask2 = gpu_ransac(lib, batch, params)
    assert res.tobytes() == res2.tobytes() and mask.tobytes() == mask2.tobytes()
    assert (res["success"] == 1).all()
    bad = 0
    for p in range(512):
        R, c, flags = gts[p]
        ang = np.rad2deg(np.arccos(np.clip((np.trace(res["rotation"][p] @ R.T) - 1) / 2, -1, 1)))
        bad += ang > 2.0
        m = mask[batch.pair_offset[p]: batch.pair_offset[p + 1]].astype(bool)
        assert m.sum() == res["num_inliers"][p]
    assert bad <= 25   # plain RANSAC (no LO) on sigma = 1e-3 noise: a few per cent of minimal-sample models are > 2 degrees off
    # recompute the mask of a few pairs from the returned model in numpy (tolerant at the threshold)
    for p in range(0, 512, 64):
        E = res["essential_matrix"][p]; Rm = res["rotation"][p]; pos = res["position"][p]
        c = batch.corr[batch.pair_offset[p]: batch.pair_offset[p + 1]]
        x = np.concatenate([c[:, :2], np.ones((len(c), 1))], 1); y = np.concatenate([c[:, 2:], np.ones((len(c), 1))], 1)
        ex = x @ E.T
        num = (y * ex).sum(1); dy = y @ E
        r = num ** 2 / (dy[:, 0] ** 2 + dy[:, 1] ** 2 + ex[:, 0] ** 2 + ex[:, 1] ** 2)
        d2 = y @ Rm
        front = ((d2 * d2).sum(1) * (x @ pos) - (x * d2).sum(1) * (d2 @ pos) > 0) & ((x * d2).sum(1) * (x @ pos) - (x * x).sum(1) * (d2 @ pos) > 0)
        ref = front & (r < params.error_thresh)
        m = mask[batch.pair_offset[p]: batch.pair_offset[p + 1]].astype(bool)
        assert (ref != m).sum() <= 2


@pytest.mark.parametrize("noise", [0.0, 1.0 / 800.0])
def test_p3p_reference_kat_on_device(lib, oracle, noise):
    x, Rg, tg = p3p_kat(noise)
    R = np.zeros((1, 4, 3, 3)); t = np.zeros((1, 4, 3)); n = np.zeros(1, np.int32)
    capi.check(lib.thb_p3p(_vp(np.ascontiguousarray(x)), _vp(np.ascontiguousarray(P3P_KAT_POINTS)), 1, _vp(R), _vp(t), _vp(n), None))
    assert n[0] == 4
    check_p3p_solutions(R[0], t[0], n[0], x, Rg, tg)


def test_minimal_solvers_match_oracle_bit_for_bit(lib, oracle):
    """P3P, four-point homography and seven-point F: same solution count, order and values as the oracle."""
    batch, _ = synthetic.make_abspose_batch(64, n=3, inlier_ratio=1.0, noise=1e-3, seed=30)
    d = batch.corr.reshape(64, 3, 5)
    feat = np.ascontiguousarray(d[:, :, :2]); world = np.ascontiguousarray(d[:, :, 2:])
    R = np.zeros((64, 4, 3, 3)); t = np.zeros((64, 4, 3)); n = np.zeros(64, np.int32)
    capi.check(lib.thb_p3p(_vp(feat), _vp(world), 64, _vp(R), _vp(t), _vp(n), None))
    Ro, to, no = oracle.p3p(feat, world)
    np.testing.assert_array_equal(n, no)
    assert np.array_equal(R, Ro, equal_nan=True) and np.array_equal(t, to, equal_nan=True)

    hb, _ = synthetic.make_homography_batch(64, n=4, inlier_ratio=1.0, noise=1e-3, seed=31)
    corr = np.ascontiguousarray(hb.corr.reshape(64, 4, 4))
    H = np.zeros((64, 3, 3)); ok = np.zeros(64, np.int32)
    capi.check(lib.thb_four_point_homography(_vp(corr), 64, _vp(H), _vp(ok), None))
    Ho, oko = oracle.four_point_homography(corr)
    np.testing.assert_array_equal(ok, oko)
    assert np.array_equal(H, Ho)

    rng = np.random.default_rng(32)
    c7 = np.ascontiguousarray(np.stack([synthetic.make_pair(rng, 7, 1.0, 1e-3)[0] for _ in range(64)]))
    F = np.zeros((64, 3, 3, 3)); n7 = np.zeros(64, np.int32)
    capi.check(lib.thb_seven_point_fundamental_matrix(_vp(c7), 64, _vp(F), _vp(n7), None))
    Fo, n7o = oracle.seven_point_fundamental(c7)
    np.testing.assert_array_equal(n7, n7o)
    assert np.array_equal(F, Fo)


def test_round_schedule_over_several_chunks(lib, monkeypatch):
    """More pairs than one chunk of the round-synchronous schedule holds (4 096): the batch is cut into equal chunks; records and
    masks still equal the fused kernel's bit for bit, ragged sizes included."""
    rng = np.random.default_rng(77)
    lists = []
    for i in range(4300):
        c, _, _, _ = synthetic.make_pair(rng, n=int(rng.integers(6, 40)), inlier_ratio=0.8)
        lists.append(c)
    batch = capi.HostPairBatch(lists, 9000 + np.arange(len(lists)))
    params = synthetic.c4_params(capi.ThbRansacParams()); params.max_iterations = 60
    out = {}
    for mode in ("fused", "rounds"):
        monkeypatch.setenv("THB_RANSAC_MODE", mode)
        out[mode] = gpu_ransac(lib, batch, params)
    assert out["fused"][0].tobytes() == out["rounds"][0].tobytes()
    np.testing.assert_array_equal(out["fused"][1], out["rounds"][1])
    assert (out["rounds"][0]["success"] == 1).all()


@pytest.mark.parametrize("kind", ["abspose", "homography"])
def test_abspose_and_homography_ransac_identical_inlier_sets(lib, oracle, kind):
    if kind == "abspose":
        batch, _ = synthetic.make_abspose_batch(24, n=400, seed=33)
    else:
        batch, _ = synthetic.make_homography_batch(24, n=400, seed=34)
    def mk(p):
        p = synthetic.c4_params(p); p.error_thresh = (3e-3) ** 2
        return p
    res, mask = gpu_ransac(lib, batch, mk(capi.ThbRansacParams()), kind=kind)
    rc, ores, omask = oracle.ransac_batch(kind, batch, mk(oracle.ransac_default_params()))
    assert rc == 0
    for f in ("success", "num_iterations", "num_inliers", "num_input_data_points"):
        np.testing.assert_array_equal(res[f], ores[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    for f in ("essential_matrix", "rotation", "position"):
        assert np.array_equal(res[f], ores[f], equal_nan=True), f
    assert (res["num_inliers"] > 100).all()


def test_c4_shaped_pairs_bit_equal_to_oracle(lib, oracle):
    """320 pairs of the benchmark's own C4 table (2000 correspondences, 60 % inliers, C4 parameters, the table's per-pair
    seeds): inlier masks, iteration counts and models identical to the oracle bit for bit; the device work counters agree
    with the iteration counts."""
    batch, _ = synthetic.make_pair_batch_indexed(range(320), n=2000, seed=21)
    params = synthetic.c4_params(capi.ThbRansacParams())
    res, mask = gpu_ransac(lib, batch, params)
    st = capi.ThbRansacStats()
    capi.check(lib.thb_ransac_last_stats(C.byref(st)))
    rc, ores, omask = oracle.ransac_relpose_batch(batch, synthetic.c4_params(oracle.ransac_default_params()))
    assert rc == 0
    for f in ("success", "num_iterations", "num_inliers", "num_input_data_points"):
        np.testing.assert_array_equal(res[f], ores[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    for f in ("essential_matrix", "rotation", "position", "confidence"):
        assert np.array_equal(res[f], ores[f]), f
    # the MLE cost is reduced in warp-tree order on the device, in index order on the CPU: equal to rounding
    np.testing.assert_allclose(res["best_cost"], ores["best_cost"], rtol=1e-13)
    assert st.pairs == 320 and st.iterations == int(res["num_iterations"].sum()) and st.samples_solved >= st.iterations
    assert st.data_scored >= 2000 * 320 and st.models_scored >= 320


def test_use_tdd_test_changes_the_iteration_bound_like_the_oracle(lib, oracle):
    """RansacParameters::use_Tdd_test: ComputeMaxIterations counts SampleSize + 1 draws (sample_consensus_estimator.h:272-279)."""
    batch, _ = synthetic.make_pair_batch_indexed(range(12), n=500, seed=3)
    def mk(p, tdd):
        p = synthetic.c4_params(p); p.use_tdd_test = tdd
        return p
    res0, _ = gpu_ransac(lib, batch, mk(capi.ThbRansacParams(), 0))
    res1, mask1 = gpu_ransac(lib, batch, mk(capi.ThbRansacParams(), 1))
    rc, ores, omask = oracle.ransac_relpose_batch(batch, mk(oracle.ransac_default_params(), 1))
    np.testing.assert_array_equal(res1["num_iterations"], ores["num_iterations"])
    np.testing.assert_array_equal(mask1, omask)
    assert (res1["num_iterations"] >= res0["num_iterations"]).all() and (res1["num_iterations"] > res0["num_iterations"]).any()


def test_pack_inlier_masks_round_trip(lib):
    import torch
    from pytheiasfm_b200 import sharding
    rng = np.random.default_rng(5)
    sizes = [0, 1, 31, 32, 33, 2000, 777, 64]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    mask = (rng.random(int(off[-1])) < 0.4).astype(np.uint8)
    woff = sharding.mask_word_offsets(off)
    words = np.zeros(int(woff[-1]), np.uint32)
    capi.check(lib.thb_pack_inlier_masks(_vp(mask), _vp(off), _vp(woff), len(sizes), capi.THB_MEM_HOST, _vp(words), None))
    for p, n in enumerate(sizes):
        np.testing.assert_array_equal(sharding.unpack_mask_words(words[woff[p]: woff[p + 1]], n), mask[off[p]: off[p + 1]])
    d = [torch.from_numpy(a).cuda() for a in (mask, off, woff)]
    dw = torch.zeros(len(words), dtype=torch.int32, device="cuda")
    capi.check(lib.thb_pack_inlier_masks(C.c_void_p(d[0].data_ptr()), C.c_void_p(d[1].data_ptr()), C.c_void_p(d[2].data_ptr()), len(sizes),
                                         capi.THB_MEM_DEVICE, C.c_void_p(dw.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dw.cpu().numpy().view(np.uint32), words)


def test_fp64_peak_probe(lib):
    t = C.c_double(0.0)
    capi.check(lib.thb_fp64_peak_tflops(3, C.byref(t), None))
    assert 20.0 < t.value < 80.0, t.value


@pytest.mark.parametrize("lo_start", [0, 50])
def test_lo_ransac_relative_pose_matches_oracle(lib, oracle, lo_start):
    """use_lo (the pipelines' default, reconstruction_estimator_options.h:133): RelativePoseEstimator::RefineModel =
    BundleAdjustTwoViewsAngular on every improved model from lo_start_iterations on and once on the final inliers. Same
    iteration counts, LO counts and inlier sets as the oracle; the refined pose agrees to rounding (the device reduces
    J^T J in block-tree order) and is closer to the truth than the plain RANSAC pose."""
    batch, gts = synthetic.make_pair_batch_indexed(range(48), n=1200, seed=9)
    def mk(p, lo):
        p = synthetic.c4_params(p); p.use_lo = lo; p.lo_start_iterations = lo_start
        return p
    res, mask = gpu_ransac(lib, batch, mk(capi.ThbRansacParams(), 1))
    plain, _ = gpu_ransac(lib, batch, mk(capi.ThbRansacParams(), 0))
    rc, ores, omask = oracle.ransac_relpose_batch(batch, mk(oracle.ransac_default_params(), 1))
    assert rc == 0
    for f in ("success", "num_iterations", "num_lo_iterations", "num_inliers"):
        np.testing.assert_array_equal(res[f], ores[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    assert np.array_equal(res["essential_matrix"], ores["essential_matrix"])      # E is not touched by the refinement
    np.testing.assert_allclose(res["rotation"], ores["rotation"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(res["position"], ores["position"], rtol=0, atol=1e-9)
    assert (res["num_lo_iterations"] >= 1).all()
    def rot_err(r):
        return np.array([np.rad2deg(np.arccos(np.clip((np.trace(r["rotation"][i] @ gts[i][0].T) - 1) / 2, -1, 1))) for i in range(batch.num_pairs)])
    assert rot_err(res).mean() < 0.7 * rot_err(plain).mean()


def test_fused_and_round_synchronous_paths_are_bit_identical(lib, oracle, monkeypatch):
    """The two schedules of the same RANSAC loop (one CTA per pair vs one kernel per phase over all pairs, chosen by batch size)
    must give the same records and masks bit for bit, for all three estimators, ragged pairs and per-pair early termination."""
    batch, _ = synthetic.make_pair_batch(40, n=700, seed=21, base_seed=1234)
    params = synthetic.c4_params(oracle.ransac_default_params())
    out = {}
    for mode in ("fused", "rounds"):
        monkeypatch.setenv("THB_RANSAC_MODE", mode)
        res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE); mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
        b = batch.struct()
        capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
        out[mode] = (res.copy(), mask.copy())
    assert out["fused"][0].tobytes() == out["rounds"][0].tobytes()
    np.testing.assert_array_equal(out["fused"][1], out["rounds"][1])
    rc, ores, omask = oracle.ransac_relpose_batch(batch, params)
    np.testing.assert_array_equal(out["rounds"][1], omask)
    np.testing.assert_array_equal(out["rounds"][0]["num_iterations"], ores["num_iterations"])
    for make, entry in ((synthetic.make_abspose_batch, "thb_ransac_abspose_batch"), (synthetic.make_homography_batch, "thb_ransac_homography_batch")):
        hb, _ = make(12, n=300, seed=5)
        params.error_thresh = (3e-3) ** 2
        got = {}
        for mode in ("fused", "rounds"):
            monkeypatch.setenv("THB_RANSAC_MODE", mode)
            res = np.zeros(hb.num_pairs, capi.RELPOSE_DTYPE); mask = np.zeros(int(hb.pair_offset[-1]), np.uint8)
            bb = hb.struct()
            capi.check(getattr(lib, entry)(C.byref(bb), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
            got[mode] = (res.tobytes(), mask.copy())
        assert got["fused"][0] == got["rounds"][0]
        np.testing.assert_array_equal(got["fused"][1], got["rounds"][1])


def test_prosac_sampler_matches_oracle(lib, oracle, monkeypatch):
    """RansacType::PROSAC (solvers/prosac_sampler.cc:53-131): quality-sorted correspondences (inliers first), the progressive
    sampling schedule and its rejection-sampled unique indices on the same mt19937 stream: iteration counts, inlier masks and
    models bit-equal to the oracle in both schedules; fewer iterations than RANSAC on sorted data."""
    batch, _ = synthetic.make_pair_batch(24, n=600, seed=31, base_seed=777)
    # sort every pair's correspondences by "quality": the generator puts outliers at random positions, so order by true residual
    params = synthetic.c4_params(oracle.ransac_default_params())
    rc, ref_res, ref_mask = oracle.ransac_relpose_batch(batch, params)
    lists = []
    for p in range(batch.num_pairs):
        a, b = int(batch.pair_offset[p]), int(batch.pair_offset[p + 1])
        order = np.argsort(-ref_mask[a:b].astype(np.int32), kind="stable")   # RANSAC inliers first
        lists.append(batch.corr[a:b][order])
    batch = capi.HostPairBatch(lists, batch.seed)
    params.ransac_type = 1
    rc, ores, omask = oracle.ransac_relpose_batch(batch, params)
    assert rc == 0
    for mode in ("fused", "rounds"):
        monkeypatch.setenv("THB_RANSAC_MODE", mode)
        res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE); mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
        b = batch.struct()
        capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
        np.testing.assert_array_equal(res["num_iterations"], ores["num_iterations"])
        np.testing.assert_array_equal(mask, omask)
        for f in ("essential_matrix", "rotation", "position"):
            assert np.array_equal(res[f], ores[f], equal_nan=True), f
    assert ores["num_iterations"].mean() < ref_res["num_iterations"].mean()
    params.ransac_type = 3   # EXHAUSTIVE
    b = batch.struct()
    assert lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None) == capi.THB_E_UNSUPPORTED


@pytest.mark.parametrize("n", [600, 601, 37, 5])
def test_lmed_matches_oracle(lib, oracle, n):
    """RansacType::LMED (solvers/lmed.h:65-72, lmed_quality_measurement.h:58-118): the median of the squared residuals found by the
    warp radix select equals the oracle's nth_element median bit for bit (even and odd counts, a count below one warp step), hence
    identical best models, iteration counts and inlier masks; n = 5 is the minimal data set (n - m = 0: the reference's threshold divides
    by zero and becomes infinite - restated as such)."""
    batch, _ = synthetic.make_pair_batch(20, n=n, inlier_ratio=0.7, seed=5, base_seed=4100)
    params = synthetic.c4_params(oracle.ransac_default_params())
    params.ransac_type = 2
    rc, ores, omask = oracle.ransac_relpose_batch(batch, params)
    assert rc == 0
    res, mask = gpu_ransac(lib, batch, params)
    np.testing.assert_array_equal(res["num_iterations"], ores["num_iterations"])
    np.testing.assert_array_equal(res["best_cost"], ores["best_cost"])
    np.testing.assert_array_equal(mask, omask)
    np.testing.assert_array_equal(res["num_inliers"], ores["num_inliers"])
    for f in ("essential_matrix", "rotation", "position"):
        assert np.array_equal(res[f], ores[f], equal_nan=True), f
    params.use_lo = 1
    b = batch.struct()
    assert lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None) == capi.THB_E_UNSUPPORTED


@pytest.mark.parametrize("kind", ["abspose", "homography"])
def test_lmed_homography_and_absolute_pose_match_oracle(lib, oracle, kind):
    """LMED through the other two estimators (P3P: up to 4 models per sample; homography), odd data count."""
    make = synthetic.make_abspose_batch if kind == "abspose" else synthetic.make_homography_batch
    batch, _ = make(12, n=301, inlier_ratio=0.75, seed=9, base_seed=5200)
    def mk(p):
        p = synthetic.c4_params(p); p.ransac_type = 2
        return p
    rc, ores, omask = oracle.ransac_batch(kind, batch, mk(oracle.ransac_default_params()))
    assert rc == 0
    res, mask = gpu_ransac(lib, batch, mk(capi.ThbRansacParams()), kind=kind)
    np.testing.assert_array_equal(res["num_iterations"], ores["num_iterations"])
    np.testing.assert_array_equal(res["best_cost"], ores["best_cost"])
    np.testing.assert_array_equal(mask, omask)
    for f in ("essential_matrix", "rotation", "position"):
        assert np.array_equal(res[f], ores[f], equal_nan=True), f


def test_distributed_front_end_single_rank(lib, oracle):
    """pytheiasfm_b200.distributed.estimate_relative_poses without a process group (world = 1): the same records and masks as the
    direct C-ABI call, through the deal / pack / gather / reassemble path the multi-GPU runs use."""
    from pytheiasfm_b200 import distributed as ptd
    batch, _ = synthetic.make_pair_batch(19, n=300, seed=12, base_seed=900)
    params = synthetic.c4_params(capi.ThbRansacParams())
    records, masks = ptd.estimate_relative_poses(batch, params)
    res, mask = gpu_ransac(lib, batch, params)
    assert records.tobytes() == res.tobytes()
    for i in range(batch.num_pairs):
        np.testing.assert_array_equal(masks[i], mask[batch.pair_offset[i]:batch.pair_offset[i + 1]])
