"""GPU parity tests (run with -m gpu on a B200): batched EstimateTwoViewInfo / TwoViewMatchGeometricVerification::VerifyMatches
from PIXEL correspondences (thb_estimate_two_view_info_batch, thb_verify_two_view_matches_batch) against the oracle's
per-pair composition (oracle/ransac_oracle.cc::TwoViewPair)."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi, synthetic

pytestmark = pytest.mark.gpu

SIX = (capi.MODEL_PINHOLE, capi.MODEL_FISHEYE, capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION, capi.MODEL_DOUBLE_SPHERE,
       capi.MODEL_EXTENDED_UNIFIED)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def gpu_two_view(lib, batch, i1, i2, opts, verify):
    info = np.zeros(batch.num_pairs, capi.TWO_VIEW_INFO_DTYPE)
    mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
    b = batch.struct()
    fn = lib.thb_verify_two_view_matches_batch if verify else lib.thb_estimate_two_view_info_batch
    capi.check(fn(C.byref(b), _vp(i1), _vp(i2), C.byref(opts), _vp(info), _vp(mask), None))
    return info, mask


def default_opts(lib):
    o = capi.ThbTwoViewOptions()
    lib.thb_two_view_default_options(C.byref(o))
    return o


def test_default_options_are_the_references(lib):
    o = default_opts(lib)  # estimate_twoview_info.h:51-80, two_view_match_geometric_verification.h:53-92
    assert (o.max_sampson_error_pixels, o.expected_ransac_confidence, o.min_ransac_iterations, o.max_ransac_iterations) == (6.0, 0.9999, 10, 1000)
    assert (o.use_mle, o.use_lo, o.lo_start_iterations) == (1, 0, 10)
    assert (o.min_num_inlier_matches, o.bundle_adjustment) == (30, 1)
    assert (o.triangulation_max_reprojection_error, o.min_triangulation_angle_degrees, o.final_max_reprojection_error) == (15.0, 4.0, 5.0)


@pytest.mark.parametrize("use_lo", [0, 1])
def test_estimate_two_view_info_batch_matches_oracle(lib, oracle, use_lo):
    """All six camera models (views of a pair use different models), per-pair thresholds from image size and focal lengths."""
    batch, i1, i2, gts = synthetic.make_two_view_batch(18, n=500, models=SIX, seed=3)
    i1["image_width"][::2] = 2048  # resolution-scaled thresholds differ between pairs
    o = default_opts(lib); o.use_lo = use_lo
    info, mask = gpu_two_view(lib, batch, i1, i2, o, False)
    rc, oinfo, omask = oracle.two_view_batch(batch, i1, i2, o, False)
    assert rc == 0
    for f in ("success", "num_verified_matches", "num_ransac_iterations", "visibility_score"):
        np.testing.assert_array_equal(info[f], oinfo[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    np.testing.assert_allclose(info["rotation_2"], oinfo["rotation_2"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(info["position_2"], oinfo["position_2"], rtol=0, atol=1e-9)
    assert (info["success"] == 1).all() and (info["focal_length_1"] == i1["params"][:, 0]).all()
    good = 0
    for p in range(batch.num_pairs):
        R = synthetic.rotmat_from_rotvec(info["rotation_2"][p])
        good += np.rad2deg(np.arccos(np.clip((np.trace(R @ gts[p][0].T) - 1) / 2, -1, 1))) < 3.0
    assert good >= 15


@pytest.mark.parametrize("models", [(capi.MODEL_PINHOLE,), SIX])
def test_verify_two_view_matches_batch_matches_oracle(lib, oracle, models):
    """VerifyMatches end to end: homography inlier count, relative-pose RANSAC on the same generator, triangulation gates,
    two-view bundle adjustment, reprojection filter."""
    batch, i1, i2, gts = synthetic.make_two_view_batch(16, n=600, models=models, seed=7, planar_fraction=0.3)
    o = default_opts(lib)
    info, mask = gpu_two_view(lib, batch, i1, i2, o, True)
    rc, oinfo, omask = oracle.two_view_batch(batch, i1, i2, o, True)
    assert rc == 0
    if models == (capi.MODEL_PINHOLE,):
        for f in ("success", "num_homography_inliers", "num_ransac_iterations", "num_triangulated", "ba_iterations", "num_verified_matches"):
            np.testing.assert_array_equal(info[f], oinfo[f], err_msg=f)
        np.testing.assert_array_equal(mask, omask)
        np.testing.assert_allclose(info["ba_initial_cost"], oinfo["ba_initial_cost"], rtol=1e-9)
        np.testing.assert_allclose(info["ba_final_cost"], oinfo["ba_final_cost"], rtol=1e-6)
        np.testing.assert_allclose(info["rotation_2"], oinfo["rotation_2"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(info["position_2"], oinfo["position_2"], rtol=0, atol=1e-6)
    else:
        # the inverse models of the other five use tan / atan2 / sqrt, which differ by an ulp between CUDA and glibc: a
        # normalised coordinate that moves by 1e-16 can move a borderline correspondence across the threshold, and with it
        # the adaptive iteration bound by one. Everything downstream is compared to that tolerance.
        np.testing.assert_array_equal(info["success"], oinfo["success"])
        np.testing.assert_array_equal(info["num_homography_inliers"], oinfo["num_homography_inliers"])   # pixel domain: exact
        assert np.abs(info["num_ransac_iterations"] - oinfo["num_ransac_iterations"]).max() <= 3
        same = info["num_ransac_iterations"] == oinfo["num_ransac_iterations"]
        assert same.sum() >= 10
        for p in np.nonzero(same)[0]:
            sl = slice(batch.pair_offset[p], batch.pair_offset[p + 1])
            assert (mask[sl] != omask[sl]).sum() <= 2
            np.testing.assert_allclose(info["rotation_2"][p], oinfo["rotation_2"][p], rtol=0, atol=1e-4)
    assert info["success"].sum() >= 14 and (info["num_homography_inliers"] > 30).all()
    for p in range(batch.num_pairs):  # bundle adjustment lands close to the generating pose
        if not info["success"][p]:
            continue
        R = synthetic.rotmat_from_rotvec(info["rotation_2"][p])
        assert np.rad2deg(np.arccos(np.clip((np.trace(R @ gts[p][0].T) - 1) / 2, -1, 1))) < 1.0
        assert abs(np.linalg.norm(info["position_2"][p]) - 1.0) < 1e-12


def test_verify_gates_and_options(lib, oracle):
    """Too few matches (:117-119), too few RANSAC inliers (:143-146), bundle adjustment off, uncalibrated views rejected."""
    batch, i1, i2, _ = synthetic.make_two_view_batch(4, n=200, seed=11)
    small, j1, j2, _ = synthetic.make_two_view_batch(2, n=20, seed=12)
    mixed = capi.HostPairBatch([batch.corr[batch.pair_offset[i]:batch.pair_offset[i + 1]] for i in range(4)] +
                               [small.corr[small.pair_offset[i]:small.pair_offset[i + 1]] for i in range(2)], np.arange(6) + 77)
    a1 = np.concatenate([i1, j1]); a2 = np.concatenate([i2, j2])
    o = default_opts(lib)
    info, mask = gpu_two_view(lib, mixed, a1, a2, o, True)
    rc, oinfo, omask = oracle.two_view_batch(mixed, a1, a2, o, True)
    for f in ("success", "num_verified_matches", "num_homography_inliers"):
        np.testing.assert_array_equal(info[f], oinfo[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    assert list(info["success"][4:]) == [0, 0] and mask[mixed.pair_offset[4]:].sum() == 0
    o.bundle_adjustment = 0
    info2, mask2 = gpu_two_view(lib, mixed, a1, a2, o, True)
    rc, oinfo2, omask2 = oracle.two_view_batch(mixed, a1, a2, o, True)
    np.testing.assert_array_equal(mask2, omask2)
    assert (info2["num_triangulated"] == 0).all() and (info2["num_verified_matches"][:4] >= info["num_verified_matches"][:4]).all()
    a1["focal_length_is_set"][1] = 0
    b = mixed.struct()
    out = np.zeros(6, capi.TWO_VIEW_INFO_DTYPE)
    assert lib.thb_verify_two_view_matches_batch(C.byref(b), _vp(a1), _vp(a2), C.byref(o), _vp(out), None, None) == capi.THB_E_UNSUPPORTED


def test_verify_c5_pairs_generator_state_after_early_termination(lib, oracle):
    """C5 pairs (box faces: the homography RANSAC ends inside a batch of 128 draws): the relative-pose RANSAC must continue
    from the generator state the SEQUENTIAL loop leaves behind, not from the batch's read-ahead."""
    sc = synthetic.config_c5(num_points=20000)
    pairs, intr = sc["pairs"], sc["intrinsics"]
    n = 40
    sub = capi.HostPairBatch([pairs.corr[pairs.pair_offset[i]:pairs.pair_offset[i + 1]] for i in range(n)], pairs.seed[:n])
    ii = np.ascontiguousarray(intr[:n])
    o = default_opts(lib)
    info, mask = gpu_two_view(lib, sub, ii, ii, o, True)
    rc, oinfo, omask = oracle.two_view_batch(sub, ii, ii, o, True)
    for f in ("success", "num_homography_inliers", "num_ransac_iterations", "num_triangulated", "ba_iterations", "num_verified_matches"):
        np.testing.assert_array_equal(info[f], oinfo[f], err_msg=f)
    np.testing.assert_array_equal(mask, omask)
    assert (info["num_ransac_iterations"] % 128 != 0).any()
