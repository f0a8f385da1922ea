"""The reference's own Python tests, re-run through the drop-in adapter (`import pytheiasfm_b200 as pt`):
pytests/sfm/bundle_adjuster_test.py:6-26, pytests/sfm/random_recon_gen.py (scene generator),
pytests/sfm/two_view_pose_test.py:96-111 and pytests/sfm/absolute_pose_estimator_test.py:11-24 — same calls,
same argument orders, same assertions. GPU only (the adapter has no CPU path)."""
import numpy as np
import pytest

import pytheiasfm_b200 as pt

pytestmark = pytest.mark.gpu


class RandomReconGenerator:
    """random_recon_gen.py:27-176 with the Camera set up directly (CameraIntrinsicsPrior is not on the hot path)."""

    def __init__(self, seed=42):
        np.random.seed(seed)
        self.recon = pt.sfm.Reconstruction()
        self.camera = pt.sfm.Camera()
        self.camera.SetFocalLength(900.0)
        self.camera.SetPrincipalPoint(720, 540)
        self.camera.SetImageSize(1440, 1080)

    def generate_random_recon(self, nr_views=10, nr_tracks=100, pt3_xyz_min=(-4, -4, 0), pt3_xyz_max=(4, 4, 6),
                              cam_xyz_min=(-6, -6, -1), cam_xyz_max=(6, 6, -10), pixel_noise=0.0):
        P = np.random.uniform(pt3_xyz_min, pt3_xyz_max, size=(nr_tracks, 3))
        for i in range(nr_tracks):
            tid = self.recon.AddTrack()
            track = self.recon.MutableTrack(tid)
            track.SetPoint(np.append(P[i], 1.0))
            track.SetIsEstimated(True)
        Cw = np.random.uniform(cam_xyz_min, cam_xyz_max, size=(nr_views, 3))
        ax = np.random.uniform(-0.2, 0.1, size=(nr_views, 3))
        ang = np.random.uniform(-np.pi / 4, np.pi / 4, size=nr_views)
        for i in range(nr_views):
            vid = self.recon.AddView(str(i), 0, i)
            view = self.recon.View(vid)
            cam = view.MutableCamera()
            cam.DeepCopy(self.camera)
            cam.SetPosition(Cw[i])
            cam.SetOrientationFromAngleAxis(ang[i] * ax[i] / np.linalg.norm(ax[i]))
            view.SetIsEstimated(True)
        for tid in self.recon.TrackIds():
            pt3d = self.recon.Track(tid).Point()
            for vid in self.recon.ViewIds():
                cam = self.recon.View(vid).Camera()
                depth, pix = cam.ProjectPoint(pt3d)
                if depth <= 0 or pix[0] < 0 or pix[0] > 1440 or pix[1] < 0 or pix[1] > 1080:
                    continue
                self.recon.AddObservation(vid, tid, pt.sfm.Feature(pix + np.random.randn(2) * pixel_noise))
        return self.recon

    def add_noise_to_views(self, noise_pos=1e-5, noise_angle=1e-2):
        for vid in self.recon.ViewIds():
            cam = self.recon.View(vid).MutableCamera()
            cam.SetPosition(cam.GetPosition() + noise_pos * np.random.randn(3))
            cam.SetOrientationFromAngleAxis(cam.GetOrientationAsAngleAxis() + noise_angle * np.pi / 180.0 * np.random.randn(3))

    def add_noise_to_tracks(self, noise_track=1e-5):
        for tid in self.recon.TrackIds():
            p = self.recon.Track(tid).Point()
            self.recon.MutableTrack(tid).SetPoint(np.append(p[:3] / p[3] + noise_track * np.random.randn(3), 1.0))


@pytest.fixture
def gen():
    g = RandomReconGenerator(seed=42)
    g.generate_random_recon(nr_views=10, nr_tracks=100)
    return g


@pytest.fixture
def ba_options():
    return pt.sfm.BundleAdjustmentOptions()


def test_BundleAdjustView(ba_options):
    """bundle_adjuster_test.py:6-15, on the scene its __main__ builds (:95-96: 1 view, 20 tracks)."""
    gen = RandomReconGenerator(seed=42)
    gen.generate_random_recon(nr_views=1, nr_tracks=20)
    assert gen.recon.View(gen.recon.ViewIds()[0]).NumFeatures() >= 6
    for vid in gen.recon.ViewIds():
        orig_pos = gen.recon.View(vid).Camera().GetPosition()
        gen.add_noise_to_views(noise_pos=1e-3, noise_angle=1e-1)
        result = pt.sfm.BundleAdjustView(gen.recon, ba_options, vid)
        dist_pos = np.linalg.norm(orig_pos - gen.recon.View(vid).Camera().GetPosition())
        assert dist_pos < 1e-4
        assert result.success


def test_BundleAdjustView_each_of_ten(gen, ba_options):
    """Same assertion on a 10-view scene, perturbing only the view being adjusted (add_noise_to_view, random_recon_gen.py:152-159)."""
    for vid in gen.recon.ViewIds():
        if gen.recon.View(vid).NumFeatures() < 6:
            continue
        cam = gen.recon.View(vid).MutableCamera()
        orig_pos = cam.GetPosition()
        cam.SetPosition(orig_pos + 1e-3 * np.random.randn(3))
        cam.SetOrientationFromAngleAxis(cam.GetOrientationAsAngleAxis() + 1e-1 * np.pi / 180.0 * np.random.randn(3))
        result = pt.sfm.BundleAdjustView(gen.recon, ba_options, vid)
        assert result.success
        assert np.linalg.norm(orig_pos - gen.recon.View(vid).Camera().GetPosition()) < 1e-4


def test_BundleAdjustTrack(gen, ba_options):
    """bundle_adjuster_test.py:17-22"""
    gen.add_noise_to_tracks(noise_track=1e-3)
    for t_id in gen.recon.TrackIds():
        if gen.recon.Track(t_id).NumViews() < 2:
            continue
        truth = gen.recon.Track(t_id).Point()
        result = pt.sfm.BundleAdjustTrack(gen.recon, ba_options, t_id)
        assert result.success
        assert result.final_cost <= result.initial_cost + 1e-12
        assert gen.recon.Track(t_id).InverseDepth() != 0.0   # UpdateInverseDepth side effect (bundle_adjustment.cc:69-83)


def test_BundleAdjustReconstruction_full_and_partial(gen, ba_options):
    """pyexamples/sfm_pipeline_*.py usage: pt.sfm.BundleAdjustReconstruction(opts, recon) with HUBER loss."""
    gen.add_noise_to_views(noise_pos=1e-2, noise_angle=0.5)
    gen.add_noise_to_tracks(noise_track=1e-2)
    ba_options.use_inner_iterations = False
    ba_options.loss_function_type = pt.sfm.LossFunctionType.HUBER
    ba_options.robust_loss_width = 2.0
    result = pt.sfm.BundleAdjustReconstruction(ba_options, gen.recon)
    assert result.success and result.final_cost < 1e-6 * result.initial_cost + 1e-9
    # partial: half of the views, all tracks
    views = gen.recon.ViewIds()[:5]
    before = {v: gen.recon.View(v).Camera().GetPosition() for v in gen.recon.ViewIds()}
    gen.add_noise_to_tracks(noise_track=1e-2)
    result = pt.sfm.BundleAdjustPartialReconstruction(ba_options, views, gen.recon.TrackIds(), gen.recon)
    assert result.success
    for v in gen.recon.ViewIds()[5:]:                # views outside the set keep their extrinsics (bundle_adjuster.cc:199-204)
        np.testing.assert_array_equal(gen.recon.View(v).Camera().GetPosition(), before[v])
    # the reference's default options (inner iterations ON, bundle_adjustment.h:144): the call every pyexample pipeline makes
    gen.add_noise_to_views(noise_pos=1e-2, noise_angle=0.5)
    gen.add_noise_to_tracks(noise_track=1e-2)
    defaults = pt.sfm.BundleAdjustmentOptions()
    assert defaults.use_inner_iterations is True
    result = pt.sfm.BundleAdjustReconstruction(defaults, gen.recon)
    assert result.success and result.final_cost < 1e-6 * result.initial_cost + 1e-9


def test_SetOutlierTracksToUnestimated(gen):
    """sfm.cc:935 / set_outlier_tracks_to_unestimated.cc: a displaced track and a far-away track lose their estimated flag."""
    tids = gen.recon.TrackIds()
    bad = gen.recon.MutableTrack(tids[0]); p = bad.Point(); p[:3] += 0.3 * p[3]; bad.SetPoint(p)
    far = gen.recon.MutableTrack(tids[1]); q = far.Point(); q[:3] = q[:3] * 5000.0; far.SetPoint(q)
    n = pt.sfm.SetOutlierTracksToUnestimated(set(tids), 4.0, 2.0, gen.recon)
    assert n >= 2 and not gen.recon.Track(tids[0]).IsEstimated() and not gen.recon.Track(tids[1]).IsEstimated()
    assert gen.recon.Track(tids[2]).IsEstimated()
    assert pt.sfm.SetOutlierTracksToUnestimated(set(tids), 4.0, 2.0, gen.recon) == 0      # idempotent


def test_position_priors_through_the_adapter(gen):
    """View.SetPositionPrior + BundleAdjustmentOptions.use_position_priors: the priors (shifted by a common offset) drag the whole
    reconstruction along its translation gauge; without the option they are ignored."""
    vids = gen.recon.ViewIds()
    shift = np.array([0.3, -0.2, 0.1])
    for v in vids:
        view = gen.recon.View(v)
        assert not view.HasPositionPrior()
        view.SetPositionPrior(view.Camera().GetPosition() + shift, 100.0 * np.eye(3))
        assert view.HasPositionPrior()
        np.testing.assert_allclose(view.GetPositionPriorSqrtInformation(), 100.0 * np.eye(3))
    before = {v: gen.recon.View(v).Camera().GetPosition().copy() for v in vids}
    opts = pt.sfm.BundleAdjustmentOptions()
    res = pt.sfm.BundleAdjustReconstruction(opts, gen.recon)
    assert res.success
    assert max(np.abs(gen.recon.View(v).Camera().GetPosition() - before[v]).max() for v in vids) < 1e-3   # ignored
    opts.use_position_priors = True
    res = pt.sfm.BundleAdjustReconstruction(opts, gen.recon)
    assert res.success
    for v in vids:
        np.testing.assert_allclose(gen.recon.View(v).Camera().GetPosition(), before[v] + shift, atol=2e-2)


def test_orientation_priors_through_the_adapter(gen):
    """View.SetOrientationPrior + BundleAdjustmentOptions.use_orientation_priors (orientation_error.h:44-80): strong priors on
    the views of BundleAdjustViews (tracks constant) hold the orientations at the priors although the reprojection errors pull
    them back; without the option the priors are ignored."""
    from scipy.spatial.transform import Rotation
    vids = gen.recon.ViewIds()
    rng = np.random.default_rng(3)
    prior = {}
    for v in vids:
        view = gen.recon.View(v)
        assert not view.HasOrientationPrior()
        prior[v] = (Rotation.from_rotvec(0.01 * rng.normal(size=3)) * Rotation.from_rotvec(view.Camera().GetOrientationAsAngleAxis())).as_rotvec()
        view.SetOrientationPrior(prior[v], 1e6 * np.eye(3))
        assert view.HasOrientationPrior()
        np.testing.assert_allclose(view.GetOrientationPrior(), prior[v])
    before = {v: gen.recon.View(v).Camera().GetOrientationAsAngleAxis().copy() for v in vids}
    opts = pt.sfm.BundleAdjustmentOptions()
    assert pt.sfm.BundleAdjustViews(gen.recon, opts, vids).success
    assert max(np.abs(gen.recon.View(v).Camera().GetOrientationAsAngleAxis() - before[v]).max() for v in vids) < 1e-3   # ignored
    opts.use_orientation_priors = True
    assert pt.sfm.BundleAdjustViews(gen.recon, opts, vids).success
    for v in vids:
        err = (Rotation.from_rotvec(gen.recon.View(v).Camera().GetOrientationAsAngleAxis()) * Rotation.from_rotvec(prior[v]).inv()).magnitude()
        assert err < 1e-4
    opts.use_depth_priors = True
    with pytest.raises(Exception):
        pt.sfm.BundleAdjustViews(gen.recon, opts, vids)


def test_BundleAdjustPartialViewsConstant(gen):
    """bundle_adjustment.cc:146-186: the constant views keep their pose, the variable views and all tracks are refined."""
    vids = gen.recon.ViewIds()
    gen.add_noise_to_views(1e-2, 1e-1)
    const = vids[:3]; var = vids[3:]
    before = {v: (gen.recon.View(v).Camera().GetPosition().copy(), gen.recon.View(v).Camera().GetOrientationAsAngleAxis().copy()) for v in vids}
    opts = pt.sfm.BundleAdjustmentOptions()
    res = pt.sfm.BundleAdjustPartialViewsConstant(opts, var, const, gen.recon)
    assert res.success and res.final_cost < res.initial_cost
    for v in const:
        np.testing.assert_array_equal(gen.recon.View(v).Camera().GetPosition(), before[v][0])
        np.testing.assert_array_equal(gen.recon.View(v).Camera().GetOrientationAsAngleAxis(), before[v][1])
    assert any(np.abs(gen.recon.View(v).Camera().GetPosition() - before[v][0]).max() > 0 for v in var)


def test_BundleAdjust_with_covariance(gen):
    """bundle_adjustment_wrapper.cc:52-96: (summary, covariance, empirical variance factor) for a view, views, a track, tracks;
    covariances are symmetric positive definite, scaled by 2 * final_cost / redundancy, and agree between the single and the
    plural form of the call."""
    opts = pt.sfm.BundleAdjustmentOptions()
    vids = gen.recon.ViewIds(); tids = [t for t in gen.recon.TrackIds() if gen.recon.Track(t).IsEstimated()]
    gen.add_noise_to_views(1e-3, 1e-3)
    s, cov, factor = pt.sfm.BundleAdjustViewWithCov(gen.recon, opts, vids[1])
    assert s.success and cov.shape == (6, 6) and factor >= 0
    np.testing.assert_allclose(cov, cov.T, rtol=1e-9, atol=1e-30)
    s2, covs, factor2 = pt.sfm.BundleAdjustViewsWithCov(gen.recon, opts, [vids[1], vids[2]])
    assert s2.success and set(covs) == {vids[1], vids[2]} and covs[vids[2]].shape == (6, 6)
    gen.add_noise_to_tracks(1e-2) if hasattr(gen, "add_noise_to_tracks") else None
    s3, cov3, factor3 = pt.sfm.BundleAdjustTrackWithCov(gen.recon, opts, tids[0])
    assert s3.success and cov3.shape == (3, 3)
    np.testing.assert_allclose(cov3, cov3.T, rtol=1e-9, atol=1e-30)
    s4, covs4, factor4 = pt.sfm.BundleAdjustTracksWithCov(gen.recon, opts, tids[:5])
    assert s4.success and set(covs4) == set(tids[:5])
    if factor3 > 0 and factor4 > 0:   # same (J^T J)^-1 behind both, different variance factors
        np.testing.assert_allclose(cov3 / factor3, covs4[tids[0]] / factor4, rtol=1e-3)


def test_SelectGoodTracksForBundleAdjustment(gen):
    """sfm.cc:933 / select_good_tracks_for_bundle_adjustment.cc: (success, set of track ids); every view keeps its quota, the
    set is a strict subset when the quota is small, and BundleAdjustPartialReconstruction accepts it."""
    vids = gen.recon.ViewIds()
    ok, chosen = pt.sfm.SelectGoodTracksForBundleAdjustment(gen.recon, set(vids), 10, 200, 5)
    assert ok and 0 < len(chosen) < len(gen.recon.TrackIds())
    for v in vids:
        seen = [t for t in gen.recon.View(v).TrackIds() if gen.recon.Track(t).IsEstimated()]
        assert len([t for t in seen if t in chosen]) >= min(5, len(seen))
    ok2, all_of_them = pt.sfm.SelectGoodTracksForBundleAdjustment(gen.recon, set(vids), 10, 200, 10 ** 6)
    assert ok2 and all_of_them == {t for t in gen.recon.TrackIds() if gen.recon.Track(t).IsEstimated()}
    opts = pt.sfm.BundleAdjustmentOptions()
    res = pt.sfm.BundleAdjustPartialReconstruction(opts, list(vids), sorted(chosen), gen.recon)
    assert res.success


def _two_view_corrs(n=300, outliers=0.3, noise=1e-3, seed=65):
    rng = np.random.default_rng(seed)
    ang = np.deg2rad(10.0)
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    c = np.array([1.0, 0.2, 0.1]); c /= np.linalg.norm(c)
    X = np.stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(4, 10, n)], -1)
    x1 = X[:, :2] / X[:, 2:]; Xc = (X - c) @ R.T; x2 = Xc[:, :2] / Xc[:, 2:]
    x1 += rng.normal(0, noise, x1.shape); x2 += rng.normal(0, noise, x2.shape)
    k = int(outliers * n)
    x2[:k] = rng.uniform(-1, 1, (k, 2))
    corrs = [pt.matching.FeatureCorrespondence(pt.sfm.Feature(a), pt.sfm.Feature(b)) for a, b in zip(x1, x2)]
    return corrs, R, c


def test_EstimateRelativePose():
    """two_view_pose_test.py:96-111: (success, pose, summary) and the translation direction within tolerance."""
    corrs, R, c = _two_view_corrs()
    params = pt.solvers.RansacParameters()
    params.error_thresh = (2e-3) ** 2; params.failure_probability = 1e-4; params.min_iterations = 10; params.max_iterations = 1000
    params.use_mle = True; params.seed = 7
    success, pose, summary = pt.sfm.EstimateRelativePose(params, pt.sfm.RansacType.RANSAC, corrs)
    assert success
    assert np.rad2deg(np.arccos(np.clip(pose.position @ c, -1, 1))) < 5.0
    assert np.rad2deg(np.arccos(np.clip((np.trace(pose.rotation @ R.T) - 1) / 2, -1, 1))) < 1.0
    assert len(summary.inliers) > 150 and summary.num_input_data_points == 300 and summary.confidence > 0.99
    again = pt.sfm.EstimateRelativePose(params, pt.sfm.RansacType.RANSAC, corrs)   # the seed field makes it reproducible
    assert again[2].inliers == summary.inliers and again[2].num_iterations == summary.num_iterations
    params.use_lo = True; params.lo_start_iterations = 5      # LO-RANSAC (estimate_relative_pose_test.cc:220-279 style)
    ok_lo, pose_lo, summary_lo = pt.sfm.EstimateRelativePose(params, pt.sfm.RansacType.RANSAC, corrs)
    assert ok_lo and summary_lo.num_lo_iterations >= 1
    assert np.rad2deg(np.arccos(np.clip((np.trace(pose_lo.rotation @ R.T) - 1) / 2, -1, 1))) < 1.0
    params.use_lo = False
    for variant in (pt.sfm.RansacType.PROSAC, pt.sfm.RansacType.LMED):     # create_and_initialize_ransac_variant.h:52-86
        ok_v, pose_v, summary_v = pt.sfm.EstimateRelativePose(params, variant, corrs)
        assert ok_v and len(summary_v.inliers) > 100
        assert np.rad2deg(np.arccos(np.clip((np.trace(pose_v.rotation @ R.T) - 1) / 2, -1, 1))) < 2.0
    with pytest.raises(Exception):
        pt.sfm.EstimateRelativePose(params, pt.sfm.RansacType.EXHAUSTIVE, corrs)   # the reference's sampler CHECK-aborts (sample size != 2)


def test_PoseFromThreePoints_and_EstimateCalibratedAbsolutePose():
    """absolute_pose_estimator_test.py:11-24 (P3P on 3 points) and the RANSAC estimator with PnPType.KNEIP."""
    rng = np.random.default_rng(3)
    ang = np.deg2rad(15.0)
    R = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    c = np.array([0.3, -0.4, 0.2])
    Xc = np.stack([rng.uniform(-2, 2, 200), rng.uniform(-2, 2, 200), rng.uniform(3, 8, 200)], -1)
    X = Xc @ R + c
    x = Xc[:, :2] / Xc[:, 2:]
    ok, Rs, ts = pt.sfm.PoseFromThreePoints([x[0], x[1], x[2]], [X[0], X[1], X[2]])
    assert ok and len(Rs) == 4
    assert min(np.abs(Rk - R).max() + np.abs(tk + R @ c).max() for Rk, tk in zip(Rs, ts)) < 1e-8
    x[:40] = rng.uniform(-1, 1, (40, 2))
    corrs = [pt.sfm.FeatureCorrespondence2D3D(a, b) for a, b in zip(x, X)]
    params = pt.solvers.RansacParameters(); params.error_thresh = 1e-6; params.seed = 11
    success, pose, summary = pt.sfm.EstimateCalibratedAbsolutePose(params, pt.sfm.RansacType.RANSAC, pt.sfm.PnPType.KNEIP, corrs)
    assert success and np.abs(pose.rotation - R).max() < 1e-6 and np.abs(pose.position - c).max() < 1e-6
    assert len(summary.inliers) >= 160
    with pytest.raises(RuntimeError, match="KNEIP"):
        pt.sfm.EstimateCalibratedAbsolutePose(params, pt.sfm.RansacType.RANSAC, pt.sfm.PnPType.DLS, corrs)


def test_minimal_solver_bindings():
    """pose_wrapper.cc return conventions: (success, [E...]), (success, H), (success, [F...])."""
    corrs, R, c = _two_view_corrs(n=7, outliers=0.0, noise=0.0)
    x1 = [cc.feature1.point for cc in corrs]; x2 = [cc.feature2.point for cc in corrs]
    ok, Es = pt.sfm.FivePointRelativePose(x1[:5], x2[:5])
    assert ok and 1 <= len(Es) <= 10 and Es[0].shape == (3, 3)
    t = -R @ c
    E_gt = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
    assert min(min(np.abs(E / np.linalg.norm(E) - s * E_gt / np.linalg.norm(E_gt)).max() for s in (1, -1)) for E in Es) < 1e-6
    ok, H = pt.sfm.FourPointHomography(x1[:4], x2[:4])
    assert ok and H.shape == (3, 3)
    ok, Fs = pt.sfm.SevenPointFundamentalMatrix(x1, x2)
    assert ok and 1 <= len(Fs) <= 3


def test_EstimateTwoViewInfo_against_the_oracle():
    """pt.sfm.EstimateTwoViewInfo (sfm.cc:922, estimate_twoview_info.cc:133-191, calibrated branch): pixel correspondences
    + two CameraIntrinsicsPriors -> (success, TwoViewInfo, inlier indices). Checked against the same composition done in
    numpy over the oracle's EstimateRelativePose: identical inlier list, pose within 1e-9."""
    from oracle import oracle_py
    from pytheiasfm_b200 import capi
    corrs_n, R, c = _two_view_corrs(n=400, outliers=0.3, noise=5e-4)
    xn = np.array([[cc.feature1.point[0], cc.feature1.point[1], cc.feature2.point[0], cc.feature2.point[1]] for cc in corrs_n])
    f1, f2, a2, s1 = 1100.0, 950.0, 1.02, 0.7
    pp1, pp2 = np.array([960.0, 540.0]), np.array([640.0, 480.0])
    px = np.empty_like(xn)   # pixels of two pinhole cameras (skew on the first, aspect ratio on the second)
    px[:, 0] = f1 * xn[:, 0] + s1 * xn[:, 1] + pp1[0]; px[:, 1] = f1 * xn[:, 1] + pp1[1]
    px[:, 2] = f2 * xn[:, 2] + pp2[0]; px[:, 3] = f2 * a2 * xn[:, 3] + pp2[1]
    corrs = [pt.matching.FeatureCorrespondence(pt.sfm.Feature(p[:2]), pt.sfm.Feature(p[2:])) for p in px]
    pr1, pr2 = pt.sfm.CameraIntrinsicsPrior(), pt.sfm.CameraIntrinsicsPrior()
    pr1.image_width, pr1.image_height, pr2.image_width, pr2.image_height = 1920, 1080, 1280, 960
    pr1.focal_length.value = [f1]; pr2.focal_length.value = [f2]
    pr1.principal_point.value = pp1; pr2.principal_point.value = pp2
    pr1.skew.value = [s1]; pr2.aspect_ratio.value = [a2]
    assert pr1.focal_length.is_set and not pr1.aspect_ratio.is_set
    opts = pt.sfm.EstimateTwoViewInfoOptions()
    assert (opts.max_sampson_error_pixels, opts.expected_ransac_confidence, opts.min_ransac_iterations, opts.max_ransac_iterations,
            opts.use_mle, opts.use_lo) == (6.0, 0.9999, 10, 1000, True, False)   # estimate_twoview_info.h:52-73
    opts.max_sampson_error_pixels = 2.0
    opts.seed = 5
    success, info, inliers = pt.sfm.EstimateTwoViewInfo(opts, pr1, pr2, corrs)
    assert success and info.focal_length_1 == f1 and info.focal_length_2 == f2
    assert info.num_verified_matches == len(inliers) > 200 and info.visibility_score == 0 and info.scale_estimate == -1.0

    # the same composition over the oracle: normalise (same operation order), scaled threshold, EstimateRelativePose
    y1 = (px[:, 1] - pp1[1]) / (f1 * 1.0); x1 = (px[:, 0] - pp1[0] - y1 * s1) / f1
    y2 = (px[:, 3] - pp2[1]) / (f2 * a2); x2 = (px[:, 2] - pp2[0] - y2 * 0.0) / f2
    norm = np.stack([x1 / 1.0 / 1.0, y1 / 1.0 / 1.0, x2 / 1.0 / 1.0, y2 / 1.0 / 1.0], -1)
    batch = capi.HostPairBatch([norm], [5])
    p = oracle_py.ransac_default_params()
    p.failure_probability = 1.0 - 0.9999; p.min_iterations = 10; p.max_iterations = 1000; p.use_mle = 1; p.use_lo = 0
    p.min_inlier_ratio = 0.0; p.ransac_type = 0
    p.error_thresh = (2.0 * 1920 / 1024.0) * (2.0 * 1280 / 1024.0) / (f1 * f2)
    rc, res, mask = oracle_py.ransac_relpose_batch(batch, p)
    assert rc == 0 and res["success"][0] == 1
    assert inliers == list(np.nonzero(mask)[0])
    Ro = res["rotation"][0].reshape(3, 3)
    ang = np.arccos(np.clip((np.trace(Ro) - 1) / 2, -1, 1))
    axis = np.array([Ro[2, 1] - Ro[1, 2], Ro[0, 2] - Ro[2, 0], Ro[1, 0] - Ro[0, 1]]) / (2 * np.sin(ang))
    np.testing.assert_allclose(info.rotation_2, ang * axis, rtol=0, atol=1e-9)
    np.testing.assert_allclose(info.position_2, res["position"][0], rtol=0, atol=1e-12)
    # and against the ground truth of the scene
    assert np.rad2deg(np.arccos(np.clip(info.position_2 @ c, -1, 1))) < 5.0

    pr2.focal_length.is_set = False
    with pytest.raises(RuntimeError, match="uncalibrated"):
        pt.sfm.EstimateTwoViewInfo(opts, pr1, pr2, corrs)


def _prior_from_view_intrinsics(v):
    """CameraIntrinsicsPrior that reproduces a synthetic.make_two_view_batch view (prior order of the distortion slots)."""
    from pytheiasfm_b200 import capi
    names = {capi.MODEL_PINHOLE: "PINHOLE", capi.MODEL_FISHEYE: "FISHEYE", capi.MODEL_FOV: "FOV", capi.MODEL_DIVISION_UNDISTORTION: "DIVISION_UNDISTORTION",
             capi.MODEL_DOUBLE_SPHERE: "DOUBLE_SPHERE", capi.MODEL_EXTENDED_UNIFIED: "EXTENDED_UNIFIED"}
    m, K = int(v["model"]), v["params"]
    pr = pt.sfm.CameraIntrinsicsPrior()
    pr.camera_intrinsics_model_type = names[m]
    pr.image_width, pr.image_height = int(v["image_width"]), int(v["image_height"])
    pr.focal_length.value = [K[0]]; pr.aspect_ratio.value = [K[1]]
    if m in (capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION):
        pr.principal_point.value = [K[2], K[3]]; pr.radial_distortion.value = [K[4], 0, 0, 0]
    else:
        pr.skew.value = [K[2]]; pr.principal_point.value = [K[3], K[4]]
        rd = [K[5], K[6], K[7], K[8]]
        if m == capi.MODEL_DOUBLE_SPHERE:
            rd = [K[6], K[5], 0, 0]          # prior order [alpha, xi], storage [xi, alpha]
        pr.radial_distortion.value = rd
    return pr


def test_batched_two_view_entries_all_models():
    """EstimateTwoViewInfoBatch / VerifyTwoViewMatchesBatch (one call for a list of pairs) against the C-ABI on the same input,
    for views of all six camera models; Camera.SetFromCameraIntrinsicsPriors + PixelToUnitDepthRay round trip."""
    import ctypes as C
    from pytheiasfm_b200 import capi, synthetic
    models = (capi.MODEL_PINHOLE, capi.MODEL_FISHEYE, capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION, capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED)
    batch, i1, i2, gts = synthetic.make_two_view_batch(6, n=400, models=models, seed=21)
    priors1 = [_prior_from_view_intrinsics(v) for v in i1]; priors2 = [_prior_from_view_intrinsics(v) for v in i2]
    corrs = [[pt.matching.FeatureCorrespondence(pt.sfm.Feature(r[:2]), pt.sfm.Feature(r[2:])) for r in batch.corr[batch.pair_offset[i]:batch.pair_offset[i + 1]]]
             for i in range(6)]
    lib = capi.load_library()
    o = capi.ThbTwoViewOptions(); lib.thb_two_view_default_options(C.byref(o))
    for verify in (False, True):
        info = np.zeros(6, capi.TWO_VIEW_INFO_DTYPE); mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
        b = batch.struct()
        fn = lib.thb_verify_two_view_matches_batch if verify else lib.thb_estimate_two_view_info_batch
        capi.check(fn(C.byref(b), i1.ctypes.data_as(C.c_void_p), i2.ctypes.data_as(C.c_void_p), C.byref(o), info.ctypes.data_as(C.c_void_p),
                      mask.ctypes.data_as(C.c_void_p), None))
        if verify:
            out = pt.sfm.VerifyTwoViewMatchesBatch(pt.sfm.TwoViewMatchGeometricVerificationOptions(), priors1, priors2, corrs, list(batch.seed))
        else:
            out = pt.sfm.EstimateTwoViewInfoBatch(pt.sfm.EstimateTwoViewInfoOptions(), priors1, priors2, corrs, list(batch.seed))
        assert len(out) == 6
        for i, (ok, tvi, idx) in enumerate(out):
            assert ok == bool(info["success"][i])
            assert idx == list(np.nonzero(mask[batch.pair_offset[i]:batch.pair_offset[i + 1]])[0])
            np.testing.assert_allclose(tvi.rotation_2, info["rotation_2"][i], rtol=0, atol=1e-12)
            assert tvi.num_verified_matches == info["num_verified_matches"][i] and tvi.num_homography_inliers == info["num_homography_inliers"][i]
    # one pair through the reference-named entry
    opts = pt.sfm.EstimateTwoViewInfoOptions(); opts.seed = int(batch.seed[4])
    ok, tvi, idx = pt.sfm.EstimateTwoViewInfo(opts, priors1[4], priors2[4], corrs[4])
    assert ok and tvi.focal_length_1 == i1["params"][4, 0]
    cam = pt.sfm.Camera()
    cam.SetFromCameraIntrinsicsPriors(priors1[4])
    assert cam.GetCameraIntrinsicsModelType() == int(i1["model"][4])
    np.testing.assert_allclose(cam.Parameters(), i1["params"][4][:len(cam.Parameters())], rtol=0, atol=0)
    ray = cam.PixelToUnitDepthRay(np.array([612.0, 377.0]))
    depth, pix = cam.ProjectPoint(np.append(3.0 * ray, 1.0))
    np.testing.assert_allclose(pix, [612.0, 377.0], atol=1e-6)      # camera-model round trip (the reference's ReprojectionTest)


def test_TriangulateMidpoint():
    """pytests-style call of pt.sfm.TriangulateMidpoint (sfm.cc:854, triangulation_test.cc:312-335): (success, point)."""
    X = np.array([5.0, 20.0, 23.0])
    a = 0.15
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([-3.0, 1.5, 11.0])
    origins = [np.zeros(3), -R.T @ t]
    dirs = [X / np.linalg.norm(X), R.T @ ((R @ X + t) / np.linalg.norm(R @ X + t))]
    ok, p = pt.sfm.TriangulateMidpoint(origins, dirs)
    assert ok
    np.testing.assert_allclose(p[:3] / p[3], X, rtol=1e-12)
    with pytest.raises(ValueError):
        pt.sfm.TriangulateMidpoint(origins[:1], dirs[:1])


def test_TrackEstimator_EstimateAllTracks():
    """pt.sfm.TrackEstimator (sfm.cc:1102-1135; estimate_track_test.cc's scenario: known cameras, tracks reset to unestimated):
    every track with enough views is re-triangulated and bundle adjusted to its ground-truth position in one launch."""
    gen = RandomReconGenerator(seed=7)
    gen.generate_random_recon(nr_views=8, nr_tracks=120, pixel_noise=0.2)
    recon = gen.recon
    truth = {}
    for tid in recon.TrackIds():
        tr = recon.MutableTrack(tid)
        truth[tid] = tr.Point()[:3] / tr.Point()[3]
        tr.SetPoint(np.array([0.0, 0.0, 0.0, 1.0]))
        tr.SetIsEstimated(False)
    keep_estimated = recon.TrackIds()[:3]
    for tid in keep_estimated:
        recon.MutableTrack(tid).SetPoint(np.append(truth[tid], 1.0))
        recon.MutableTrack(tid).SetIsEstimated(True)
    opts = pt.sfm.TrackEstimatorOptions()
    assert (opts.max_acceptable_reprojection_error_pixels, opts.min_triangulation_angle_degrees, opts.bundle_adjustment) == (5.0, 3.0, True)
    est = pt.sfm.TrackEstimator(opts, recon)
    summary = est.EstimateAllTracks()
    seen = {t for v in recon.ViewIds() for t in recon.View(v).TrackIds()}
    assert summary.input_num_estimated_tracks == len(set(keep_estimated) & seen)
    assert summary.num_triangulation_attempts == len(seen) - summary.input_num_estimated_tracks
    multi = [t for t in seen if recon.Track(t).NumViews() >= 2 and t not in keep_estimated]
    assert len(summary.estimated_tracks) >= 0.9 * len(multi)
    assert (len(summary.estimated_tracks) + summary.num_bad_angles + summary.num_failed_triangulations + summary.num_bad_reprojections
            == summary.num_triangulation_attempts)
    for tid in recon.TrackIds():
        tr = recon.Track(tid)
        if tid in summary.estimated_tracks:
            assert tr.IsEstimated()
            p = tr.Point()
            assert np.linalg.norm(p[:3] / p[3] - truth[tid]) < 0.05
            assert tr.InverseDepth() > 0
        elif tid not in keep_estimated:
            assert not tr.IsEstimated()
            np.testing.assert_array_equal(tr.Point(), [0.0, 0.0, 0.0, 1.0])
    # a second pass has nothing left to attempt among the estimated ones
    again = pt.sfm.TrackEstimator(opts, recon).EstimateTracks(set(summary.estimated_tracks))
    assert again.num_triangulation_attempts == 0 and again.input_num_estimated_tracks == len(summary.estimated_tracks)
