"""Batched theia::BundleAdjustTrack (bundle_adjustment.cc:261-285): every track its own trust-region solve with the cameras
constant, one launch for all tracks. Checked track by track against the oracle run on the single-track problem the
reference would build (bundle_adjuster.cc:176-221), and through properties at C5 size."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi, synthetic

pytestmark = pytest.mark.gpu

ALL_MODELS = [capi.MODEL_PINHOLE, capi.MODEL_FISHEYE, capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION,
              capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED]


def gpu_tracks(lib, prob, opts):
    res = np.zeros(prob.num_points, capi.TRACK_BA_DTYPE)
    p = prob.struct()
    capi.check(lib.thb_ba_tracks_batch(C.byref(p), C.byref(opts), res.ctypes.data_as(C.c_void_p), None))
    return res


def single_track_problem(prob, t):
    """What BundleAdjustTrack(options, t, reconstruction) hands to Ceres: the track's observations, every camera constant."""
    a = dict(prob.a)
    keep = prob.a["obs_pt"] == t
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        a[k] = None if prob.a[k] is None else prob.a[k][keep]
    a["cam_const"] = np.full(prob.num_cameras, 3, np.uint8)
    a["intr_const"] = None
    return capi.HostBaProblem(a)


def perturbed(ncam, npts, obs_per_pt, seed, sigma=0.05, **kw):
    prob, _ = synthetic.make_ba_problem(ncam, npts, obs_per_pt, seed=seed, **kw)
    rng = np.random.default_rng(seed + 1000)
    prob.a["pts"][:, :3] += rng.normal(0, sigma, (npts, 3))
    return prob


def check_against_oracle(lib, oracle, prob, opts, tracks):
    pg = prob.copy()
    res = gpu_tracks(lib, pg, opts)
    for t in tracks:
        sub = single_track_problem(prob, t)
        o = oracle.ba_solve(sub, opts)
        assert o["rc"] == 0
        r = res[t]
        assert r["termination_type"] == o["termination_type"], (t, r, o["termination_type"])
        assert abs(r["initial_cost"] - o["initial_cost"]) <= 1e-10 * max(o["initial_cost"], 1e-30)
        assert abs(r["final_cost"] - o["final_cost"]) <= 1e-6 * max(o["final_cost"], 1e-12), (t, r, o["final_cost"])
        assert r["num_iterations"] == o["num_iterations"], (t, r["num_iterations"], o["num_iterations"])
        np.testing.assert_allclose(pg.a["pts"][t], sub.a["pts"][t], rtol=1e-6, atol=1e-9)
    return res, pg


@pytest.mark.parametrize("homogeneous", [True, False])
def test_tracks_match_the_oracle_track_by_track(lib, oracle, homogeneous):
    prob = perturbed(12, 150, 5, seed=301)
    opts = capi.default_options(lib)
    opts.use_inner_iterations = 0
    opts.use_homogeneous_point_parametrization = int(homogeneous)
    res, pg = check_against_oracle(lib, oracle, prob, opts, range(0, 150, 3))
    assert (res["num_iterations"] > 0).all()
    assert (res["final_cost"] < res["initial_cost"]).all()
    # cameras and intrinsics are inputs only
    np.testing.assert_array_equal(pg.a["cam_ext"], prob.a["cam_ext"])


@pytest.mark.parametrize("model", ALL_MODELS)
def test_every_camera_model_and_a_robust_loss(lib, oracle, model):
    prob = perturbed(10, 60, 4, seed=310 + model, sigma=0.02, models=(model,))
    if model == capi.MODEL_DIVISION_UNDISTORTION:
        prob.a["intr"][0, 4] = -5e-7
    opts = capi.default_options(lib)
    opts.use_inner_iterations = 0
    opts.loss_function_type = capi.LOSS_HUBER if model % 2 else capi.LOSS_TRIVIAL
    opts.robust_loss_width = 2.0
    check_against_oracle(lib, oracle, prob, opts, range(0, 60, 4))


def test_constant_and_unobserved_tracks_are_left_alone(lib):
    prob = perturbed(8, 40, 4, seed=320)
    prob.a["pt_const"][::4] = 1
    keep = prob.a["obs_pt"] != 5
    a = dict(prob.a)
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        a[k] = prob.a[k][keep]
    prob = capi.HostBaProblem(a)
    before = prob.a["pts"].copy()
    opts = capi.default_options(lib)
    opts.use_inner_iterations = 0
    res = gpu_tracks(lib, prob, opts)
    skipped = np.zeros(40, bool); skipped[::4] = True; skipped[5] = True
    assert (res["num_iterations"][skipped] == -1).all() and (res["num_iterations"][~skipped] > 0).all()
    np.testing.assert_array_equal(prob.a["pts"][skipped], before[skipped])
    assert (prob.a["pts"][~skipped] != before[~skipped]).any(axis=1).all()
    opts.use_inner_iterations = 1
    p = prob.struct()
    assert lib.thb_ba_tracks_batch(C.byref(p), C.byref(opts), None, None) == capi.THB_E_UNSUPPORTED


def test_c5_sized_batch_properties(lib):
    """128 cameras / 60k tracks / 360k observations (BASELINE configs[4] shape): every track converges, no track's cost goes up,
    the summed cost lands on the chi^2 floor of the pixel noise, a second call is a fixed point."""
    prob, gt = synthetic.make_ba_problem(128, 60000, 6, seed=330, pos_sigma=0.0, rot_sigma=0.0)
    rng = np.random.default_rng(331)
    prob.a["pts"][:, :3] += rng.normal(0, 0.05, (60000, 3))
    opts = capi.default_options(lib)
    opts.use_inner_iterations = 0
    res = gpu_tracks(lib, prob, opts)
    assert (res["termination_type"] == 0).all()
    assert (res["final_cost"] <= res["initial_cost"]).all()
    assert res["final_cost"].sum() < 0.05 * res["initial_cost"].sum()
    again = gpu_tracks(lib, prob, opts)
    np.testing.assert_allclose(again["final_cost"], res["final_cost"], rtol=1e-6, atol=1e-12)
    assert again["num_iterations"].max() <= 2


# ---- TrackEstimator::EstimateTrack (estimate_track.cc:209-321) in one launch --------------------------------------------

def rodrigues(aa):
    th = np.linalg.norm(aa)
    if th < 1e-12:
        return np.eye(3)
    k = aa / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def pinhole_rays(prob):
    """Camera::PixelToUnitDepthRay(feature).normalized() (camera.cc:218-226) for undistorted pinhole cameras."""
    a = prob.a
    rays = np.zeros((prob.num_observations, 3))
    for i in range(prob.num_observations):
        c = a["obs_cam"][i]
        f, ar, s, cx, cy = a["intr"][a["cam_group"][c], :5]
        y = (a["obs_xy"][i, 1] - cy) / (f * ar)
        x = (a["obs_xy"][i, 0] - cx - s * y) / f
        d = rodrigues(a["cam_ext"][c, 3:6]).T @ np.array([x, y, 1.0])
        rays[i] = d / np.linalg.norm(d)
    return rays


def gpu_estimate(lib, prob, rays, eopts, opts):
    status = np.full(prob.num_points, 99, np.int32)
    res = np.zeros(prob.num_points, capi.TRACK_BA_DTYPE)
    p = prob.struct()
    capi.check(lib.thb_estimate_tracks_batch(C.byref(p), rays.ctypes.data_as(C.c_void_p), C.byref(eopts), C.byref(opts),
                                             status.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p), None))
    return status, res


def oracle_estimate_track(oracle, prob, rays, t, eopts, opts):
    """The stages of EstimateTrack composed from the oracle's pieces; returns (status, point or None)."""
    a = prob.a
    idx = np.nonzero(a["obs_pt"] == t)[0]
    if a["pt_const"] is not None and a["pt_const"][t]:
        return capi.TRACK_SKIPPED, None
    d = rays[idx]
    cos_min = np.cos(np.deg2rad(eopts.min_triangulation_angle_degrees))
    if len(idx) < 2 or not any(d[i] @ d[j] < cos_min for i in range(len(idx)) for j in range(i + 1, len(idx))):
        return capi.TRACK_BAD_ANGLE, None
    org = a["cam_ext"][a["obs_cam"][idx], :3]
    X, ok = oracle.triangulate_midpoint_batch(org, d, np.array([0, len(idx)], np.int64))
    if not ok[0]:
        return capi.TRACK_FAILED_TRIANGULATION, None
    X = X[0]
    if eopts.bundle_adjustment:
        start = prob.copy()
        start.a["pts"][t] = X
        sub = single_track_problem(start, t)
        o = oracle.ba_solve(sub, opts)
        if not o["success"]:
            return capi.TRACK_BA_FAILED, None
        X = sub.a["pts"][t]
    err = 0.0
    for i in idx:
        c = a["obs_cam"][i]
        pc = rodrigues(a["cam_ext"][c, 3:6]) @ (X[:3] - X[3] * a["cam_ext"][c, :3])
        if pc[2] / X[3] < 0:
            return capi.TRACK_BAD_REPROJECTION, None
        f, ar, s, cx, cy = a["intr"][a["cam_group"][c], :5]
        n = pc[:2] / pc[2]
        pix = np.array([f * n[0] + s * n[1] + cx, f * ar * n[1] + cy])
        err += np.sum((pix - a["obs_xy"][i]) ** 2)
    if not err / len(idx) < eopts.max_acceptable_reprojection_error_pixels ** 2:
        return capi.TRACK_BAD_REPROJECTION, None
    return capi.TRACK_ESTIMATED, X


@pytest.mark.parametrize("min_angle,with_ba", [(3.0, True), (120.0, True), (3.0, False)])
def test_estimate_tracks_matches_the_composed_oracle(lib, oracle, min_angle, with_ba):
    prob, _ = synthetic.make_ba_problem(12, 160, 4, seed=340, pos_sigma=0.0, rot_sigma=0.0)
    a = prob.a
    a["obs_xy"][np.nonzero(a["obs_pt"] == 7)[0][0]] += 300.0       # an outlier observation: bad reprojection
    a["obs_xy"][np.nonzero(a["obs_pt"] == 11)[0][1]] += 40.0
    a["pt_const"][[3, 50]] = 1                                     # already estimated tracks are skipped
    keep = np.ones(prob.num_observations, bool)
    keep[np.nonzero(a["obs_pt"] == 20)[0][1:]] = False             # a single observation: no angle
    keep[np.nonzero(a["obs_pt"] == 21)[0]] = False                 # no observation at all
    arr = dict(a)
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        arr[k] = a[k][keep]
    prob = capi.HostBaProblem(arr)
    rays = pinhole_rays(prob)
    rng = np.random.default_rng(341)
    prob.a["pts"][:] = rng.normal(0, 50, prob.a["pts"].shape)      # starting values are ignored
    before = prob.a["pts"].copy()
    eopts = capi.ThbTrackEstimatorOptions()
    eopts.min_triangulation_angle_degrees = min_angle
    eopts.bundle_adjustment = int(with_ba)
    opts = capi.default_options(lib)
    opts.use_inner_iterations = 0
    pg = prob.copy()
    status, res = gpu_estimate(lib, pg, rays, eopts, opts)
    want = [oracle_estimate_track(oracle, prob, rays, t, eopts, opts) for t in range(prob.num_points)]
    np.testing.assert_array_equal(status, [w[0] for w in want])
    for t, (st, X) in enumerate(want):
        if st == capi.TRACK_ESTIMATED:
            np.testing.assert_allclose(pg.a["pts"][t], X, rtol=1e-6, atol=1e-9)
        else:
            np.testing.assert_array_equal(pg.a["pts"][t], before[t])
    assert status[3] == status[50] == capi.TRACK_SKIPPED and status[20] == status[21] == capi.TRACK_BAD_ANGLE
    assert status[7] == capi.TRACK_BAD_REPROJECTION
    assert (status == capi.TRACK_ESTIMATED).sum() > (5 if min_angle > 10 else 140)
    if min_angle > 10:
        assert (status == capi.TRACK_BAD_ANGLE).sum() > 5


def test_set_outlier_tracks_batch_matches_oracle(lib, oracle):
    """SetOutlierTracksToUnestimated (set_outlier_tracks_to_unestimated.cc:62-137) for all tracks in one launch: tracks behind
    a camera, with a large mean reprojection error, or seen under too small an angle, for mixed camera models; host and
    device memory."""
    import torch
    prob, gt = synthetic.make_ba_problem(12, 2000, 4, models=(capi.MODEL_PINHOLE, capi.MODEL_DOUBLE_SPHERE, capi.MODEL_FISHEYE), seed=51)
    prob.a["cam_ext"][:] = gt["cam_ext"]; prob.a["pts"][:] = gt["pts"]
    rng = np.random.default_rng(3)
    P = prob.a["pts"]
    P[::9, :3] += rng.normal(0, 0.05, (len(P[::9]), 3))          # large reprojection errors
    far = np.arange(5, 2000, 37)
    P[far, :3] = P[far, :3] * 400.0                                # far away: rays nearly parallel ...
    for i in np.nonzero(np.isin(prob.a["obs_pt"], far))[0]:        # ... but perfectly reprojected: only the angle test fails
        c = prob.a["obs_cam"][i]; g = prob.a["cam_group"][c]
        pc = synthetic.rotmat_from_rotvec(prob.a["cam_ext"][c, 3:]) @ (P[prob.a["obs_pt"][i], :3] - prob.a["cam_ext"][c, :3])
        if pc[2] > 0:
            prob.a["obs_xy"][i] = synthetic.project(int(prob.a["intr_model"][g]), prob.a["intr"][g], pc[None])[0]
    behind = np.arange(7, 2000, 53)
    P[behind, :3] = prob.a["cam_ext"][prob.a["obs_cam"][np.searchsorted(np.sort(prob.a["obs_pt"]), behind)] % 12, :3] * 1.5   # pushed behind some camera
    prob.a["pt_const"][::11] = 1
    P[3] *= 2.5                                                     # homogeneous scale must not matter
    removed_o, status_o = oracle.set_outlier_tracks(prob, 4.0, 3.0)
    status = np.zeros(prob.num_points, np.int32); removed = C.c_int32(0)
    p = prob.struct()
    capi.check(lib.thb_set_outlier_tracks_batch(C.byref(p), 4.0, 3.0, status.ctypes.data_as(C.c_void_p), C.byref(removed), None))
    np.testing.assert_array_equal(status, status_o)
    assert removed.value == removed_o == int((status > 0).sum())
    assert {capi_s for capi_s in set(status.tolist())} >= {-1, 0, 1, 2}
    dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in prob.a.items()}
    pd = prob.struct(); pd.memory_space = capi.THB_MEM_DEVICE
    for k, v in dev.items():
        setattr(pd, k, None if v is None else v.data_ptr())
    d_status = torch.zeros(prob.num_points, dtype=torch.int32, device="cuda")
    capi.check(lib.thb_set_outlier_tracks_batch(C.byref(pd), 4.0, 3.0, C.c_void_p(d_status.data_ptr()), C.byref(removed), None))
    np.testing.assert_array_equal(d_status.cpu().numpy(), status_o)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["default", "coarse_grid_small_quota", "view_subset_preselected", "c2_scaled"])
def test_select_good_tracks_batch_matches_oracle(lib, oracle, case):
    """SelectGoodTracksForBundleAdjustment (select_good_tracks_for_bundle_adjustment.cc:263-325): per-track statistics, the
    best track of every image grid cell, then the per-view quota filled in ascending track order; index work, so the chosen set
    must equal the oracle's exactly. Mixed camera models, a view subset with a caller-provided starting set, host and device memory."""
    import torch
    if case == "c2_scaled":
        prob, gt = synthetic.config_c2(scale=0.05)
        args = (10, 100, 100)
    else:
        prob, gt = synthetic.make_ba_problem(14, 3000, 5, models=(capi.MODEL_PINHOLE, capi.MODEL_EXTENDED_UNIFIED, capi.MODEL_FOV), seed=77)
        args = {"default": (10, 100, 100), "coarse_grid_small_quota": (3, 400, 20), "view_subset_preselected": (4, 150, 60)}[case]
    rng = np.random.default_rng(9)
    prob.a["pts"][:, :3] += rng.normal(0, 0.01, (prob.num_points, 3))      # distinct mean reprojection errors
    cam_sel = None
    start = np.zeros(prob.num_points, np.uint8)
    if case == "view_subset_preselected":
        cam_sel = (np.arange(prob.num_cameras) % 3 != 1).astype(np.uint8)
        start[::17] = 1
    n_o, sel_o = oracle.select_good_tracks(prob, *args, cam_selected=cam_sel, selected=start)
    sel = start.copy(); count = C.c_int32(0)
    p = prob.struct()
    capi.check(lib.thb_select_good_tracks_batch(C.byref(p), None if cam_sel is None else cam_sel.ctypes.data_as(C.c_void_p), *args,
                                                sel.ctypes.data_as(C.c_void_p), C.byref(count), None))
    np.testing.assert_array_equal(sel != 0, sel_o != 0)
    assert count.value == n_o == int((sel != 0).sum())
    assert 0 < n_o < prob.num_points and np.all(sel[start != 0] != 0)
    # every selected view has its quota (or all of its tracks)
    chosen = sel != 0
    for c in range(prob.num_cameras):
        if cam_sel is not None and not cam_sel[c]:
            continue
        pts = prob.a["obs_pt"][prob.a["obs_cam"] == c]
        assert chosen[pts].sum() >= min(args[2], len(pts))
    dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in prob.a.items()}
    pd = prob.struct(); pd.memory_space = capi.THB_MEM_DEVICE
    for k, v in dev.items():
        setattr(pd, k, None if v is None else v.data_ptr())
    d_sel = torch.from_numpy(start.copy()).cuda()
    d_cs = None if cam_sel is None else torch.from_numpy(cam_sel).cuda()
    capi.check(lib.thb_select_good_tracks_batch(C.byref(pd), None if d_cs is None else C.c_void_p(d_cs.data_ptr()), *args,
                                                C.c_void_p(d_sel.data_ptr()), C.byref(count), None))
    np.testing.assert_array_equal(d_sel.cpu().numpy() != 0, sel_o != 0)
    assert lib.thb_select_good_tracks_batch(C.byref(p), None, 10, 0, 100, sel.ctypes.data_as(C.c_void_p), None, None) == capi.THB_E_INVALID_ARGUMENT


def test_track_entries_take_device_memory(lib):
    """thb_ba_tracks_batch / thb_estimate_tracks_batch with every pointer in HBM (the C5 pipeline keeps the scene on the device):
    identical results to the host-buffer call, observations in arbitrary (not track-major) order."""
    import torch
    prob = perturbed(10, 600, 4, seed=77)
    order = np.random.default_rng(5).permutation(prob.num_observations)          # shuffled: the grouping happens on the device
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        if prob.a[k] is not None:
            prob.a[k] = np.ascontiguousarray(prob.a[k][order])
    prob = capi.HostBaProblem(prob.a)
    opts = capi.default_options(lib); opts.use_inner_iterations = 0
    host = prob.copy()
    res_h = gpu_tracks(lib, host, opts)

    def to_device(p):
        dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in p.a.items()}
        pd = p.struct(); pd.memory_space = capi.THB_MEM_DEVICE
        for k, v in dev.items():
            setattr(pd, k, None if v is None else v.data_ptr())
        return dev, pd
    dev, pd = to_device(prob.copy())
    d_res = torch.zeros(prob.num_points * capi.TRACK_BA_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    capi.check(lib.thb_ba_tracks_batch(C.byref(pd), C.byref(opts), C.c_void_p(d_res.data_ptr()), None))
    res_d = d_res.cpu().numpy().view(capi.TRACK_BA_DTYPE)
    assert res_d.tobytes() == res_h.tobytes()
    np.testing.assert_array_equal(dev["pts"].cpu().numpy(), host.a["pts"])
    # EstimateTrack for all tracks, device memory against host memory
    rays = pinhole_rays(prob)
    eopts = capi.ThbTrackEstimatorOptions()
    eopts.min_triangulation_angle_degrees = 2.0; eopts.bundle_adjustment = 1
    ph = prob.copy()
    status_h, res_h2 = gpu_estimate(lib, ph, rays, eopts, opts)
    dev, pd = to_device(prob.copy())
    d_rays = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
    d_status = torch.zeros(prob.num_points, dtype=torch.int32, device="cuda")
    capi.check(lib.thb_estimate_tracks_batch(C.byref(pd), C.c_void_p(d_rays.data_ptr()), C.byref(eopts), C.byref(opts),
                                             C.c_void_p(d_status.data_ptr()), C.c_void_p(d_res.data_ptr()), None))
    np.testing.assert_array_equal(d_status.cpu().numpy(), status_h)
    np.testing.assert_array_equal(dev["pts"].cpu().numpy(), ph.a["pts"])
    assert (status_h == capi.TRACK_ESTIMATED).sum() > 0
