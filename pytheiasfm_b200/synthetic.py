"""Deterministic synthetic reconstructions / pair batches for BASELINE.json's configs.

Pure numpy (seeded default_rng); shapes follow SURVEY.md §8(d). The arrays are already
in the C-ABI layout of include/theia_b200.h.
"""
import numpy as np

from . import capi


def _rotvec_from_matrix(R):
    """Batched rotation matrix -> angle-axis (world->camera, Camera::ORIENTATION)."""
    R = np.asarray(R, dtype=np.float64)
    tr = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(tr)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    s = 2.0 * np.sin(th)
    small = np.abs(s) < 1e-12
    k = np.where(small, 0.5, th / np.where(small, 1.0, s))
    return v * k[..., None]


def rotmat_from_rotvec(w):
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1)
    out = np.zeros(w.shape[:-1] + (3, 3))
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    th2 = th * th
    a = np.where(th < 1e-8, 1.0 - th2 / 6.0, np.sin(th) / np.where(th < 1e-8, 1.0, th))
    b = np.where(th < 1e-8, 0.5 - th2 / 24.0, (1.0 - np.cos(th)) / np.where(th < 1e-8, 1.0, th2))
    out = np.eye(3) + a[..., None, None] * K + b[..., None, None] * (K @ K)
    return out


def _look_at(C, target, up=np.array([0.0, 0.0, 1.0])):
    z = target - C
    z /= np.linalg.norm(z, axis=-1, keepdims=True)
    x = np.cross(z, up)
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    y = np.cross(z, x)
    return np.stack([x, y, z], axis=-2)  # rows = camera axes in world coords


def project(model, K, p):
    """numpy restatement used ONLY to synthesise observations (double precision)."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    if model == capi.MODEL_PINHOLE:
        nx, ny = x / z, y / z
        r2 = nx * nx + ny * ny
        d = 1.0 + r2 * (K[5] + K[6] * r2)
        dx, dy = nx * d, ny * d
        return np.stack([K[0] * dx + K[2] * dy + K[3], K[0] * K[1] * dy + K[4]], -1)
    if model == capi.MODEL_DOUBLE_SPHERE:
        xi, al = K[5], K[6]
        d1 = np.sqrt(x * x + y * y + z * z)
        k = xi * d1 + z
        d2 = np.sqrt(x * x + y * y + k * k)
        n = al * d2 + (1 - al) * k
        dx, dy = x / n, y / n
        return np.stack([K[0] * dx + K[2] * dy + K[3], K[0] * K[1] * dy + K[4]], -1)
    if model == capi.MODEL_EXTENDED_UNIFIED:
        al, be = K[5], K[6]
        rho = np.sqrt(be * (x * x + y * y) + z * z)
        n = al * rho + (1 - al) * z
        dx, dy = x / n, y / n
        return np.stack([K[0] * dx + K[2] * dy + K[3], K[0] * K[1] * dy + K[4]], -1)
    if model == capi.MODEL_FISHEYE:
        r = np.sqrt(x * x + y * y)
        th = np.arctan2(r, np.abs(z))
        t2 = th * th
        td = th * (1 + K[5] * t2 + K[6] * t2 ** 2 + K[7] * t2 ** 3 + K[8] * t2 ** 4)
        s = np.where(r > 1e-4, td / np.maximum(r, 1e-300), 1.0)
        dx, dy = x * s, y * s
        return np.stack([K[0] * dx + K[2] * dy + K[3], K[0] * K[1] * dy + K[4]], -1)
    if model == capi.MODEL_FOV:
        nx, ny = x / z, y / z
        r = np.sqrt(nx * nx + ny * ny)
        om = K[4]
        rd = np.arctan(2 * r * np.tan(om / 2)) / np.maximum(r * om, 1e-300)
        return np.stack([K[0] * rd * nx + K[2], K[0] * K[1] * rd * ny + K[3]], -1)
    if model == capi.MODEL_DIVISION_UNDISTORTION:
        ux, uy = K[0] * x / z, K[0] * K[1] * y / z
        r2 = ux * ux + uy * uy
        k = K[4]
        den = 2 * k * r2
        sc = np.where(np.abs(den) < 1e-15, 1.0, (1 - np.sqrt(np.maximum(1 - 4 * k * r2, 0))) / np.where(np.abs(den) < 1e-15, 1.0, den))
        return np.stack([ux * sc + K[2], uy * sc + K[3]], -1)
    raise ValueError(model)


def default_intrinsics(model, f=1000.0, cx=500.0, cy=500.0):
    K = np.zeros(capi.THB_INTR_STRIDE)
    if model in (capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION):
        K[:4] = [f, 1.0, cx, cy]
        K[4] = 0.75 if model == capi.MODEL_FOV else -5e-7
    else:
        K[:5] = [f, 1.0, 0.0, cx, cy]
        if model == capi.MODEL_PINHOLE:
            K[5:7] = [0.0, 0.0]
        elif model == capi.MODEL_DOUBLE_SPHERE:
            K[5:7] = [-0.27, 0.57]
        elif model == capi.MODEL_EXTENDED_UNIFIED:
            K[5:7] = [0.6, 1.1]
        elif model == capi.MODEL_FISHEYE:
            K[5:9] = [-0.02, 0.003, 0.0, 0.0]
    return K


def make_ba_problem(num_cameras=10, num_points=500, obs_per_point=4, models=(capi.MODEL_PINHOLE,),
                    seed=0, pixel_sigma=0.5, pos_sigma=0.02, rot_sigma=0.01, pt_sigma=0.02,
                    num_rings=1, ring_radius=6.0, box=(2.0, 2.0, 2.0), focal=1000.0, image=1000.0,
                    intr_const_mask=None, w_scale=False):
    """Cameras on rings looking at the origin, points uniform in a box, `obs_per_point` random
    in-frame observers per point, Gaussian pixel noise, perturbed initial parameters.
    One shared intrinsics group per entry of `models` (cameras are assigned round-robin by halves).
    Returns (HostBaProblem, ground_truth dict)."""
    rng = np.random.default_rng(seed)
    nc, npnt = num_cameras, num_points
    per_ring = (nc + num_rings - 1) // num_rings
    idx = np.arange(nc)
    ring = idx // per_ring
    ang = 2 * np.pi * (idx % per_ring) / per_ring + 0.37 * ring
    rad = ring_radius * (1.0 + 0.08 * ring)
    height = (ring - (num_rings - 1) / 2.0) * (box[2] * 0.35) + 0.0
    Cw = np.stack([rad * np.cos(ang), rad * np.sin(ang), height + 0.5 * box[2]], -1)
    target = rng.normal(0.0, 0.05 * min(box[0], box[1]), size=(nc, 3))
    R = _look_at(Cw, target)
    aa = _rotvec_from_matrix(R)
    pts = (rng.random((npnt, 3)) * 2.0 - 1.0) * np.array(box)

    ng = len(models)
    cam_group = (idx * ng // nc).astype(np.int32)
    intr = np.stack([default_intrinsics(m, focal if m not in (capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED) else 0.4 * focal,
                                         image / 2, image / 2) for m in models])
    intr_model = np.array(models, dtype=np.int32)

    # choose observers: random candidates, keep the first `obs_per_point` that are in frame
    k = obs_per_point
    obs_cam = np.empty((npnt, k), dtype=np.int32)
    filled = np.zeros(npnt, dtype=np.int32)
    for _ in range(64):
        todo = np.nonzero(filled < k)[0]
        if todo.size == 0:
            break
        cand = rng.integers(0, nc, size=todo.size).astype(np.int32)
        pc = np.einsum("nij,nj->ni", R[cand], pts[todo] - Cw[cand])
        ok = pc[:, 2] > 0.5
        pix = np.zeros((todo.size, 2))
        for g, m in enumerate(models):
            sel = ok & (cam_group[cand] == g)
            if sel.any():
                pix[sel] = project(m, intr[g], pc[sel])
        ok &= np.all((pix > 0.02 * image) & (pix < 0.98 * image), axis=1)
        # reject duplicates
        dup = (obs_cam[todo] == cand[:, None]) & (np.arange(k)[None, :] < filled[todo][:, None])
        ok &= ~dup.any(axis=1)
        t = todo[ok]
        obs_cam[t, filled[t]] = cand[ok]
        filled[t] += 1
    if (filled < k).any():
        raise RuntimeError("could not find %d in-frame observers for every point" % k)
    obs_pt = np.repeat(np.arange(npnt, dtype=np.int32), k)
    obs_cam = obs_cam.reshape(-1)
    pc = np.einsum("nij,nj->ni", R[obs_cam], pts[obs_pt] - Cw[obs_cam])
    xy = np.zeros((obs_cam.size, 2))
    for g, m in enumerate(models):
        sel = cam_group[obs_cam] == g
        xy[sel] = project(m, intr[g], pc[sel])
    xy += rng.normal(0.0, pixel_sigma, size=xy.shape) if pixel_sigma > 0 else 0.0
    # shuffle observation order (the reference iterates hash maps: no particular order)
    perm = rng.permutation(obs_cam.size)
    obs_cam, obs_pt, xy = obs_cam[perm], obs_pt[perm], xy[perm]

    gt = dict(cam_ext=np.concatenate([Cw, aa], 1), pts=np.concatenate([pts, np.ones((npnt, 1))], 1), intr=intr.copy())
    cam0 = gt["cam_ext"].copy()
    cam0[:, :3] += rng.normal(0, pos_sigma, (nc, 3)) if pos_sigma > 0 else 0.0
    cam0[:, 3:] += rng.normal(0, rot_sigma, (nc, 3)) if rot_sigma > 0 else 0.0
    pts0 = gt["pts"].copy()
    pts0[:, :3] += rng.normal(0, pt_sigma, (npnt, 3)) if pt_sigma > 0 else 0.0
    if w_scale:  # exercise the free homogeneous scale
        s = rng.uniform(0.5, 2.0, (npnt, 1))
        pts0 *= s
    if intr_const_mask is None:
        intr_const = np.array([(1 << capi.MODEL_NUM_PARAMS[m]) - 1 for m in models], dtype=np.uint16)
    else:
        intr_const = np.array(intr_const_mask, dtype=np.uint16)
    prob = capi.HostBaProblem(dict(
        cam_ext=cam0, cam_const=np.zeros(nc, np.uint8), cam_group=cam_group, intr=intr, intr_model=intr_model,
        intr_const=intr_const, pts=pts0, pt_const=np.zeros(npnt, np.uint8),
        obs_cam=obs_cam, obs_pt=obs_pt, obs_xy=xy, obs_sqrt_info=np.ones_like(xy)))
    return prob, gt


def config_c1(seed=1):
    """BASELINE.json configs[0]: Pinhole, 10 cams / 500 pts / 2k obs."""
    return make_ba_problem(10, 500, 4, seed=seed, num_rings=1, ring_radius=6.0, box=(2.0, 2.0, 2.0))


def config_c2(seed=2, scale=1.0):
    """BASELINE.json configs[1]: Pinhole, 1k cams / 100k pts / 1M obs (scale<1 shrinks all three)."""
    nc = max(10, int(round(1000 * scale)))
    npnt = max(50, int(round(100000 * scale)))
    return make_ba_problem(nc, npnt, 10, seed=seed, num_rings=10 if nc >= 100 else 2, ring_radius=24.0,
                           box=(10.0, 10.0, 3.0))


def config_c3(seed=3, scale=1.0):
    """BASELINE.json configs[2]: DoubleSphere + ExtendedUnified, 500 cams / 50k pts / 8 obs per point,
    intrinsics_to_optimize = FOCAL_LENGTH | RADIAL_DISTORTION (constant: a, s, cx, cy)."""
    nc = max(10, int(round(500 * scale)))
    npnt = max(50, int(round(50000 * scale)))
    const = 0b0011110  # aspect, skew, cx, cy constant; f, and the two distortion slots free
    return make_ba_problem(nc, npnt, 8, models=(capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED), seed=seed,
                           num_rings=5 if nc >= 50 else 2, ring_radius=24.0, box=(10.0, 10.0, 3.0),
                           intr_const_mask=[const, const])


def random_rotation(rng, max_deg):
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    return rotmat_from_rotvec(axis * np.deg2rad(rng.uniform(0, max_deg)))


def make_pair(rng, n=2000, inlier_ratio=0.6, noise=1e-3, max_rot_deg=15.0):
    """One image pair in normalised coordinates: inliers from random 3-D points (depth 4-10)
    under a random rotation and a unit baseline, Gaussian noise; outliers uniform in [-1,1]^2.
    Returns (corr [n,4], R, position, inlier flags)."""
    R = random_rotation(rng, max_rot_deg)
    c = rng.normal(size=3); c /= np.linalg.norm(c)          # camera-2 position (unit baseline)
    ni = int(round(inlier_ratio * n))
    X = np.stack([rng.uniform(-3, 3, ni), rng.uniform(-3, 3, ni), rng.uniform(4, 10, ni)], -1)
    x1 = X[:, :2] / X[:, 2:3]
    Xc = (X - c) @ R.T
    x2 = Xc[:, :2] / Xc[:, 2:3]
    x1 = x1 + rng.normal(0, noise, x1.shape); x2 = x2 + rng.normal(0, noise, x2.shape)
    out = rng.uniform(-1, 1, (n - ni, 4))
    corr = np.concatenate([np.concatenate([x1, x2], 1), out], 0)
    flags = np.concatenate([np.ones(ni, bool), np.zeros(n - ni, bool)])
    perm = rng.permutation(n)
    return corr[perm], R, c, flags[perm]


def make_pair_batch(num_pairs, n=2000, inlier_ratio=0.6, noise=1e-3, seed=0, base_seed=1000):
    """BASELINE configs[3]-shaped batch: `num_pairs` pairs x `n` correspondences, per-pair seed = base + index."""
    rng = np.random.default_rng(seed)
    corrs, gts = [], []
    for _ in range(num_pairs):
        c, R, p, f = make_pair(rng, n, inlier_ratio, noise)
        corrs.append(c); gts.append((R, p, f))
    return capi.HostPairBatch(corrs, base_seed + np.arange(num_pairs)), gts


def make_pair_batch_indexed(indices, n=2000, inlier_ratio=0.6, noise=1e-3, seed=0, base_seed=1000):
    """The pairs `indices` of a BASELINE configs[3]-shaped table in which every pair has its OWN generator
    (default_rng([seed, index])) and seed base_seed + index: a rank builds exactly the pairs it owns, and any
    partition of the table yields identical pairs."""
    corrs, gts = [], []
    for i in indices:
        c, R, p, f = make_pair(np.random.default_rng([seed, int(i)]), n, inlier_ratio, noise)
        corrs.append(c); gts.append((R, p, f))
    return capi.HostPairBatch(corrs, base_seed + np.asarray(indices, dtype=np.int64)), gts


def c4_params(lib_or_oracle_params):
    """RansacParameters of BASELINE configs[3]."""
    p = lib_or_oracle_params
    p.error_thresh = (2e-3) ** 2; p.failure_probability = 1e-4; p.min_iterations = 10; p.max_iterations = 1000
    p.use_mle = 1; p.use_lo = 0; p.min_inlier_ratio = 0.0; p.ransac_type = 0
    return p


def make_abspose_batch(num, n=500, inlier_ratio=0.7, noise=1e-3, seed=0, base_seed=2000):
    """2D-3D batches for EstimateCalibratedAbsolutePose: datum = (x, y, X, Y, Z) in normalised image coordinates."""
    rng = np.random.default_rng(seed)
    data, gts = [], []
    for _ in range(num):
        R = random_rotation(rng, 40.0)
        c = rng.normal(size=3)
        ni = int(round(inlier_ratio * n))
        Xc = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(3, 9, n)], -1)   # in the camera frame
        X = Xc @ R + c                                                                            # world = R^T Xc + c
        x = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, noise, (n, 2))
        x[ni:] = rng.uniform(-1, 1, (n - ni, 2))
        flags = np.arange(n) < ni
        perm = rng.permutation(n)
        data.append(np.concatenate([x, X], 1)[perm]); gts.append((R, c, flags[perm]))
    return capi.HostPairBatch(data, base_seed + np.arange(num), width=5), gts


def make_homography_batch(num, n=500, inlier_ratio=0.6, noise=1e-3, seed=0, base_seed=3000):
    rng = np.random.default_rng(seed)
    data, gts = [], []
    for _ in range(num):
        H = np.eye(3) + rng.normal(0, 0.2, (3, 3)); H /= H[2, 2]
        ni = int(round(inlier_ratio * n))
        x1 = rng.uniform(-1, 1, (n, 2))
        y = np.c_[x1, np.ones(n)] @ H.T
        x2 = y[:, :2] / y[:, 2:] + rng.normal(0, noise, (n, 2))
        x2[ni:] = rng.uniform(-1, 1, (n - ni, 2))
        flags = np.arange(n) < ni
        perm = rng.permutation(n)
        data.append(np.c_[x1, x2][perm]); gts.append((H, flags[perm]))
    return capi.HostPairBatch(data, base_seed + np.arange(num)), gts


def make_two_view_batch(num_pairs, n=600, models=(capi.MODEL_PINHOLE,), inlier_ratio=0.7, noise_px=0.5, seed=0, base_seed=5000, image=1000,
                        planar_fraction=0.0):
    """Image pairs with PIXEL correspondences for EstimateTwoViewInfo / VerifyMatches: camera 1 at the origin, camera 2 a few
    degrees and a unit baseline away, points in front of both, each view projected through its own camera model
    (models[p % len] for view 1, models[(p + 1) % len] for view 2), Gaussian pixel noise, uniform outliers; a fraction of the
    inliers may lie on a plane (homography inliers). Returns (HostPairBatch, intr1, intr2, ground truth list)."""
    rng = np.random.default_rng(seed)
    corrs, gts = [], []
    intr1 = np.zeros(num_pairs, capi.VIEW_INTRINSICS_DTYPE); intr2 = np.zeros(num_pairs, capi.VIEW_INTRINSICS_DTYPE)
    for p in range(num_pairs):
        m1, m2 = models[p % len(models)], models[(p + 1) % len(models)]
        K1 = default_intrinsics(m1, 0.9 * image if m1 not in (capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED) else 0.4 * image, image / 2, image / 2)
        K2 = default_intrinsics(m2, 0.95 * image if m2 not in (capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED) else 0.42 * image, image / 2, image / 2)
        for I, m, K in ((intr1, m1, K1), (intr2, m2, K2)):
            I[p]["model"] = m; I[p]["image_width"] = image; I[p]["image_height"] = image; I[p]["focal_length_is_set"] = 1; I[p]["params"] = K
        R = random_rotation(rng, 12.0)
        c = rng.normal(size=3); c[2] *= 0.3; c /= np.linalg.norm(c)
        ni = int(round(inlier_ratio * n))
        X = np.stack([rng.uniform(-2.5, 2.5, ni), rng.uniform(-2.5, 2.5, ni), rng.uniform(5, 11, ni)], -1)
        npl = int(planar_fraction * ni)
        if npl:
            X[:npl, 2] = 8.0 + 0.1 * X[:npl, 0]
        x1 = project(m1, K1, X)
        x2 = project(m2, K2, (X - c) @ R.T)
        x1 = x1 + rng.normal(0, noise_px, x1.shape); x2 = x2 + rng.normal(0, noise_px, x2.shape)
        out = rng.uniform(0.1 * image, 0.9 * image, (n - ni, 4))
        corr = np.concatenate([np.concatenate([x1, x2], 1), out], 0)
        flags = np.concatenate([np.ones(ni, bool), np.zeros(n - ni, bool)])
        perm = rng.permutation(n)
        corrs.append(corr[perm]); gts.append((R, c, flags[perm]))
    return capi.HostPairBatch(corrs, base_seed + np.arange(num_pairs)), intr1, intr2, gts


def config_c5(seed=5, num_cameras=128, num_points=60000, window=6, pair_reach=4, outlier_ratio=0.15, pixel_sigma=0.5, image=1000.0, focal=1000.0):
    """BASELINE configs[4]: south-building-shaped scene - `num_cameras` Pinhole cameras on two rings around a box of
    `num_points` points, every point seen by `window` neighbouring cameras of one ring (~6 observations per track), image
    pairs = cameras within `pair_reach` ring neighbours (4 * num_cameras pairs of ~1-2k matches), `outlier_ratio` wrong matches
    per pair. Returns a dict: the flattened reconstruction (HostBaProblem with perturbed cameras / ground-truth-free points),
    the pair table in PIXELS (HostPairBatch + per-pair view indices + ThbViewIntrinsics arrays) and the ground truth."""
    rng = np.random.default_rng(seed)
    per_ring = num_cameras // 2
    idx = np.arange(num_cameras)
    ring = idx // per_ring
    ang = 2 * np.pi * (idx % per_ring) / per_ring + 0.05 * ring
    Cw = np.stack([12.0 * np.cos(ang), 12.0 * np.sin(ang), np.where(ring == 0, 1.0, 3.5)], -1)
    R = _look_at(Cw, rng.normal(0.0, 0.1, size=(num_cameras, 3)) + np.array([0.0, 0.0, 1.5]))
    aa = _rotvec_from_matrix(R)
    K = default_intrinsics(capi.MODEL_PINHOLE, focal, image / 2, image / 2)
    # points on / near the faces of a box, so that neighbouring cameras see them at similar scale
    pts = (rng.random((num_points, 3)) * 2.0 - 1.0) * np.array([3.0, 3.0, 2.0]) + np.array([0.0, 0.0, 1.5])
    pring = rng.integers(0, 2, num_points)
    centre = (np.round((np.arctan2(pts[:, 1], pts[:, 0]) % (2 * np.pi)) / (2 * np.pi) * per_ring).astype(int)) % per_ring
    offs = np.arange(window) - window // 2
    obs_cam = (pring[:, None] * per_ring + (centre[:, None] + offs[None, :]) % per_ring).astype(np.int32).reshape(-1)
    obs_pt = np.repeat(np.arange(num_points, dtype=np.int32), window)
    pc = np.einsum("nij,nj->ni", R[obs_cam], pts[obs_pt] - Cw[obs_cam])
    xy_true = project(capi.MODEL_PINHOLE, K, pc)
    ok = (pc[:, 2] > 0.5) & np.all((xy_true > 0.02 * image) & (xy_true < 0.98 * image), axis=1)
    obs_cam, obs_pt, xy_true = obs_cam[ok], obs_pt[ok], xy_true[ok]
    xy = xy_true + rng.normal(0.0, pixel_sigma, xy_true.shape)
    # pair table: shared tracks of cameras within pair_reach neighbours on the same ring
    order = np.lexsort((obs_pt, obs_cam))
    oc, op, oxy = obs_cam[order], obs_pt[order], xy[order]
    start = np.searchsorted(oc, np.arange(num_cameras + 1))
    corrs, pair_views = [], []
    for a in range(num_cameras):
        for d in range(1, pair_reach + 1):
            b = (a // per_ring) * per_ring + ((a % per_ring) + d) % per_ring
            pa, pb = op[start[a]:start[a + 1]], op[start[b]:start[b + 1]]
            common, ia, ib = np.intersect1d(pa, pb, return_indices=True)
            m = np.concatenate([oxy[start[a]:start[a + 1]][ia], oxy[start[b]:start[b + 1]][ib]], 1)
            nout = int(outlier_ratio * len(m))
            if nout and len(pa) and len(pb):
                wrong = np.concatenate([oxy[start[a]:start[a + 1]][rng.integers(0, len(pa), nout)], oxy[start[b]:start[b + 1]][rng.integers(0, len(pb), nout)]], 1)
                m = np.concatenate([m, wrong], 0)
            corrs.append(m[rng.permutation(len(m))]); pair_views.append((a, b))
    pairs = capi.HostPairBatch(corrs, 9000 + np.arange(len(corrs)))
    intr = np.zeros(len(corrs), capi.VIEW_INTRINSICS_DTYPE)
    intr["model"] = capi.MODEL_PINHOLE; intr["image_width"] = int(image); intr["image_height"] = int(image); intr["focal_length_is_set"] = 1
    intr["params"] = K
    cam0 = np.concatenate([Cw, aa], 1)
    cam0[:, :3] += rng.normal(0, 0.01, (num_cameras, 3)); cam0[:, 3:] += rng.normal(0, 0.001, (num_cameras, 3))
    prob = capi.HostBaProblem(dict(
        cam_ext=cam0, cam_const=np.zeros(num_cameras, np.uint8), cam_group=np.zeros(num_cameras, np.int32), intr=K[None].copy(),
        intr_model=np.array([capi.MODEL_PINHOLE], np.int32), intr_const=np.array([(1 << 7) - 1], np.uint16),
        pts=np.concatenate([np.zeros((num_points, 3)), np.ones((num_points, 1))], 1), pt_const=np.zeros(num_points, np.uint8),
        obs_cam=obs_cam, obs_pt=obs_pt, obs_xy=xy, obs_sqrt_info=np.ones_like(xy)))
    return dict(problem=prob, pairs=pairs, pair_views=np.array(pair_views, np.int32), intrinsics=intr,
                gt=dict(cam_ext=np.concatenate([Cw, aa], 1), pts=pts, K=K))


def pinhole_rays(prob):
    """Camera::PixelToUnitDepthRay(feature).normalized() for every observation of an undistorted-Pinhole problem (what the
    caller of thb_estimate_tracks_batch provides)."""
    a = prob.a
    K = a["intr"][a["cam_group"][a["obs_cam"]]]
    y = (a["obs_xy"][:, 1] - K[:, 4]) / (K[:, 0] * K[:, 1])
    x = (a["obs_xy"][:, 0] - K[:, 3] - y * K[:, 2]) / K[:, 0]
    u = np.stack([x, y, np.ones_like(x)], -1)
    Rm = rotmat_from_rotvec(a["cam_ext"][a["obs_cam"], 3:])
    d = np.einsum("nji,nj->ni", Rm, u)
    return np.ascontiguousarray(d / np.linalg.norm(d, axis=1, keepdims=True))
