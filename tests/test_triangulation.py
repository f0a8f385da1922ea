"""theia::TriangulateMidpoint (sfm/triangulation/triangulation.cc:130-157) for batches of tracks: the oracle against numpy
and the reference's own test scene (triangulation_test.cc:312-360) on the CPU, the CUDA path against the oracle on the GPU."""
import ctypes as C

import numpy as np
import pytest

from pytheiasfm_b200 import capi


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def numpy_midpoint(org, dirs):
    A = np.zeros((4, 4)); b = np.zeros(4)
    for o, d in zip(org, dirs):
        dh = np.append(d, 0.0)
        T = np.eye(4) - np.outer(dh, dh)
        A += T
        b += T @ np.append(o, 1.0)
    return np.linalg.solve(A, b)


def ragged_scene(num_tracks, seed, min_rays=2, max_rays=30, noise=1e-3):
    """Cameras on a ring looking at points in a box; unit ray directions with a little angular noise."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(min_rays, max_rays + 1, num_tracks)
    off = np.zeros(num_tracks + 1, np.int64)
    off[1:] = np.cumsum(counts)
    X = rng.uniform(-2, 2, (num_tracks, 3))
    org = np.zeros((off[-1], 3)); dirs = np.zeros((off[-1], 3))
    for t in range(num_tracks):
        n = counts[t]
        ang = rng.uniform(0, 2 * np.pi, n)
        o = np.stack([6 * np.cos(ang), 6 * np.sin(ang), rng.uniform(-1, 1, n)], 1)
        d = X[t] - o
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d += rng.normal(0, noise, d.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        org[off[t]:off[t + 1]] = o
        dirs[off[t]:off[t + 1]] = d
    return org, dirs, off, X


def reference_test_scene(noise, seed=0):
    """triangulation_test.cc:312-360: two points seen from the identity pose and from (R(0.15 rad about y), t)."""
    rng = np.random.default_rng(seed)
    pts = np.array([[5.0, 20.0, 23.0], [-6.0, 16.0, 33.0]])
    a = 0.15
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([-3.0, 1.5, 11.0])
    poses = [(np.eye(3), np.zeros(3)), (R, t)]
    org = []; dirs = []; pix = []
    for X in pts:
        for Rm, tv in poses:
            x = Rm @ X + tv
            p = x[:2] / x[2] + (rng.normal(0, noise, 2) if noise else 0.0)
            pix.append(p)
            org.append(-Rm.T @ tv)
            d = Rm.T @ np.append(p, 1.0)
            dirs.append(d / np.linalg.norm(d))
    return np.array(org), np.array(dirs), np.array([0, 2, 4], np.int64), pts, poses, np.array(pix).reshape(2, 2, 2)


def reprojection_error(pose, Xh, pix):
    R, t = pose
    x = R @ (Xh[:3] / Xh[3]) + t
    return np.sum((x[:2] / x[2] - pix) ** 2)


@pytest.mark.parametrize("noise,tol", [(0.0, 1e-12), (1.0 / 512.0, 1e-5)])
def test_oracle_reference_midpoint_scene(oracle, noise, tol):
    org, dirs, off, pts, poses, pix = reference_test_scene(noise)
    X, ok = oracle.triangulate_midpoint_batch(org, dirs, off)
    assert ok.all()
    for i in range(2):
        for j in range(2):
            assert reprojection_error(poses[j], X[i], pix[i, j]) <= tol


def test_oracle_matches_numpy_on_ragged_tracks(oracle):
    org, dirs, off, Xgt = ragged_scene(300, seed=3)
    X, ok = oracle.triangulate_midpoint_batch(org, dirs, off)
    assert ok.all()
    for t in range(300):
        ref = numpy_midpoint(org[off[t]:off[t + 1]], dirs[off[t]:off[t + 1]])
        np.testing.assert_allclose(X[t], ref, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(X[:, 3], 1.0, rtol=1e-14)
    assert np.median(np.abs(X[:, :3] - Xgt).max(axis=1)) < 0.01   # 2-ray tracks with a narrow baseline are far noisier


def test_oracle_failure_cases(oracle):
    """Fewer than two rays (the reference CHECK-aborts) and an indefinite A (LLT reports NumericalIssue) give ok = 0."""
    org = np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0.0, 0, 0], [1.0, 0, 0]])
    dirs = np.array([[0.0, 0, 1], [0.0, 0, 3], [0.0, 0, 3], [0.0, 0.6, 0.8], [0.6, 0.0, 0.8]])
    off = np.array([0, 1, 3, 3, 5], np.int64)
    X, ok = oracle.triangulate_midpoint_batch(org, dirs, off)
    np.testing.assert_array_equal(ok, [0, 0, 0, 1])
    np.testing.assert_array_equal(X[:3], 0.0)


@pytest.mark.gpu
def test_device_matches_oracle_on_ragged_tracks(lib, oracle):
    org, dirs, off, _ = ragged_scene(5000, seed=5)
    # one indefinite system (|d| = 3) in the middle of the batch
    dirs[off[17]:off[18]] *= 3.0
    X = np.full((5000, 4), np.nan); ok = np.zeros(5000, np.uint8)
    capi.check(lib.thb_triangulate_midpoint_batch(_vp(org), _vp(dirs), _vp(off), 5000, capi.THB_MEM_HOST, _vp(X), _vp(ok), None))
    Xo, oko = oracle.triangulate_midpoint_batch(org, dirs, off)
    np.testing.assert_array_equal(ok, oko)
    assert ok[17] == 0 and ok.sum() == 4999
    # same operations in the same order; FMA contraction on the device is the only difference
    np.testing.assert_allclose(X, Xo, rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
def test_device_edge_cases(lib, oracle):
    org = np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0.0, 0, 0], [1.0, 0, 0]])
    dirs = np.array([[0.0, 0, 1], [0.0, 0, 3], [0.0, 0, 3], [0.0, 0.6, 0.8], [0.6, 0.0, 0.8]])
    off = np.array([0, 1, 3, 3, 5], np.int64)
    X = np.full((4, 4), np.nan); ok = np.full(4, 9, np.uint8)
    capi.check(lib.thb_triangulate_midpoint_batch(_vp(org), _vp(dirs), _vp(off), 4, capi.THB_MEM_HOST, _vp(X), _vp(ok), None))
    Xo, oko = oracle.triangulate_midpoint_batch(org, dirs, off)
    np.testing.assert_array_equal(ok, oko)
    np.testing.assert_allclose(X, Xo, rtol=1e-12, atol=1e-13)
    # an empty batch is a no-op
    capi.check(lib.thb_triangulate_midpoint_batch(None, None, None, 0, capi.THB_MEM_HOST, None, None, None))


@pytest.mark.gpu
def test_device_full_size_properties(lib):
    """C5-sized and larger (1M tracks, ~16M rays): w = 1, the point minimises the summed squared ray distances (gradient of
    sum |(I - d d^T)(X - o)|^2 vanishes), device-resident buffers."""
    import torch
    org, dirs, off, Xgt = ragged_scene(20000, seed=9)
    reps = 50
    counts = np.diff(off)
    big_off = np.zeros(20000 * reps + 1, np.int64)
    big_off[1:] = np.cumsum(np.tile(counts, reps))
    d_org = torch.from_numpy(np.tile(org, (reps, 1))).cuda(); d_dir = torch.from_numpy(np.tile(dirs, (reps, 1))).cuda()
    d_off = torch.from_numpy(big_off).cuda()
    d_X = torch.empty(20000 * reps, 4, dtype=torch.float64, device="cuda"); d_ok = torch.empty(20000 * reps, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    capi.check(lib.thb_triangulate_midpoint_batch(d_org.data_ptr(), d_dir.data_ptr(), d_off.data_ptr(), 20000 * reps, capi.THB_MEM_DEVICE,
                                                  d_X.data_ptr(), d_ok.data_ptr(), st))
    torch.cuda.synchronize()
    X = d_X.cpu().numpy(); ok = d_ok.cpu().numpy()
    assert ok.all()
    assert np.array_equal(X[:20000], X[-20000:])
    np.testing.assert_allclose(X[:, 3], 1.0, rtol=1e-13)
    assert np.median(np.abs(X[:20000, :3] - Xgt).max(axis=1)) < 0.01
    for t in range(0, 20000, 997):
        g = np.zeros(3)
        for o, d in zip(org[off[t]:off[t + 1]], dirs[off[t]:off[t + 1]]):
            g += (np.eye(3) - np.outer(d, d)) @ (X[t, :3] - o)
        assert np.abs(g).max() < 1e-10
