"""C5-shaped track stage (128 cams / 60k tracks / 360k observations): batched TriangulateMidpoint + batched BundleAdjustTrack
on the GPU (host buffers, copies included) against the oracle looping over the same tracks on the host cores."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pytheiasfm_b200 import capi, synthetic
from oracle import oracle_py

lib = capi.load_library()
prob, gt = synthetic.make_ba_problem(128, 60000, 6, seed=330, pos_sigma=0.0, rot_sigma=0.0)
rng = np.random.default_rng(331)
prob.a["pts"][:, :3] += rng.normal(0, 0.05, (60000, 3))
opts = capi.default_options(lib)
opts.use_inner_iterations = 0
res = np.zeros(prob.num_points, capi.TRACK_BA_DTYPE)
os.environ["THB_TRACK_TIMING"] = "1"
for rep in range(5):
    pg = prob.copy()
    p = pg.struct()
    t0 = time.perf_counter()
    capi.check(lib.thb_ba_tracks_batch(C.byref(p), C.byref(opts), res.ctypes.data_as(C.c_void_p), None))
    dt = time.perf_counter() - t0
    print("thb_ba_tracks_batch: %d tracks in %.2f ms = %.0f tracks/s (mean %.2f iterations)" % (
        prob.num_points, dt * 1e3, prob.num_points / dt, res["num_iterations"].mean()), flush=True)

# oracle: one single-track solve per track, as TrackEstimator calls BundleAdjustTrack (sample of 300 tracks)
order = np.argsort(prob.a["obs_pt"], kind="stable")
starts = np.searchsorted(prob.a["obs_pt"][order], np.arange(prob.num_points + 1))
sample = range(0, 60000, 200)
subs = []
for t in sample:
    a = dict(prob.a)
    sel = order[starts[t]:starts[t + 1]]
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        a[k] = prob.a[k][sel]
    a["cam_const"] = np.full(prob.num_cameras, 3, np.uint8)
    a["intr_const"] = None
    a["pts"] = prob.a["pts"][t:t + 1]          # the single track
    a["pt_const"] = None
    a["obs_pt"] = np.zeros(len(sel), np.int32)
    subs.append(capi.HostBaProblem(a))
t0 = time.perf_counter()
for s in subs:
    oracle_py.ba_solve(s, opts)
dt = time.perf_counter() - t0
print("oracle, one solve per track: %d tracks in %.1f ms = %.0f tracks/s on 1 core (x%d cores if the thread pool scaled perfectly)" % (
    len(subs), dt * 1e3, len(subs) / dt, os.cpu_count()), flush=True)

# triangulation
counts = np.diff(starts)
off = starts.astype(np.int64)
cams = prob.a["obs_cam"][order]
org = prob.a["cam_ext"][cams, :3].copy()
X = gt["pts"][prob.a["obs_pt"][order], :3] if "pts" in gt else prob.a["pts"][prob.a["obs_pt"][order], :3]
d = X - org
d /= np.linalg.norm(d, axis=1, keepdims=True)
out = np.zeros((prob.num_points, 4)); ok = np.zeros(prob.num_points, np.uint8)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
for rep in range(3):
    t0 = time.perf_counter()
    capi.check(lib.thb_triangulate_midpoint_batch(vp(org), vp(d), vp(off), prob.num_points, capi.THB_MEM_HOST, vp(out), vp(ok), None))
    dt = time.perf_counter() - t0
    print("thb_triangulate_midpoint_batch: %d tracks / %d rays in %.2f ms" % (prob.num_points, len(org), dt * 1e3), flush=True)
t0 = time.perf_counter()
oracle_py.triangulate_midpoint_batch(org, d, off)
print("oracle triangulation: %.2f ms on 1 core" % ((time.perf_counter() - t0) * 1e3))
