"""Generates the golden vectors under tests/golden/ from the ORACLE (oracle/, the CPU restatement of the reference algorithms).

They are RESTATEMENT-DERIVED: the reference itself cannot be built or imported in this environment (Ceres / Eigen / glog
absent), so these files pin the oracle against drift and give the CUDA path a fixed target that does
not depend on running the oracle at test time; they are not outputs of pyTheia. Inputs are stored explicitly (not only
seeds), so the fixtures stay valid if the synthetic generators change.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle_py  # noqa: E402
from pytheiasfm_b200 import capi, synthetic  # noqa: E402

BA_FIELDS = [name for name, _ in capi.HostBaProblem.FIELDS]


def problem_arrays(prob, prefix):
    return {prefix + k: v for k, v in prob.a.items() if v is not None}


def ba_golden():
    """C1 (BASELINE configs[0]) full solve with the reference defaults, and K1 on a mixed-model problem."""
    prob, _ = synthetic.config_c1()
    out = problem_arrays(prob, "in_")
    work = prob.copy()
    o = oracle_py.default_options()
    s = oracle_py.ba_solve(work, o)
    assert s["rc"] == 0 and s["success"] == 1
    out.update(final_cost=s["final_cost"], initial_cost=s["initial_cost"], num_iterations=s["num_iterations"],
               iter_cost=np.array(s["iter_cost"]), out_cam_ext=work.a["cam_ext"], out_pts=work.a["pts"])
    np.savez_compressed(os.path.join(HERE, "ba_c1_solve.npz"), **out)

    models = (capi.MODEL_PINHOLE, capi.MODEL_FISHEYE, capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION,
              capi.MODEL_DOUBLE_SPHERE, capi.MODEL_EXTENDED_UNIFIED)
    prob, _ = synthetic.make_ba_problem(12, 60, 4, models=models, seed=77, w_scale=True)
    for g in range(prob.num_groups):
        if prob.a["intr_model"][g] == capi.MODEL_DIVISION_UNDISTORTION:
            prob.a["intr"][g, 4] = -5e-7
    r, jc, ji, jp, ok = oracle_py.ba_evaluate(prob)
    out = problem_arrays(prob, "in_")
    out.update(residuals=r, jac_cam=jc, jac_intr=ji, jac_pt=jp, ok=ok)
    np.savez_compressed(os.path.join(HERE, "ba_k1_six_models.npz"), **out)


def ransac_golden():
    """EstimateRelativePose on 6 pairs x 300 correspondences (C4 parameters), five-point on 32 samples."""
    batch, _ = synthetic.make_pair_batch(6, n=300, seed=5, base_seed=4242)
    p = synthetic.c4_params(oracle_py.ransac_default_params())
    rc, res, mask = oracle_py.ransac_relpose_batch(batch, p)
    assert rc == 0
    np.savez_compressed(os.path.join(HERE, "ransac_relpose_6pairs.npz"), corr=batch.corr, pair_offset=batch.pair_offset,
                        seed=batch.seed, results=res.view(np.uint8).reshape(len(res), -1), inlier_mask=mask)
    samples, _ = synthetic.make_pair_batch(32, n=5, inlier_ratio=1.0, noise=1e-3, seed=9)
    x = samples.corr.reshape(32, 5, 4)
    E, n = oracle_py.five_point(x[:, :, :2], x[:, :, 2:])
    np.savez_compressed(os.path.join(HERE, "five_point_32samples.npz"), x1=x[:, :, :2].copy(), x2=x[:, :, 2:].copy(), E=E, num_solutions=n)


def single_track_problem(prob, t):
    """What BundleAdjustTrack(options, t, reconstruction) hands to Ceres (bundle_adjuster.cc:176-221)."""
    a = dict(prob.a)
    keep = prob.a["obs_pt"] == t
    for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
        a[k] = None if prob.a[k] is None else prob.a[k][keep]
    a["cam_const"] = np.full(prob.num_cameras, 3, np.uint8)
    a["intr_const"] = None
    return capi.HostBaProblem(a)


def track_golden():
    """Track stage: TriangulateMidpoint on 80 ragged tracks, BundleAdjustTrack on each of the 80 tracks of a 10-camera scene."""
    rng = np.random.default_rng(21)
    counts = rng.integers(2, 9, 80)
    off = np.zeros(81, np.int64); off[1:] = np.cumsum(counts)
    X = rng.uniform(-2, 2, (80, 3))
    ang = rng.uniform(0, 2 * np.pi, off[-1])
    org = np.stack([6 * np.cos(ang), 6 * np.sin(ang), rng.uniform(-1, 1, off[-1])], 1)
    d = np.repeat(X, counts, axis=0) - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d += rng.normal(0, 1e-3, d.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[off[5]:off[6]] *= 3.0                                  # an indefinite system: LLT fails
    pts, ok = oracle_py.triangulate_midpoint_batch(org, d, off)
    out = dict(tri_origins=org, tri_directions=d, tri_ray_offset=off, tri_points=pts, tri_ok=ok)

    prob, _ = synthetic.make_ba_problem(10, 80, 4, seed=23, pos_sigma=0.0, rot_sigma=0.0)
    prob.a["pts"][:, :3] += rng.normal(0, 0.05, (80, 3))
    prob.a["pt_const"][[4, 40]] = 1
    out.update(problem_arrays(prob, "in_"))
    o = oracle_py.default_options()
    o.use_inner_iterations = 0
    res = np.zeros(80, capi.TRACK_BA_DTYPE)
    refined = prob.a["pts"].copy()
    for t in range(80):
        if prob.a["pt_const"][t]:
            res[t] = (0.0, 0.0, -1, 0)
            continue
        sub = single_track_problem(prob, t)
        s = oracle_py.ba_solve(sub, o)
        assert s["rc"] == 0 and s["success"] == 1
        res[t] = (s["initial_cost"], s["final_cost"], s["num_iterations"], s["termination_type"])
        refined[t] = sub.a["pts"][t]
    out.update(track_results=res.view(np.uint8).reshape(80, -1), out_pts=refined)
    np.savez_compressed(os.path.join(HERE, "track_stage_80tracks.npz"), **out)


if __name__ == "__main__":
    ba_golden()
    ransac_golden()
    track_golden()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
