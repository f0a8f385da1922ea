#!/usr/bin/env python
"""Pin the oracle (and, on a GPU box, the CUDA path) to the REAL reference when it is importable.

  python tools/compare_pytheia.py [--gpu] [--out tests/golden/pytheia_c1.json]

`import pytheia` fails in the build container and on the GPU boxes of this project (Ceres / Eigen / glog are absent and
there is no wheel offline - SURVEY F2), so this script prints "pytheia not importable" and exits 0 there. On a machine
with pyTheia installed it
  1. builds BASELINE configs[0] (C1: 10 cams / 500 pts / 2k obs, the arrays of pytheiasfm_b200.synthetic.config_c1) as a
     pt.sfm.Reconstruction through the reference's own Python API (the calls pytests/sfm/random_recon_gen.py makes),
  2. runs pt.sfm.BundleAdjustReconstruction with use_inner_iterations False and True, the reference's other defaults
     (bundle_adjustment.cc:188-217), and compares summary.final_cost and the refined parameters with the oracle
     (oracle/ba_oracle.cc) and, with --gpu, with thb_ba_solve: <= 1e-6 relative on the final cost (north_star),
  3. runs pt.sfm.EstimateRelativePose on C4-shaped pairs; Python cannot fix the reference's RANSAC seed (`rng` is not
     bound, src/pytheia/solvers/solvers.cc:89-102), so only the pose (<= 1 degree) and the inlier count (within 2 %)
     are compared, not the inlier sets,
  4. writes everything it measured to --out so the numbers can be committed as golden vectors WITH PROVENANCE
     (pytheia version, Ceres version string if exposed, platform).
"""
import argparse
import json
import os
import platform
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_reconstruction(pt, prob):
    """HostBaProblem -> pt.sfm.Reconstruction (one intrinsics group per problem group, every view / track estimated)."""
    from pytheiasfm_b200 import capi
    names = {capi.MODEL_PINHOLE: "PINHOLE", capi.MODEL_FISHEYE: "FISHEYE", capi.MODEL_FOV: "FOV",
             capi.MODEL_DIVISION_UNDISTORTION: "DIVISION_UNDISTORTION", capi.MODEL_DOUBLE_SPHERE: "DOUBLE_SPHERE",
             capi.MODEL_EXTENDED_UNIFIED: "EXTENDED_UNIFIED"}
    recon = pt.sfm.Reconstruction()
    a = prob.a
    view_ids, track_ids = [], []
    for c in range(prob.num_cameras):
        g = int(a["cam_group"][c]); K = a["intr"][g]; model = int(a["intr_model"][g])
        prior = pt.sfm.CameraIntrinsicsPrior()
        prior.camera_intrinsics_model_type = names[model]
        prior.focal_length.value = [float(K[0])]
        prior.aspect_ratio.value = [float(K[1])]
        skewed = model not in (capi.MODEL_FOV, capi.MODEL_DIVISION_UNDISTORTION)
        if skewed:
            prior.skew.value = [float(K[2])]
            prior.principal_point.value = [float(K[3]), float(K[4])]
            dist = K[5:capi.MODEL_NUM_PARAMS[model]]
            if model == capi.MODEL_DOUBLE_SPHERE:  # prior order is [alpha, xi], storage [xi, alpha] (double_sphere_camera_model.cc:103-106)
                dist = dist[::-1]
        else:
            prior.principal_point.value = [float(K[2]), float(K[3])]
            dist = K[4:5]
        prior.radial_distortion.value = [float(x) for x in dist] + [0.0] * (4 - len(dist))
        prior.image_width = 1000; prior.image_height = 1000
        vid = recon.AddView(str(c), g, float(c))
        view = recon.View(vid)
        cam = view.MutableCamera()
        cam.SetFromCameraIntrinsicsPriors(prior)
        cam.SetPosition(a["cam_ext"][c, :3].copy())
        cam.SetOrientationFromAngleAxis(a["cam_ext"][c, 3:].copy())
        view.SetIsEstimated(True)
        view_ids.append(vid)
    for p in range(prob.num_points):
        tid = recon.AddTrack()
        tr = recon.MutableTrack(tid)
        tr.SetPoint(a["pts"][p].tolist())
        tr.SetIsEstimated(True)
        track_ids.append(tid)
    for i in range(prob.num_observations):
        recon.AddObservation(view_ids[int(a["obs_cam"][i])], track_ids[int(a["obs_pt"][i])], pt.sfm.Feature(a["obs_xy"][i].copy()))
    return recon, view_ids, track_ids


def read_back(recon, view_ids, track_ids):
    cams = np.array([np.concatenate([recon.View(v).Camera().GetPosition(), recon.View(v).Camera().GetOrientationAsAngleAxis()]) for v in view_ids])
    pts = np.array([recon.Track(t).Point() for t in track_ids])
    return cams, pts / pts[:, 3:4]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true", help="also run the CUDA path (needs a B200)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "pytheia_compare.json"))
    args = ap.parse_args()
    try:
        import pytheia as pt
    except Exception as e:  # noqa: BLE001
        print("pytheia not importable (%s): nothing compared; the oracle stays PARITY UNPINNED (DESIGN.md section 5)" % type(e).__name__)
        return 0
    from oracle import oracle_py
    from pytheiasfm_b200 import synthetic
    report = {"provenance": {"pytheia": getattr(pt, "__version__", "unknown"), "python": sys.version.split()[0], "platform": platform.platform()}}
    worst = 0.0
    for inner in (False, True):
        prob, _ = synthetic.config_c1()
        recon, vids, tids = build_reconstruction(pt, prob)
        opts = pt.sfm.BundleAdjustmentOptions()
        opts.use_inner_iterations = inner
        summ = pt.sfm.BundleAdjustReconstruction(opts, recon)
        cams, pts = read_back(recon, vids, tids)
        entry = {"pytheia_final_cost": summ.final_cost, "pytheia_initial_cost": summ.initial_cost, "success": bool(summ.success)}
        oo = oracle_py.default_options(); oo.use_inner_iterations = int(inner)
        po = prob.copy()
        o = oracle_py.ba_solve(po, oo)
        if o["rc"] == 0:
            entry["oracle_final_cost"] = o["final_cost"]; entry["oracle_iterations"] = o["num_iterations"]
            entry["rel_final_cost_oracle"] = abs(o["final_cost"] - summ.final_cost) / summ.final_cost
            entry["max_abs_camera_diff_oracle"] = float(np.abs(po.a["cam_ext"] - cams).max())
            entry["max_abs_point_diff_oracle"] = float(np.abs(po.a["pts"][:, :3] / po.a["pts"][:, 3:4] - pts[:, :3]).max())
            worst = max(worst, entry["rel_final_cost_oracle"])
        else:
            entry["oracle"] = "rc %d" % o["rc"]
        if args.gpu:
            import ctypes as C
            from pytheiasfm_b200 import capi
            lib = capi.load_library()
            go = capi.default_options(lib); go.use_inner_iterations = int(inner)
            pg = prob.copy(); s = capi.ThbBaSummary(); p = pg.struct()
            capi.check(lib.thb_ba_solve(C.byref(p), C.byref(go), C.byref(s), None))
            entry["gpu_final_cost"] = s.final_cost
            entry["rel_final_cost_gpu"] = abs(s.final_cost - summ.final_cost) / summ.final_cost
            worst = max(worst, entry["rel_final_cost_gpu"])
        report["c1_inner_iterations_%s" % inner] = entry
        print("C1 inner=%s: %s" % (inner, json.dumps(entry)))
    # C4-shaped RANSAC: pose and inlier count only (no seed control from Python)
    batch, gts = synthetic.make_pair_batch_indexed(range(8), n=2000, seed=21)
    params = pt.solvers.RansacParameters()
    params.error_thresh = (2e-3) ** 2; params.failure_probability = 1e-4; params.min_iterations = 10; params.max_iterations = 1000
    params.use_mle = True
    orc, ores, omask = oracle_py.ransac_relpose_batch(batch, synthetic.c4_params(oracle_py.ransac_default_params()))
    rows = []
    for i in range(batch.num_pairs):
        c = batch.corr[batch.pair_offset[i]: batch.pair_offset[i + 1]]
        corrs = []
        for x in c:
            fc = pt.matching.FeatureCorrespondence(pt.sfm.Feature(x[:2].copy()), pt.sfm.Feature(x[2:].copy()))
            corrs.append(fc)
        ok, pose, summ = pt.sfm.EstimateRelativePose(params, pt.sfm.RansacType(0), corrs)
        R = np.asarray(pose.rotation)
        ang = float(np.rad2deg(np.arccos(np.clip((np.trace(R @ ores["rotation"][i].T) - 1) / 2, -1, 1))))
        rows.append({"pair": i, "pytheia_inliers": len(summ.inliers), "oracle_inliers": int(ores["num_inliers"][i]), "rotation_diff_deg": ang})
    report["c4_pairs"] = rows
    print("C4:", json.dumps(rows))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)
    print("wrote", args.out, "worst relative final-cost difference %.3g (bar 1e-6)" % worst)
    return 0 if worst <= 1e-6 else 1


if __name__ == "__main__":
    sys.exit(main())
