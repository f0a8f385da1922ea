#!/usr/bin/env python
"""Headline benchmark: BA iterations/s on BASELINE.json configs[1]
(synthetic Pinhole reconstruction, 1k cameras / 100k points / 1M observations).

  python bench.py --gpus N --steps K --warmup W          our arm (CUDA path through the C-ABI): the line is the BA half of
                                                          the metric, its `ransac` object the RANSAC half (C4, pairs/s,
                                                          strong-scaled over the N ranks)
  python bench.py --impl reference --gpus N ...          the reference's CPU algorithm (oracle restatement,
                                                          all host cores) on the same workload/metric

A "step" is one Levenberg-Marquardt iteration of BundleAdjustReconstruction (Jacobian evaluation, Schur
complement build, reduced-camera-system solve, back-substitution, candidate cost). The K timed iterations are
the iterations of COMPLETE solves with the reference's default tolerances, each started from the same perturbed
estimate and run until Ceres' convergence test stops it (about 7 iterations on this workload; the last solve
is cut at the K-th iteration). Every solve pays its own problem setup inside the timed region. Forcing K
iterations of ONE solve with the tolerances disabled would mostly time rejected round-off steps after convergence.
BA does not shard: with --gpus N every rank runs an independent replica ("replicas only").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "BA iterations/s (1k cams, 100k pts, 1M obs)"
UNIT = "iterations/s"


def quiet_nccl():
    """NCCL prints its version banner on STDOUT (any NCCL_DEBUG level from VERSION up); rank 0's stdout must be the one JSON line."""
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"


_REAL_STDOUT = None


def capture_stdout():
    """Everything libraries write to file descriptor 1 during the run (NCCL's banner, OpenMP notices) goes to stderr; the JSON
    line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def workload_config(args):
    return {"workload": "C2 synthetic Pinhole BundleAdjustReconstruction: %d cams / %d pts / %d obs" % (
                int(round(1000 * args.scale)), int(round(100000 * args.scale)), int(round(1000000 * args.scale))),
            "intrinsics_to_optimize": "NONE", "loss": "TRIVIAL", "use_homogeneous_point_parametrization": True,
            "use_inner_iterations": False, "linear_solver": "SCHUR + dense Cholesky (exact)",
            "tolerances": "reference defaults (function 1e-6, gradient 1e-10, parameter 1e-8)",
            "protocol": "K iterations = consecutive complete solves from the same initial estimate, setup included", "l2": "working set (J planes 160 MB + S 290 MB) exceeds the 126 MB L2",
            "parallelism": "replicas only (BA is single-GPU)"}


def make_options(lib_or_oracle_default, iters):
    o = lib_or_oracle_default  # BundleAdjustmentOptions defaults (inner iterations off), at most `iters` iterations
    o.max_num_iterations = iters
    return o


def run_solves(solve_once, restore, iters):
    """`iters` LM iterations as consecutive complete solves; returns (#solves, list of summaries)."""
    remaining, sums = iters, []
    while remaining > 0:
        restore()
        summ = solve_once(remaining)
        n = summ["num_iterations"]
        assert n > 0, summ
        remaining -= n
        sums.append(summ)
    return sums


class ClockSampler(threading.Thread):
    """nvidia-smi SM clock / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        # NVML in-process when available: an nvidia-smi subprocess per sample stalls the driver for tens of ms, which is
        # visible in a 150 ms timed region
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [(0x8, 2), (0x40, 3), (0x20, 4), (0x4, 5)]  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
            while not self.stop_flag.is_set():
                r = int(reasons_fn(h))
                row = [str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)), "", "", "", ""]
                for mask, pos in bits:
                    row[pos] = "Active" if r & mask else "Not Active"
                self.samples.append(row)
                self.stop_flag.wait(0.02)
            return
        except Exception:  # noqa: BLE001
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def algorithmic_bytes_k1(prob):
    """per observation 40 B in + 160 B out; each camera (48 + 7*8 B) and point (32 B) once."""
    return prob.num_observations * 200 + prob.num_cameras * 104 + prob.num_points * 32


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles():
    """dram bytes per K1 launch from the committed ncu summary, if any."""
    path = os.path.join(ROOT, "profiles", "k1_dram_bytes.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["dram_bytes_per_launch"])
        except Exception:  # noqa: BLE001
            return None
    return None


def oracle_threads():
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm sets its team size explicitly and reports what it got."""
    from oracle import oracle_py
    return oracle_py.set_num_threads(os.cpu_count() or 1)


def oracle_solves(prob, iters):
    """The oracle under the same protocol; returns (iterations/s over wall time, wall seconds, summaries)."""
    from oracle import oracle_py
    cur = {}

    def restore():
        cur["p"] = prob.copy()

    t0 = time.time()
    sums = run_solves(lambda rem: oracle_py.ba_solve(cur["p"], make_options(oracle_py.default_options(), rem)), restore, iters)
    wall = time.time() - t0
    return iters / wall, wall, sums


def cpu_baseline(prob, iters=2):
    """The oracle (CPU restatement of the reference algorithm, OpenMP over the host cores) on the same
    workload for the first `iters` LM iterations of a solve: a reported baseline, not the target."""
    threads = oracle_threads()
    value, wall, sums = oracle_solves(prob, iters)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d LM iterations of one solve of the full C2 workload, setup included "
                      "(oracle/ba_oracle.cc, %d OpenMP threads used of %d host cores, %.1f s)" % (iters, threads, os.cpu_count() or 1, wall),
            "iter_cost": sums[0]["iter_cost"]}


RANSAC_METRIC = "RANSAC pair-verifications/s (10k pairs x 2k correspondences, FivePointRelativePose)"
PAIR_BLOCK = 8       # pairs per block of the block-cyclic pair schedule
C4_SEED = 21
# FP64 operations per unit of work of k_ransac<RelPoseEst>, measured with ncu's executed-instruction counters
# (smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on, fma = 2) - profiles/r02_ransac_flop_calibration.txt
FLOP_PER_SAMPLE = 53459.0  # five-point solve of one minimal sample
FLOP_PER_MODEL = 4500.0    # essential-matrix decomposition + cheirality vote of one candidate
FLOP_PER_DATUM = 79.6      # cheirality test + Sampson distance + cost of one correspondence


def ransac_flops(stats):
    return stats["samples_solved"] * FLOP_PER_SAMPLE + stats["models_scored"] * FLOP_PER_MODEL + stats["data_scored"] * FLOP_PER_DATUM


def ransac_config(num_pairs, world):
    return {"workload": "C4 synthetic EstimateRelativePose: %d pairs x 2000 correspondences, 60%% inliers, sigma 1e-3" % num_pairs,
            "ransac": "RANSAC, error_thresh (2e-3)^2, failure_probability 1e-4, iterations 10..1000, MLE, no LO, per-pair seed",
            "parallelism": "pair table dealt block-cyclically (blocks of %d pairs) over %d rank(s); per-rank device-side pair counter; "
                           "one all_gather per step of the result records + bit-packed inlier masks" % (PAIR_BLOCK, world),
            "l2": "each step re-reads the rank's correspondences (640 MB / ranks, larger than the 126 MB L2 up to 4 ranks)"}


def oracle_ransac_sample(num_pairs, steps):
    """Oracle port on the first `num_pairs` pairs of the C4 table, all host threads; returns (pairs/s, threads, results, masks, batch)."""
    from oracle import oracle_py
    from pytheiasfm_b200 import synthetic
    threads = oracle_threads()
    batch, _ = synthetic.make_pair_batch_indexed(range(num_pairs), n=2000, seed=C4_SEED)
    params = synthetic.c4_params(oracle_py.ransac_default_params())
    t0 = time.time()
    for _ in range(steps):
        rc, res, mask = oracle_py.ransac_relpose_batch(batch, params, threads)
    secs = (time.time() - t0) / steps
    assert rc == 0
    return num_pairs / secs, threads, res, mask, batch


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path cannot be built here (Ceres/Eigen/glog absent),
    so the oracle port stands in for it, on all host cores. Rank 0 only."""
    if rank != 0:
        return
    from pytheiasfm_b200 import synthetic
    threads = oracle_threads()
    line = {"impl": "reference"}
    if args.workload in ("both", "ba"):
        prob, _ = synthetic.config_c2(scale=args.scale)
        if args.warmup > 0:
            oracle_solves(prob, 1)
        value, wall, sums = oracle_solves(prob, args.steps)
        line.update({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                     "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                                      "sample": "%d LM iterations (%d solves) of the full workload, oracle port of the reference algorithm "
                                                "(the reference itself needs Ceres+Eigen, absent here), %d OpenMP threads used of %d host cores, wall %.1f s"
                                                % (args.steps, len(sums), threads, os.cpu_count() or 1, wall)},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0, "final_cost": sums[0]["final_cost"], "initial_cost": sums[0]["initial_cost"],
                     "iterations_per_solve": [x["num_iterations"] for x in sums]})
    if args.workload in ("both", "ransac"):
        sample = min(args.pairs, args.ref_pairs)
        if args.warmup > 0:
            oracle_ransac_sample(min(sample, 16), 1)
        value, threads, res, mask, _ = oracle_ransac_sample(sample, max(1, args.steps if args.workload == "ransac" else min(args.steps, 3)))
        r = {"metric": RANSAC_METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "scaling": "strong", "higher_is_better": True,
             "config": ransac_config(args.pairs, 1),
             "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                              "sample": "first %d of the %d pairs per step, oracle port, OpenMP over pairs with %d threads of %d host cores "
                                        "(the reference's own loop is single-threaded)" % (sample, args.pairs, threads, os.cpu_count() or 1)},
             "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
             "mean_ransac_iterations": float(res["num_iterations"].mean()), "total_inliers_sample": int(mask.sum())}
        if args.workload == "ransac":
            line.update(r)
            line.update({"steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "vs_baseline": None, "dtype": "f64",
                         "data": "synthetic", "gpu_launches": 0})
        else:
            line["ransac"] = r
    emit(line)


def run_ransac_leg(args, lib, rank, local_rank, world, stream, K, W):
    """Second half of the metric: two-view RANSAC verification throughput on C4, strong-scaled over the ranks. A step = the
    whole pair table: every rank verifies the blocks it owns, packs the inlier masks to bits and one all_gather puts the
    result records and masks of all pairs on every rank. Returns the `ransac` object (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from pytheiasfm_b200 import capi, sharding, synthetic
    total_pairs = args.pairs
    sptr = C.c_void_p(stream.cuda_stream)
    idx = sharding.block_cyclic_indices(total_pairs, rank, world, PAIR_BLOCK)
    mine, _ = synthetic.make_pair_batch_indexed(idx, n=2000, seed=C4_SEED)
    params = synthetic.c4_params(capi.ThbRansacParams())
    rec = capi.RELPOSE_DTYPE.itemsize
    npair, ncorr = mine.num_pairs, int(mine.pair_offset[-1])
    woff = sharding.mask_word_offsets(mine.pair_offset)
    nwords = int(woff[-1])
    pin_corr = torch.from_numpy(mine.corr).pin_memory(); pin_off = torch.from_numpy(mine.pair_offset).pin_memory()
    pin_seed = torch.from_numpy(mine.seed.astype(np.int64)).to(torch.int32).pin_memory()
    d_corr = torch.empty_like(pin_corr, device="cuda"); d_off = torch.empty_like(pin_off, device="cuda"); d_seed = torch.empty_like(pin_seed, device="cuda")
    d_woff = torch.from_numpy(woff).cuda()
    d_mask = torch.zeros(max(ncorr, 1), dtype=torch.uint8, device="cuda")
    payload_bytes = npair * rec + nwords * 4
    d_payload = torch.zeros(max(payload_bytes, 8), dtype=torch.uint8, device="cuda")  # [result records | packed mask words]
    res_ptr = d_payload.data_ptr(); words_ptr = res_ptr + npair * rec
    assert (npair * rec) % 8 == 0
    if world > 1:
        counts_t = torch.tensor([payload_bytes, npair, nwords], dtype=torch.int64, device="cuda")
        allc = [torch.zeros(3, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(allc, counts_t)
        counts = [int(c[0]) for c in allc]
    else:
        counts = [payload_bytes]
    pad = max(counts)
    d_send = torch.zeros(pad, dtype=torch.uint8, device="cuda")
    d_gather = torch.zeros(world * pad, dtype=torch.uint8, device="cuda")
    pin_gather = torch.zeros(world * pad, dtype=torch.uint8).pin_memory()
    b = capi.ThbPairBatch(); b.num_pairs = npair; b.memory_space = capi.THB_MEM_DEVICE
    b.pair_offset = d_off.data_ptr(); b.corr = d_corr.data_ptr(); b.seed = d_seed.data_ptr()

    def upload():
        d_corr.copy_(pin_corr, non_blocking=True); d_off.copy_(pin_off, non_blocking=True); d_seed.copy_(pin_seed, non_blocking=True)

    def step(from_host):
        if from_host:
            upload()
        if npair > 0:
            capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), C.c_void_p(res_ptr), C.c_void_p(d_mask.data_ptr()), sptr))
            capi.check(lib.thb_pack_inlier_masks(C.c_void_p(d_mask.data_ptr()), C.c_void_p(d_off.data_ptr()), C.c_void_p(d_woff.data_ptr()),
                                                 npair, capi.THB_MEM_DEVICE, C.c_void_p(words_ptr), sptr))
        if world > 1:
            d_send[:payload_bytes] = d_payload[:payload_bytes]
            dist.all_gather_into_tensor(d_gather, d_send)
            src = d_gather
        else:
            src = d_payload
        if from_host:  # the verified table (records + masks of ALL pairs) back on the host of every rank
            pin_gather[: src.numel()].copy_(src[: pin_gather.numel()], non_blocking=True)
            stream.synchronize()

    def timed(from_host):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            step(from_host)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    upload()
    for _ in range(W):
        step(False)
    sampler = ClockSampler(local_rank); sampler.start()
    ms = timed(False)
    sampler.stop_flag.set(); sampler.join()
    st = capi.ThbRansacStats()
    capi.check(lib.thb_ransac_last_stats(C.byref(st)))
    stats = st.as_dict()
    step(True)  # warm-up of the host path
    ms_e2e = timed(True)
    peak = C.c_double(0.0)
    capi.check(lib.thb_fp64_peak_tflops(5, C.byref(peak), sptr))
    flops = ransac_flops(stats)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([flops, float(stats["iterations"]), float(stats["samples_solved"]), float(stats["models_scored"]), float(stats["data_scored"]), peak.value,
                        float(stats["cycles_draw"]), float(stats["cycles_solve"]), float(stats["cycles_score"]), float(stats["cycles_scan"])],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    ms_max, e2e_max = float(t[0]), float(t[1])
    # the gathered table of the last step, in global pair order
    table = pin_gather.numpy()
    res_all = np.zeros(total_pairs, capi.RELPOSE_DTYPE)
    inliers_from_masks = np.zeros(total_pairs, np.int64)
    wpp = (2000 + 31) // 32
    for r in range(world):
        ridx = sharding.block_cyclic_indices(total_pairs, r, world, PAIR_BLOCK)
        seg = table[r * pad: r * pad + len(ridx) * rec + len(ridx) * wpp * 4]
        res_all[ridx] = np.frombuffer(seg[: len(ridx) * rec].tobytes(), capi.RELPOSE_DTYPE)
        words = np.frombuffer(seg[len(ridx) * rec:].tobytes(), np.uint32).reshape(len(ridx), wpp)
        inliers_from_masks[ridx] = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little").sum(1)
    assert (res_all["success"] == 1).all(), "a pair of the table was not verified"
    assert np.array_equal(inliers_from_masks, res_all["num_inliers"].astype(np.int64)), "gathered bit-masks disagree with the result records"
    flops_all, peak_all = float(tot[0]), float(tot[5])
    sec = ms_max / K * 1e-3
    obj = {"metric": RANSAC_METRIC, "value": total_pairs / sec, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "config": ransac_config(total_pairs, world),
           "e2e": {"value": total_pairs * K / (e2e_max * 1e-3), "unit": "pairs/s",
                   "h2d_bytes_per_step": int(mine.corr.nbytes + mine.pair_offset.nbytes + npair * 4), "d2h_bytes_per_step": int(world * pad),
                   "note": "per rank: H2D of its pairs from pinned memory, k_ransac, mask packing, all_gather, D2H of the whole verified table"},
           "gpu_launches": 2 * K, "mean_ransac_iterations": float(res_all["num_iterations"].mean()),
           "total_inliers": int(res_all["num_inliers"].sum()), "pairs_verified": int(res_all["success"].sum()),
           "gathered_bytes_per_step": int(world * pad), "clocks": sampler.summary(),
           "work_per_step": {"iterations": float(tot[1]), "samples_solved": float(tot[2]), "models_scored": float(tot[3]), "data_scored": float(tot[4])},
           "cta_phase_share": {k: float(tot[6 + i]) / max(1.0, float(tot[6:10].sum())) for i, k in enumerate(("draw", "solve", "score", "scan"))},
           "roofline": {"bound": "fp64", "kernel": "k_ransac<RelPoseEst> (persistent CTAs, one pair at a time per CTA)",
                        "achieved": flops_all / sec / 1e12, "peak": peak_all, "unit": "TFLOP/s", "frac": flops_all / sec / 1e12 / peak_all if peak_all > 0 else None,
                        "traffic": None, "flops_per_step": flops_all,
                        "flop_model": {"per_sample_solved": FLOP_PER_SAMPLE, "per_model": FLOP_PER_MODEL, "per_datum_scored": FLOP_PER_DATUM,
                                       "source": "device work counters (thb_ransac_last_stats) x ncu-calibrated FP64 ops per unit"},
                        "peak_source": "thb_fp64_peak_tflops: DFMA issue-rate probe run in this process on every rank (summed); MEASURED_PEAKS.json has no FP64 figure"}}
    if not args.no_cpu_baseline and world == 1:
        n_chk = min(64, total_pairs)
        value, threads, ores, omask, ob = oracle_ransac_sample(n_chk, 1)
        # rank 0 owns every pair at N = 1: the first n_chk records / masks are those pairs
        same = bool(np.array_equal(ores["num_iterations"], res_all["num_iterations"][:n_chk]) and
                    np.array_equal(ores["num_inliers"], res_all["num_inliers"][:n_chk]) and
                    np.array_equal(ores["essential_matrix"], res_all["essential_matrix"][:n_chk]))
        obj["cpu_baseline"] = {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                               "sample": "first %d pairs of the workload, oracle/ransac_oracle.cc, OpenMP over pairs, %d threads of %d host cores" % (n_chk, threads, os.cpu_count() or 1)}
        obj["parity"] = {"pairs_checked": n_chk, "identical_iterations_inliers_models": same}
        assert same, "RANSAC results differ from the oracle on the checked pairs"
    return obj


def adapter_e2e(prob, iterations_per_solve, repeats=3):
    """The call a pipeline user makes: pt.sfm.BundleAdjustReconstruction(options, reconstruction) through the pybind adapter on
    the full C2 scene, INCLUDING the adapter's gather of the hash-map Reconstruction into the flat arrays and the scatter
    back. The reconstruction is rebuilt to the initial estimate before every timed call (not timed)."""
    from pytheiasfm_b200 import _pt as pt
    a = prob.a
    recon = pt.sfm.Reconstruction()
    vids, tids = [], []
    for c in range(prob.num_cameras):
        v = recon.AddView(str(c), 0, float(c))
        cam = recon.View(v).MutableCamera()
        cam.SetFocalLength(float(a["intr"][0, 0])); cam.SetPrincipalPoint(float(a["intr"][0, 3]), float(a["intr"][0, 4]))
        recon.View(v).SetIsEstimated(True)
        vids.append(v)
    for p in range(prob.num_points):
        t = recon.AddTrack()
        recon.MutableTrack(t).SetIsEstimated(True)
        tids.append(t)
    oc, op, xy = a["obs_cam"], a["obs_pt"], a["obs_xy"]
    for i in range(prob.num_observations):
        recon.AddObservation(vids[oc[i]], tids[op[i]], pt.sfm.Feature(xy[i]))

    def reset():
        for c, v in enumerate(vids):
            cam = recon.View(v).MutableCamera()
            cam.SetPosition(a["cam_ext"][c, :3].copy()); cam.SetOrientationFromAngleAxis(a["cam_ext"][c, 3:].copy())
        for p, t in enumerate(tids):
            recon.MutableTrack(t).SetPoint(a["pts"][p].copy())

    opts = pt.sfm.BundleAdjustmentOptions()
    opts.use_inner_iterations = False
    walls, solves, setups, cost = [], [], [], None
    for r in range(repeats + 1):
        reset()
        t0 = time.perf_counter()
        summ = pt.sfm.BundleAdjustReconstruction(opts, recon)
        wall = time.perf_counter() - t0
        if r > 0:
            walls.append(wall); solves.append(summ.solve_time_in_seconds); setups.append(summ.setup_time_in_seconds)
        cost = summ.final_cost
    wall = float(np.median(walls))
    return {"value": iterations_per_solve / wall, "unit": UNIT, "ms_per_call": 1e3 * wall, "iterations_per_call": iterations_per_solve,
            "solver_ms": 1e3 * float(np.median(solves)), "device_setup_ms": 1e3 * float(np.median(setups)),
            "gather_scatter_ms": 1e3 * (wall - float(np.median(solves)) - float(np.median(setups))), "final_cost": cost,
            "note": "pt.sfm.BundleAdjustReconstruction on a Reconstruction of hash maps (1 call = 1 complete solve); gather_scatter = wall - device setup - solve"}


def run_c5_leg(args, lib, rank, local_rank, world, stream, K):
    """BASELINE configs[4]: the hot-path stages of a global SfM pipeline on the south-building-shaped scene, end to end from
    host buffers: two-view verification of every image pair (TwoViewMatchGeometricVerification, pairs dealt over the ranks),
    TrackEstimator (triangulation + per-track BA), BundleAdjustReconstruction with the pipeline's options (Huber width 10,
    inner iterations) and SetOutlierTracksToUnestimated. View-graph filtering and rotation / position averaging between
    verification and the track stage are outside the hot path: the cameras enter the track stage as the
    generator's poses plus noise. Returns the `c5` object (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from pytheiasfm_b200 import capi, sharding, synthetic
    sptr = C.c_void_p(stream.cuda_stream)
    scene = synthetic.config_c5()
    pairs, intr, prob0 = scene["pairs"], scene["intrinsics"], scene["problem"]
    idx = sharding.block_cyclic_indices(pairs.num_pairs, rank, world, 4)
    mine = capi.HostPairBatch([pairs.corr[pairs.pair_offset[i]:pairs.pair_offset[i + 1]] for i in idx], pairs.seed[idx])
    mi = np.ascontiguousarray(intr[idx])
    tvo = capi.ThbTwoViewOptions(); lib.thb_two_view_default_options(C.byref(tvo))
    info = np.zeros(mine.num_pairs, capi.TWO_VIEW_INFO_DTYPE); vmask = np.zeros(int(mine.pair_offset[-1]), np.uint8)
    rays = synthetic.pinhole_rays(prob0)
    teo = capi.ThbTrackEstimatorOptions()
    bo = capi.default_options(lib)
    bo.loss_function_type = capi.LOSS_HUBER; bo.robust_loss_width = 10.0; bo.use_inner_iterations = 1; bo.max_num_iterations = 50
    tbo = capi.default_options(lib)
    counts = [len(sharding.block_cyclic_indices(pairs.num_pairs, r, world, 4)) for r in range(world)]
    rec = capi.TWO_VIEW_INFO_DTYPE.itemsize
    out = {}

    def vp(a):
        return a.ctypes.data_as(C.c_void_p)

    def once():
        t = {}
        t0 = time.perf_counter()
        if mine.num_pairs:
            b = mine.struct()
            capi.check(lib.thb_verify_two_view_matches_batch(C.byref(b), vp(mi), vp(mi), C.byref(tvo), vp(info), vp(vmask), sptr))
        if world > 1:
            local = torch.from_numpy(info.view(np.uint8).reshape(-1).copy()).cuda()
            parts = sharding.all_gather_padded(local, [c * rec for c in counts])
            out["info_all"] = np.concatenate([np.frombuffer(x.cpu().numpy().tobytes(), capi.TWO_VIEW_INFO_DTYPE) for x in parts])
        else:
            out["info_all"] = info
        torch.cuda.synchronize()
        t["verification"] = time.perf_counter() - t0
        if rank == 0:
            prob = prob0.copy()
            status = np.zeros(prob.num_points, np.int32)
            t0 = time.perf_counter()
            p = prob.struct()
            capi.check(lib.thb_estimate_tracks_batch(C.byref(p), vp(rays), C.byref(teo), C.byref(tbo), vp(status), None, sptr))
            t["tracks"] = time.perf_counter() - t0
            # BundleAdjustReconstruction adds estimated tracks only (bundle_adjuster.cc:142,178): drop the others' observations
            t0 = time.perf_counter()
            keep = np.flatnonzero(status[prob.a["obs_pt"]] == capi.TRACK_ESTIMATED)
            a = dict(prob.a)
            for k in ("obs_cam", "obs_pt", "obs_xy", "obs_sqrt_info"):
                a[k] = np.take(prob.a[k], keep, axis=0)   # (a boolean mask on the [n, 2] arrays is 8x slower than take)
            a["pt_const"] = (status != capi.TRACK_ESTIMATED).astype(np.uint8)
            prob = capi.HostBaProblem(a)
            t["select_estimated"] = time.perf_counter() - t0
            summ = capi.ThbBaSummary()
            t0 = time.perf_counter()
            p = prob.struct()
            capi.check(lib.thb_ba_solve(C.byref(p), C.byref(bo), C.byref(summ), sptr))
            t["bundle_adjustment"] = time.perf_counter() - t0
            ostat = np.zeros(prob.num_points, np.int32); removed = C.c_int32(0)
            t0 = time.perf_counter()
            p = prob.struct()
            capi.check(lib.thb_set_outlier_tracks_batch(C.byref(p), 4.0, 2.0, vp(ostat), C.byref(removed), sptr))
            t["outlier_filter"] = time.perf_counter() - t0
            out.update(status=status, summ=summ.as_dict(), removed=removed.value, prob=prob)
        if world > 1:
            dist.barrier()
        return t

    once()
    runs = [once() for _ in range(K)]
    if rank != 0:
        return None
    stages = {k: 1e3 * float(np.mean([r[k] for r in runs])) for k in runs[0]}
    total = sum(stages.values())
    gt = scene["gt"]
    est = out["status"] == capi.TRACK_ESTIMATED
    X = out["prob"].a["pts"]
    err = np.linalg.norm(X[est, :3] / X[est, 3:4] - gt["pts"][est], axis=1)
    obj = {"metric": "C5 hot-path pipeline (verification + tracks + BA + filter) runs/s", "value": 1e3 / total, "unit": "runs/s", "n_gpus": world, "steps": K,
           "config": {"workload": "C5 south-building-shaped synthetic: %d cams, %d tracks, %d observations, %d image pairs x ~%d matches" % (
               prob0.num_cameras, prob0.num_points, prob0.num_observations, pairs.num_pairs, int(np.diff(pairs.pair_offset).mean())),
               "stages": "VerifyMatches batched over the ranks (pairs dealt block-cyclically), then on rank 0: TrackEstimator, BundleAdjustReconstruction "
                         "(Huber 10, inner iterations, <= 50 iterations), SetOutlierTracksToUnestimated; all from host buffers",
               "not_timed": "view-graph filtering, rotation / position averaging (outside the hot path)"},
           "stage_ms": stages, "total_ms": total,
           "pairs_verified": int(out["info_all"]["success"].sum()), "pairs": int(pairs.num_pairs),
           "tracks_estimated": int(est.sum()), "ba": {k: out["summ"][k] for k in ("num_iterations", "initial_cost", "final_cost", "termination_type")},
           "tracks_removed_by_filter": int(out["removed"]), "median_point_error": float(np.median(err))}
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle_py
        threads = oracle_threads()
        nsub = 64
        sub = capi.HostPairBatch([pairs.corr[pairs.pair_offset[i]:pairs.pair_offset[i + 1]] for i in range(nsub)], pairs.seed[:nsub])
        t0 = time.time(); rc, oinfo, omask = oracle_py.two_view_batch(sub, intr[:nsub], intr[:nsub], tvo, True); tv = time.time() - t0
        same = bool(np.array_equal(oinfo["success"], out["info_all"]["success"][:nsub]) and
                    np.array_equal(oinfo["num_verified_matches"], out["info_all"]["num_verified_matches"][:nsub]))
        obj["cpu_baseline"] = {"kind": "port", "cores": threads, "verification_ms_all_pairs": 1e3 * tv * pairs.num_pairs / nsub,
                               "sample": "oracle VerifyMatches on the first %d of the %d pairs, scaled to all pairs (%d threads)" % (nsub, pairs.num_pairs, threads),
                               "verification_identical_on_sample": same}
        assert same, "C5 verification differs from the oracle on the sampled pairs"
    return obj


def run_ba_leg(args, lib, rank, local_rank, world, stream, K, W):
    """First half of the metric: BA iterations/s on C2. Returns the line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from pytheiasfm_b200 import capi, synthetic
    sptr = C.c_void_p(stream.cuda_stream)
    prob, _ = synthetic.config_c2(scale=args.scale)  # every rank: an identical replica

    def make_arm(tensors, space):
        """(solve_once, restore) over one set of buffers; the solve refines cam_ext / pts in place, like the reference."""
        p = prob.struct()
        p.memory_space = space
        for k, v in tensors.items():
            setattr(p, k, None if v is None else v.data_ptr())
        init = {k: tensors[k].clone() for k in ("cam_ext", "pts", "intr")}

        def restore():
            for k, v in init.items():
                tensors[k].copy_(v)

        def solve_once(max_iters):
            summ = capi.ThbBaSummary()
            opts = make_options(capi.default_options(lib), max_iters)
            capi.check(lib.thb_ba_solve(C.byref(p), C.byref(opts), C.byref(summ), sptr))
            return summ.as_dict()
        return solve_once, restore

    # ---- device-resident arm: inputs already in HBM when the timed region starts ----
    dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in prob.a.items()}
    solve_dev, restore_dev = make_arm(dev, capi.THB_MEM_DEVICE)
    run_solves(solve_dev, restore_dev, W)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sums = run_solves(solve_dev, restore_dev, K)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag.set()
    sampler.join()
    assert sum(x["num_iterations"] for x in sums) == K
    launches_timed = sum(x["gpu_launches"] for x in sums)
    phase = {k: sum(x[k] for x in sums) / K for k in ("ms_jacobian", "ms_normal", "ms_solve", "ms_update")}
    setup_ms = 1e3 * sum(x["setup_time_in_seconds"] for x in sums) / len(sums)

    # roofline of the dominant HBM-bound kernel (K1), timed alone with an L2 flush between launches
    restore_dev()
    pd = prob.struct()
    pd.memory_space = capi.THB_MEM_DEVICE
    for k, v in dev.items():
        setattr(pd, k, None if v is None else v.data_ptr())
    sess = C.c_void_p()
    o1 = make_options(capi.default_options(lib), 1)
    capi.check(lib.thb_ba_create(C.byref(pd), C.byref(o1), sptr, C.byref(sess)))
    k1_ms = C.c_double(0.0)
    capi.check(lib.thb_ba_time_jacobian(sess, 20, 1, C.byref(k1_ms)))
    capi.check(lib.thb_ba_finish(sess, None))
    fp64_peak = C.c_double(0.0)
    capi.check(lib.thb_fp64_peak_tflops(5, C.byref(fp64_peak), sptr))

    # ---- the reference's other solver option on the same workload: ITERATIVE_SCHUR (THB_SOLVER_SCHUR_PCG), complete solves ----
    def solve_pcg():
        restore_dev()
        summ_ = capi.ThbBaSummary()
        opts_ = make_options(capi.default_options(lib), 100)
        opts_.linear_solver = capi.SOLVER_SCHUR_PCG
        capi.check(lib.thb_ba_solve(C.byref(pd), C.byref(opts_), C.byref(summ_), sptr))
        return summ_.as_dict()
    solve_pcg()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    pcg_sums = [solve_pcg() for _ in range(5)]
    p1.record(stream)
    torch.cuda.synchronize()
    pcg_ms = p0.elapsed_time(p1) / len(pcg_sums)
    ps = pcg_sums[0]
    pcg = {"linear_solver": "ITERATIVE_SCHUR + SCHUR_JACOBI (eta 0.1), complete solves to the reference's default tolerances, inputs in HBM",
           "ms_per_solve": pcg_ms, "iterations_per_solve": ps["num_iterations"], "cg_iterations_per_solve": ps["num_linear_solver_iterations"],
           "iterations_per_s": ps["num_iterations"] / (pcg_ms * 1e-3), "ms_linear_solve": ps["ms_solve"] / max(1, ps["num_linear_solves"]),
           "final_cost": ps["final_cost"], "exact_solver_ms_per_solve": ms / len(sums)}

    # ---- end-to-end arm: the call a user makes, pinned HOST buffers, H2D + setup + iterations + D2H timed ----
    pin = {k: (None if v is None else torch.from_numpy(v.copy()).pin_memory()) for k, v in prob.a.items()}
    solve_host, restore_host = make_arm(pin, capi.THB_MEM_HOST)
    run_solves(solve_host, restore_host, W)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sums_h = run_solves(solve_host, restore_host, K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = len(sums_h) * sum(v.numel() * v.element_size() for v in pin.values() if v is not None)
    d2h = len(sums_h) * (pin["cam_ext"].numel() * 8 + pin["pts"].numel() * 8) + K * 96
    summ = sums[0]

    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = float(t[0]), float(t[1])
    if rank != 0:
        return None
    peak, peak_src = measured_peak_hbm()
    ab = algorithmic_bytes_k1(prob)
    achieved = ab / (k1_ms.value * 1e-3) / 1e9
    k4_flops = (6.0 * prob.num_cameras) ** 3 / 3.0
    line = {
        "metric": METRIC, "value": world * K / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": world * K / e2e_max, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                "note": "thb_ba_solve calls on pinned host buffers: H2D + device-side setup + iterations + D2H of the refined "
                        "parameters; the per-iteration scalar read-back (96 B) is in d2h"},
        "gpu_launches": launches_timed,
        "roofline": {"bound": "hbm", "kernel": "k_jacobian_sc (K1, shared-memory camera table, materialised tangent-space Jacobian planes)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_from_profiles(), "peak_source": peak_src + ", burst",
                     "algorithmic_bytes": ab, "avg_launch_ms": k1_ms.value},
        # the kernel that takes most of the step (K4, dense FP64 Cholesky of the reduced camera system: n^3/3 flops per
        # iteration) against the FP64 FMA rate measured in this process (thb_fp64_peak_tflops)
        "roofline_k4": {"bound": "fp64", "kernel": "DenseChol (K4: dense FP64 Cholesky of the reduced camera system)",
                        "achieved": k4_flops / (phase["ms_solve"] * 1e-3) / 1e12 if phase["ms_solve"] > 0 else None,
                        "peak": fp64_peak.value, "unit": "TFLOP/s",
                        "frac": k4_flops / (phase["ms_solve"] * 1e-3) / 1e12 / fp64_peak.value if phase["ms_solve"] > 0 and fp64_peak.value > 0 else None,
                        "peak_source": "thb_fp64_peak_tflops: DFMA issue-rate probe run in this process (DMMA m8n8k4 issues at the same rate); MEASURED_PEAKS.json has no FP64 figure",
                        "share_of_step": phase["ms_solve"] / (ms / K)},
        "phase_ms_per_step": {"jacobian": phase["ms_jacobian"], "normal_equations": phase["ms_normal"],
                              "reduced_solve": phase["ms_solve"], "update_and_cost": phase["ms_update"]},
        "solves": len(sums), "iterations_per_solve": [x["num_iterations"] for x in sums], "setup_ms_per_solve": setup_ms,
        "successful_steps": sum(x["num_successful_steps"] for x in sums),
        "final_cost": summ["final_cost"], "initial_cost": summ["initial_cost"],
        "termination_type": summ["termination_type"], "clocks": sampler.summary(),
    }
    line["iterative_schur"] = pcg
    assert abs(pcg["final_cost"] - summ["final_cost"]) <= 1e-4 * summ["final_cost"], ("ITERATIVE_SCHUR ends at a different minimum", pcg["final_cost"])
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(prob, iters=2)
        ours, theirs = summ["iter_cost"][:3], cb.pop("iter_cost")[:3]
        rel = max(abs(a - b) / abs(b) for a, b in zip(ours, theirs))
        line["cpu_baseline"] = cb
        line["parity"] = {"iter_cost_gpu": ours, "iter_cost_oracle": theirs, "max_rel_diff": rel, "tolerance": 1e-6}
        assert len(ours) == len(theirs) and rel <= 1e-6, ("BA cost sequence differs from the oracle", ours, theirs)
    if world == 1 and not args.no_adapter:
        line["adapter_e2e"] = adapter_e2e(prob, summ["num_iterations"])
        assert abs(line["adapter_e2e"]["final_cost"] - summ["final_cost"]) <= 1e-6 * summ["final_cost"]
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="both", choices=["both", "ba", "ransac"],
                    help="both = the two halves of BASELINE.json's metric: BA (configs[1]) as the line, RANSAC (configs[3]) as its `ransac` object")
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--ref-pairs", type=int, default=512, help="--impl reference: pairs of the C4 table per step (bounded sample)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the BA workload (debug only; 1.0 = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 pipeline leg (configs[4])")
    ap.add_argument("--no-adapter", action="store_true", help="skip the pybind adapter timing")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    capture_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pytheiasfm_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_nccl()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.load_library()  # raises if the CUDA library is missing: no fallback
    W, K = max(args.warmup, 3), args.steps
    stream = torch.cuda.current_stream()
    line = None
    if args.workload in ("both", "ba"):
        line = run_ba_leg(args, lib, rank, local_rank, world, stream, K, W)
    if args.workload in ("both", "ransac"):
        r = run_ransac_leg(args, lib, rank, local_rank, world, stream, K if args.workload == "ransac" else min(K, 10), W)
        if rank == 0:
            if line is None:
                line = dict(r, vs_baseline=None, dtype="f64", data="synthetic")
            else:
                line["ransac"] = r
    if args.workload == "both" and not args.no_c5:
        c5 = run_c5_leg(args, lib, rank, local_rank, world, stream, min(K, 5))
        if rank == 0:
            line["c5"] = c5
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
