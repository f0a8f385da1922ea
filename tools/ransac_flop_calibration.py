#!/usr/bin/env python
"""Run under ncu to count the FP64 instructions of the RANSAC kernels (profiles/r02_ransac_flop_calibration.txt):

  ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/ransac_flops.csv python tools/ransac_flop_calibration.py

Launch 1: k_five_point on 4096 minimal samples drawn from C4 pairs (FP64 ops per five-point solve).
Launch 2: k_ransac<RelPoseEst> on 256 C4 pairs; the script prints the device work counters of that launch so that
          flops = samples * F_sample + models * F_model + data * F_datum can be solved for the per-unit constants bench.py uses."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytheiasfm_b200 import capi, synthetic  # noqa: E402

lib = capi.load_library()
batch, _ = synthetic.make_pair_batch_indexed(range(256), n=2000, seed=21)
rng = np.random.default_rng(0)
x1 = np.zeros((4096, 5, 2)); x2 = np.zeros((4096, 5, 2))
for i in range(4096):
    p = i % 256
    c = batch.corr[batch.pair_offset[p]: batch.pair_offset[p + 1]]
    s = c[rng.choice(len(c), 5, replace=False)]
    x1[i] = s[:, :2]; x2[i] = s[:, 2:]
E = np.zeros((4096, 10, 3, 3)); n = np.zeros(4096, np.int32)
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
capi.check(lib.thb_five_point_relative_pose(vp(x1), vp(x2), 4096, vp(E), vp(n), None))
params = synthetic.c4_params(capi.ThbRansacParams())
res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE)
mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
b = batch.struct()
capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), vp(res), vp(mask), None))
st = capi.ThbRansacStats()
capi.check(lib.thb_ransac_last_stats(C.byref(st)))
print(json.dumps({"five_point_samples": 4096, "mean_solutions": float(n.mean()), "ransac": st.as_dict()}))
