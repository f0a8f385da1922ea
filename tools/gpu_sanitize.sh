#!/bin/bash
# compute-sanitizer passes over the small parity tests (memcheck everywhere, racecheck on the shared-memory heavy kernels)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; timeout 900 $S "$@" > gpurun_out/r02_sanitize_$name.log 2>&1; echo "$name rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitize_$name.log | tail -3; }
run mem_ransac --tool memcheck --error-exitcode 9 python -m pytest tests/test_ransac_gpu.py -x -q -m gpu -k "lmed or prosac or fused_and_rounds or kats or identical_inlier"
run mem_ba --tool memcheck --error-exitcode 9 python -m pytest tests/test_ba_gpu.py -x -q -m gpu -k "prior or pcg or k4 or covariance or inner"
run mem_tv --tool memcheck --error-exitcode 9 python -m pytest tests/test_two_view_gpu.py tests/test_track_ba.py -x -q -m gpu
run race_ransac --tool racecheck --error-exitcode 9 python -m pytest tests/test_ransac_gpu.py -x -q -m gpu -k "lmed_matches_oracle and 37 or prosac"
run race_ba --tool racecheck --error-exitcode 9 python -m pytest tests/test_ba_gpu.py -x -q -m gpu -k "orientation_inner or k4_dense and 300"
