#!/bin/bash
THB_TV_TIMING=1 timeout 300 python tools/microbench/c5_verify_stats.py 2>&1 | grep -E "call|two-view batch|rounds<" | tail -4
timeout 900 python -m pytest tests/test_ransac_gpu.py tests/test_two_view_gpu.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -2
THB_RANSAC_MODE=rounds timeout 300 python tools/microbench/ransac_small_batches.py 2>&1 | tail -8
