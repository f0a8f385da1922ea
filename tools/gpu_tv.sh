#!/bin/bash
THB_TV_TIMING=1 timeout 300 python tools/microbench/c5_verify_stats.py 2>&1 | grep -E "call|two-view batch|rounds<" | tail -8
timeout 600 python -m pytest tests/test_two_view_gpu.py -x -q -m gpu 2>&1 | tail -2
