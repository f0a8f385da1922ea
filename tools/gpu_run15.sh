THB_RANSAC_MODE=fused timeout 300 python tools/microbench/ransac_small_batches.py 2>&1 | tail -8
THB_RANSAC_MODE=rounds timeout 300 python tools/microbench/ransac_small_batches.py 2>&1 | tail -8
