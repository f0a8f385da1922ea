mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_track_ba.py tests/test_adapter_gpu.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/microbench/track_time.py 2>&1 | tail -8
