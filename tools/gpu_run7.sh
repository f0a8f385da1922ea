set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_track_ba.py tests/test_adapter_gpu.py -x -q -k "select or Select" 2>&1 | tail -30
