timeout 900 python -m pytest tests/test_adapter_gpu.py tests/test_track_ba.py -m gpu -q 2>&1 | tail -30
