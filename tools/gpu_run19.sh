timeout 900 python -m pytest tests/test_ransac_gpu.py tests/test_two_view_gpu.py tests/test_adapter_gpu.py -q -m gpu 2>&1 | tail -12
