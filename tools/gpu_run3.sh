timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_adapter_gpu.py -m gpu -x -q -k "inner or BundleAdjustReconstruction or invalid" 2>&1 | tail -25
