mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_adapter_gpu.py -q -k "covariance or Cov" > gpurun_out/cov_test.log 2>&1
tail -60 gpurun_out/cov_test.log
