mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_adapter_gpu.py -q -m gpu -k "prior or Prior or covariance" > gpurun_out/gprior_test.log 2>&1; tail -30 gpurun_out/gprior_test.log
