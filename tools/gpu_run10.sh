mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; tail -c 600 gpurun_out/r02c_bench_1gpu.json; tail -3 gpurun_out/r02c_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; tail -c 900 gpurun_out/r02c_bench_ref.json
