mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ransac_gpu.py tests/test_two_view_gpu.py tests/test_golden.py -q -m gpu -x > gpurun_out/rs_test.log 2>&1; tail -15 gpurun_out/rs_test.log
timeout 600 python bench.py --workload ransac --steps 5 --warmup 2 --no-c5 > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err; tail -c 1500 gpurun_out/rs_bench.json; tail -5 gpurun_out/rs_bench.err
THB_RANSAC_MODE=fused timeout 600 python bench.py --workload ransac --steps 5 --warmup 2 --no-c5 --no-cpu-baseline > gpurun_out/rs_bench_fused.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/rs_bench_fused.json').read().strip().splitlines()[-1]); print('fused', d['value'], d['e2e']['value'])"
