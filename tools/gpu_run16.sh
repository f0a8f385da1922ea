mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02d_pytest_gpu.log
timeout 600 python bench.py --workload ransac --steps 5 --warmup 2 --no-c5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ransac', d['value'], d['e2e']['value'], d['cta_phase_share'])"
