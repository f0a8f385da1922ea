mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02e_pytest_gpu.log
timeout 600 python bench.py --workload ba --steps 10 --warmup 3 --no-cpu-baseline --no-adapter 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BA', d['value'], d['e2e']['value'], d['phase_ms_per_step'])"
