timeout 1200 python -m pytest tests/test_ba_gpu.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -3
run() { timeout 600 python bench.py --workload ba --steps 10 --warmup 3 --no-cpu-baseline --no-adapter 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['phase_ms_per_step'].items()}, d['final_cost'])"; }
run bulk
THB_K3_MODE=red run red
