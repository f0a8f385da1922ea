run() { timeout 600 python bench.py --workload ba --steps 10 --warmup 3 --no-cpu-baseline --no-adapter 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['phase_ms_per_step'].items()}, d['final_cost'])"; }
THB_K3_BULK_ROWS=6 run bulk6
THB_K3_BULK_ROWS=5 run bulk5
THB_K3_BULK_ROWS=4 run bulk4
THB_K3_BULK_ROWS=3 run bulk3
THB_K3_BULK_ROWS=2 run bulk2
