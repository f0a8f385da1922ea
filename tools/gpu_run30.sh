run() { timeout 600 python bench.py --workload ba --steps 10 --warmup 3 --no-cpu-baseline --no-adapter 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['phase_ms_per_step'].items()}, d['final_cost'])"; }
THB_K3_BANDS=1 run bulk_1band
THB_K3_BANDS=2 run bulk_2bands
THB_K3_BANDS=3 run bulk_3bands
THB_K3_BANDS=4 run bulk_4bands
THB_K3_BANDS=6 run bulk_6bands
THB_K3_MODE=red THB_K3_BANDS=3 run red_3bands
