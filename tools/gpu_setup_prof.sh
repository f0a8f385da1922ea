#!/bin/bash
THB_SETUP_PROF=1 timeout 300 python bench.py --workload ba --steps 20 --warmup 3 --no-cpu-baseline --no-adapter 2> gpurun_out/setup_prof.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['setup_ms_per_solve'], d['roofline_k4']['frac'])"
grep THB_SETUP_PROF gpurun_out/setup_prof.log | tail -4
