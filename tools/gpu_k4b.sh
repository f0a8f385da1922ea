#!/bin/bash
# backsolve hand-off change: K4 tests, timing, bench BA leg
timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu -k "k4 or c2_full or c1" 2>&1 | tail -3
timeout 120 python tools/microbench/k4_time.py 6000 2500 300
timeout 300 python bench.py --workload ba --steps 20 --warmup 3 --no-cpu-baseline --no-adapter 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d.get('phases_ms'), d['roofline_k4']['frac'])"
