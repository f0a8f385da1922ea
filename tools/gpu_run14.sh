mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ransac_gpu.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -3
run() { timeout 300 python bench.py --workload ransac --steps 4 --warmup 2 --no-c5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ms=d['ms_per_step']; sh=d['cta_phase_share']; print('$1', '%.0f pairs/s' % d['value'], '%.1f ms' % ms, {k: round(v*ms,1) for k,v in sh.items()})"; }
THB_RS_SOLVE_SMEM=1 run smem
THB_RS_SOLVE_SMEM=0 run local
THB_RS_PROF=1 timeout 300 python bench.py --workload ransac --steps 1 --warmup 1 --no-c5 --no-cpu-baseline 2>&1 | grep RSPROF | tail -5
