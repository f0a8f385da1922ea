timeout 900 python -m pytest tests/test_ransac_gpu.py tests/test_two_view_gpu.py tests/test_golden.py tests/test_adapter_gpu.py -q -m gpu 2>&1 | tail -4
run() { timeout 300 python bench.py --workload ransac --steps 5 --warmup 2 --no-c5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ms=d['ms_per_step']; sh=d['cta_phase_share']; print('$1', '%.0f pairs/s' % d['value'], '%.1f ms' % ms, {k: round(v,3) for k,v in sh.items()})"; }
run fused_default
THB_RANSAC_MODE=rounds run rounds
