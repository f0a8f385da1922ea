run() { timeout 300 python bench.py --workload ransac --steps 4 --warmup 2 --no-c5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ms=d['ms_per_step']; sh=d['cta_phase_share']; print('$1', '%.0f pairs/s' % d['value'], '%.1f ms' % ms, {k: round(v,3) for k,v in sh.items()})"; }
THB_RANSAC_CTAS=3 run ctas3
THB_RANSAC_CTAS=4 run ctas4
