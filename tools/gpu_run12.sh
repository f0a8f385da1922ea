mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload ransac --steps 4 --warmup 2 --no-c5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ms=d['ms_per_step']; sh=d['cta_phase_share']; print('$1', '%.0f pairs/s' % d['value'], '%.1f ms' % ms, {k: round(v*ms,1) for k,v in sh.items()})"; }
THB_RS_SOLVE_CTAS=1 THB_RS_SCORE_MODE=0 run solve1_score0
THB_RS_SOLVE_CTAS=2 THB_RS_SCORE_MODE=1 run solve2_score1
THB_RS_SOLVE_CTAS=3 THB_RS_SCORE_MODE=2 run solve3_score2
THB_RS_SOLVE_CTAS=4 THB_RS_SCORE_MODE=1 run solve4_score1
