THB_RS_PROF=1 THB_RS_SOLVE_CTAS=3 timeout 300 python bench.py --workload ransac --steps 1 --warmup 1 --no-c5 --no-cpu-baseline 2>&1 | grep RSPROF | tail -12
THB_RS_PROF=1 THB_RS_SOLVE_CTAS=1 timeout 300 python bench.py --workload ransac --steps 1 --warmup 1 --no-c5 --no-cpu-baseline 2>&1 | grep RSPROF | tail -6
