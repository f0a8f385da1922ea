THB_TV_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-adapter --pairs 1000 2>&1 | grep "two-view batch" | tail -4
