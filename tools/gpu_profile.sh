# r02 profile captures (run on the B200 box through gpurun); summaries are written by tools/microbench/*.py here afterwards
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_ba.csv python bench.py --workload ba --steps 6 --warmup 1 --no-cpu-baseline --no-adapter > gpurun_out/r02_launches_ba.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_ll_kernel -c 1 -f -o gpurun_out/r02_k4_full python tools/microbench/k4_time.py 6000 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ransac -c 1 -f -o gpurun_out/r02_ransac_full python tools/ransac_flop_calibration.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobian_sc -s 2 -c 1 -f -o gpurun_out/r02_k1_full python bench.py --workload ba --steps 3 --warmup 1 --no-cpu-baseline --no-adapter > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
