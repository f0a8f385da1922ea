mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02g_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02g_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r02g_bench_1gpu.json 2> gpurun_out/r02g_bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r02g_bench_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/r02g_bench_ref.json 2> gpurun_out/r02g_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02g_launches_ba.csv python bench.py --workload ba --steps 6 --warmup 1 --no-cpu-baseline --no-adapter > /dev/null 2>&1
ls -la gpurun_out/r02g_launches_ba.csv
