# final validation of the round: full GPU test suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02f_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r02f_bench_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "ref rc=$?"
