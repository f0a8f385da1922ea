timeout 900 python bench.py --no-adapter > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo rc=$?; tail -5 gpurun_out/r02b_bench.err
