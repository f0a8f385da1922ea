set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_tests.txt
python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_bench.err
ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_ransac_flops.csv python tools/ransac_flop_calibration.py > gpurun_out/r02a_ransac_flops.json 2>&1
cat gpurun_out/r02a_tests.txt
python bench.py --impl reference --steps 6 > gpurun_out/r02a_bench_ref.json 2>/dev/null
