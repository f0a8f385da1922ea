mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg$ -s 4 -c 1 -f -o gpurun_out/r02_pcg_full python tools/microbench/pcg_time.py > gpurun_out/pcg_ncu.log 2>&1; tail -3 gpurun_out/pcg_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pcg -c 30 --csv python tools/microbench/pcg_time.py 2>/dev/null | grep -i "k_pcg" | head -12 | cut -d, -f5,12-15
