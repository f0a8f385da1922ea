set -x
mkdir -p gpurun_out
THB_K4_PROF=1 timeout 300 python tools/microbench/k4_time.py 6000 2> gpurun_out/k4prof.txt | tail -3
tail -96 gpurun_out/k4prof.txt > gpurun_out/k4prof_last.txt
wc -l gpurun_out/k4prof.txt
