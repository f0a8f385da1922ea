mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_schur_offdiag -s 2 -c 1 -f -o gpurun_out/r02_k3_bulk python bench.py --workload ba --steps 3 --warmup 1 --no-cpu-baseline --no-adapter > /dev/null 2>&1
ls -la gpurun_out/r02_k3_bulk.ncu-rep
