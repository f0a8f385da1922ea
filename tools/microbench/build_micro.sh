#!/bin/bash
# builds tools/microbench/k4_micro (needs SetLastError/GetLastError from ba_solver.o; run __graft_entry__.build() first)
set -e
cd "$(dirname "$0")/.."
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  -o tools/microbench/k4_micro.bin tools/microbench/k4_micro.cu build/ba_solver.o build/ba_setup.o -cudart static
