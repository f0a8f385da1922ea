#!/bin/bash
THB_ADAPTER_PROF=1 timeout 600 python tools/microbench/adapter_time.py 2>&1 | tail -6
timeout 600 python -m pytest tests/test_adapter_gpu.py -x -q -m gpu 2>&1 | tail -2
