mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/microbench/distributed_check.py 2>&1 | grep "rank "
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02g_bench_2gpu.json').read().strip().splitlines()[-1]); print('BA', d['value'], d['e2e']['value'], 'ransac', d['ransac']['value'], d['ransac']['e2e']['value'], 'c5', d['c5']['total_ms'], d['c5']['stage_ms'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02g_bench_ref_2gpu.json 2>/dev/null; echo "ref rc=$?"; tail -c 400 gpurun_out/r02g_bench_ref_2gpu.json
