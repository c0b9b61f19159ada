#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -q > gpurun_out/r2_pytest_mgpu.log 2>&1
echo "mgpu tests rc=$?"; tail -15 gpurun_out/r2_pytest_mgpu.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
XQ_BENCH_DP_MODE=shard XQ_BENCH_ITER_GAMES=256 XQ_BENCH_ITER_EVAL=8 timeout 900 $TR bench.py --gpus 2 --workload iteration --steps 1 --warmup 1 > gpurun_out/r2_iter_2gpu_shard.json 2> gpurun_out/r2_iter_2gpu_shard.err
echo "iteration shard rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iter_2gpu_shard.json')); print(d['value'], d['phases'])"
