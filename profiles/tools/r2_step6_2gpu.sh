#!/bin/bash
# 2 GPUs: NCCL tests (self-play fan-in of records, data-parallel training with the overlapped FC all-reduce, replicate mode),
# weak-scaling bench, iteration bench train phase
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -q > gpurun_out/r2_pytest_mgpu.log 2>&1
echo "mgpu tests rc=$?"; tail -15 gpurun_out/r2_pytest_mgpu.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
XQ_BENCH_NO_CONFIGS3=1 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench 2gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['e2e'])[:400])"
for mode in shard replicate; do
  XQ_BENCH_DP_MODE=$mode XQ_BENCH_ITER_GAMES=256 XQ_BENCH_ITER_EVAL=8 timeout 900 $TR bench.py --gpus 2 --workload iteration --steps 1 --warmup 1 > gpurun_out/r2_iter_2gpu_$mode.json 2> gpurun_out/r2_iter_2gpu_$mode.err
  echo "iteration $mode rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iter_2gpu_$mode.json')); print(d['value'], d['phases'])"
done
XQ_BENCH_ITER_GAMES=256 XQ_BENCH_ITER_EVAL=8 timeout 900 python bench.py --workload iteration --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_iter_1gpu_small.json 2> gpurun_out/r2_iter_1gpu_small.err
echo "iteration 1gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iter_1gpu_small.json')); print(d['value'], d['phases'])"
