#!/bin/bash
# 8 GPUs: configs[4] iteration (dp_mode auto = replicated hand-written training step) and the weak-scaling self-play line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
timeout 900 $TR bench.py --gpus 8 --workload iteration --steps 1 --warmup 1 > gpurun_out/r2_iteration_bench_8gpu_final.json 2> gpurun_out/r2_iteration_bench_8gpu_final.err
echo "iteration 8gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iteration_bench_8gpu_final.json')); print(d['value'], d['config']['dp_mode'], d['config']['train_step'], d['phases'])"
XQ_BENCH_NO_CONFIGS3=1 timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_selfplay_bench_8gpu_final.json 2> gpurun_out/r2_selfplay_bench_8gpu_final.err
echo "bench 8gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_selfplay_bench_8gpu_final.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['e2e'])[:300])"
tail -n 3 gpurun_out/r2_iteration_bench_8gpu_final.err; tail -n 3 gpurun_out/r2_selfplay_bench_8gpu_final.err
