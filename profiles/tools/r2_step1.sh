#!/bin/bash
# round 2, step 1: the halo-free conv kernel (lane-masked taps) -- parity first, then timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_net_gpu.py -x -q > gpurun_out/r2_pytest_net.log 2>&1
echo "net tests rc=$?" >> gpurun_out/r2_pytest_net.log
tail -15 gpurun_out/r2_pytest_net.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -8 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_s1.json 2> gpurun_out/r2_bench_s1.err
echo "bench rc=$?"
cat gpurun_out/r2_bench_s1.json | cut -c1-1500
