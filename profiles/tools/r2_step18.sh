#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_multileaf_gpu.py tests/test_net_gpu.py -q > gpurun_out/r2_pytest_s18.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_pytest_s18.log | cut -c1-300
timeout 900 python bench.py --workload iteration --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_iteration_bench_1gpu.json 2> gpurun_out/r2_iteration_bench_1gpu.err
echo "iteration rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iteration_bench_1gpu.json')); print(d['value'], d['phases'])"
XQ_BENCH_EVAL_LEAVES=8 XQ_BENCH_SP_LEAVES=4 timeout 900 python bench.py --workload iteration --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_iteration_bench_1gpu_multileaf.json 2> gpurun_out/r2_iteration_bench_1gpu_multileaf.err
echo "iteration multileaf rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iteration_bench_1gpu_multileaf.json')); print(d['value'], d['phases'])"
