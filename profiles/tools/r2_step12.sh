#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?"; tail -8 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests/test_train_gpu.py -q > gpurun_out/r2_pytest_train.log 2>&1
echo "train tests rc=$?"; tail -4 gpurun_out/r2_pytest_train.log | cut -c1-300
timeout 600 python profiles/tools/prof_train_dp.py > gpurun_out/r2_prof_train_1gpu.txt 2>&1
echo "1gpu rc=$?"; grep "== world" gpurun_out/r2_prof_train_1gpu.txt; grep "bn_\|cudnn::bn\|batch_norm" gpurun_out/r2_prof_train_1gpu.txt | cut -c1-60,130-215 | head -8
timeout 600 python bench.py --workload train --steps 30 --warmup 5 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err
echo "bench train rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
