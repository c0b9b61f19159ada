#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_selfplay_gpu.py tests/test_multileaf_gpu.py tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_rng_gpu.py -q > gpurun_out/r2_pytest_s17.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_pytest_s17.log | cut -c1-300
for g in 1 0; do
XQ_STEP_GRAPH=$g XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_graph$g.json 2> gpurun_out/r2_bench_graph$g.err
echo "bench graph=$g rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_graph$g.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['forward_ms_isolated'], r['dominant_kernel']['ms_per_launch'], d['gpu_launches'], d['clocks']['sm_mhz'])"
XQ_STEP_GRAPH=$g XQ_BENCH_GAMES=64 XQ_BENCH_SIMS=200 XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_small_graph$g.json 2> gpurun_out/r2_bench_small_graph$g.err
echo "small bench graph=$g rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_small_graph$g.json')); print(d['value'], d['ms_per_step'])"
done
timeout 900 python bench.py --workload iteration --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_iteration_bench_1gpu.json 2> gpurun_out/r2_iteration_bench_1gpu.err
echo "iteration rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_iteration_bench_1gpu.json')); print(d['value'], d['phases'])"
