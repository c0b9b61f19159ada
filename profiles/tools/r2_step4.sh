#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_rng_gpu.py tests/test_multileaf_gpu.py tests/test_selfplay_gpu.py -q > gpurun_out/r2_pytest_s4.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r2_pytest_s4.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_s4.json 2> gpurun_out/r2_bench_s4.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_s4.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], r['frac'], r['forward_ms_isolated'], r['dominant_kernel']['ms_per_launch'], d['clocks'])
PY
XQ_BENCH_SIMS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2_selfplay_launches_s4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches_s4.log 2>&1
echo "ncu launches rc=$?"
