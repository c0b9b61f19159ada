#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_selfplay_gpu.py tests/test_multileaf_gpu.py tests/test_dropin_gpu.py -q > gpurun_out/r2_pytest_s15.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_pytest_s15.log | cut -c1-300
XQ_BENCH_NO_CONFIGS3=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_s15.json 2> gpurun_out/r2_bench_s15.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_s15.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], r['frac'], r['forward_ms_isolated'], r.get('forward_ms_back_to_back_400ms'), r['dominant_kernel']['ms_per_launch'], d['clocks'])
print(json.dumps(d['e2e'])[:700])
PY
