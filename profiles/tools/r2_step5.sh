#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log | cut -c1-300
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_default.json'))
r=d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'fwd', r['forward_ms_isolated'])
print('e2e', json.dumps(d['e2e'])[:900])
print('cpu', json.dumps(d['cpu_baseline'])[:300])
print('extra', json.dumps(d['extra'])[:900])
print('secondary', json.dumps(d['secondary'])[:1800])
PY
