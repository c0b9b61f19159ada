#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log | cut -c1-300
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_default.json'))
r=d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'fwd', r['forward_ms_isolated'], r.get('forward_ms_back_to_back_400ms'), d['clocks'])
print('e2e', json.dumps(d['e2e'])[:1200])
print('extra', json.dumps(d['extra'])[:600])
PY
