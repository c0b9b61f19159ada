#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -x -q > gpurun_out/r2_pytest_net2.log 2>&1
rc=$?; echo "net tests (2cta) rc=$rc"; tail -12 gpurun_out/r2_pytest_net2.log | cut -c1-300
if [ $rc -ne 0 ]; then
  echo "2-CTA kernel failed: single-CTA check"; XQ_NET_2CTA=0 timeout 300 python -m pytest tests/test_net_gpu.py -x -q 2>&1 | tail -3
  exit 0
fi
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_selfplay_gpu.py tests/test_multileaf_gpu.py -q > gpurun_out/r2_pytest_s10.log 2>&1
echo "more tests rc=$?"; tail -5 gpurun_out/r2_pytest_s10.log | cut -c1-300
for m in 1 0; do
XQ_NET_2CTA=$m XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2cta$m.json 2> gpurun_out/r2_bench_2cta$m.err
echo "bench 2cta=$m rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_2cta$m.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], r['frac'], r['forward_ms_isolated'], r['dominant_kernel']['ms_per_launch'], d['clocks'])
PY
done
