#!/bin/bash
# round 2, step 2: leaf compaction + multi-leaf search + seam exports + N=224 FC -- tests, then timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_net_gpu.py tests/test_multileaf_gpu.py tests/test_rng_gpu.py tests/test_seam_gpu.py -x -q > gpurun_out/r2_pytest_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/r2_pytest_new.log
tail -30 gpurun_out/r2_pytest_new.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -30 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_s2.json 2> gpurun_out/r2_bench_s2.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_s2.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], r['frac'], r['forward_ms_isolated'], r['dominant_kernel']['ms_per_launch'], d['clocks'])
PY
tail -5 gpurun_out/r2_bench_s2.err
