# round-end gates on one B200: the whole GPU suite, smoke(), the default bench line (with the f1 block) and the reference arm
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
tail -5 gpurun_out/r2_pytest_gpu_final.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_smoke_final.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke_final.log
tail -3 gpurun_out/r2_smoke_final.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default_final.err
tail -4 gpurun_out/r2_bench_default_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_default_final.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
print(json.dumps(d["extra"]["f1_train_step"])[:900])
PY
