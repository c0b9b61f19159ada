set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?" >> gpurun_out/bench_default.err
( time timeout 600 python bench.py --workload movegen ) > gpurun_out/bench_movegen.json 2> gpurun_out/bench_movegen.err
echo "bench movegen rc=$?" >> gpurun_out/bench_movegen.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_movegen_launches_tpb.csv python bench.py --workload movegen --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_movegen_launches.log 2>&1
echo "ncu movegen rc=$?" >> gpurun_out/ncu_movegen_launches.log
XQ_BENCH_SIMS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r1_selfplay_launches_v7.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_selfplay_launches.log 2>&1
echo "ncu selfplay rc=$?" >> gpurun_out/ncu_selfplay_launches.log
ls -la gpurun_out
