set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_movegen_gpu.py -x -q ) > gpurun_out/pytest_movegen.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_movegen.log
for impl in warp thread; do
  XQ_MOVEGEN_IMPL=$impl timeout 300 python bench.py --workload movegen --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_movegen_$impl.json 2> gpurun_out/bench_movegen_$impl.err
  echo "bench $impl rc=$?" >> gpurun_out/bench_movegen_$impl.err
done
XQ_MOVEGEN_IMPL=thread timeout 600 ncu --set full --clock-control none --import-source on -k regex:movegen_tpb -s 3 -c 1 -f -o gpurun_out/prof_r1_movegen_tpb python bench.py --workload movegen --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tpb.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_tpb.log
ls -la gpurun_out
