set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?" >> gpurun_out/bench_default.err
XQ_BENCH_SIMS=4 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sims4.json 2> gpurun_out/bench_sims4.err
echo "sims4 rc=$?" >> gpurun_out/bench_sims4.err
XQ_BENCH_SIMS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mcts_select|mcts_expand|sp_after_root|sp_end_move|sp_new' -s 16 -c 12 -f -o gpurun_out/prof_r1_tree python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tree.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_tree.log
ls -la gpurun_out
