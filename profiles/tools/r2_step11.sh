#!/bin/bash
# final-kernel evidence: launch list and ncu --set full of one forward, after the same command ran clean without ncu
mkdir -p gpurun_out
XQ_BENCH_SIMS=4 XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_sims4.json 2> gpurun_out/r2_bench_sims4.err
echo "sims4 rc=$?"
XQ_BENCH_SIMS=4 XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2_selfplay_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
XQ_BENCH_SIMS=4 XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv2_kernel|conv_kernel|fc_kernel|value_head' -s 60 -c 18 -f -o gpurun_out/prof_r2_net_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_net.log 2>&1
echo "ncu net rc=$?"
XQ_BENCH_SIMS=4 XQ_BENCH_NO_CONFIGS3=1 XQ_BENCH_NO_API_E2E=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mcts_select_multi|mcts_expand_backup_multi|sp_after_root|sp_end_move|sp_root_begin' -s 10 -c 10 -f -o gpurun_out/prof_r2_tree python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_tree.log 2>&1
echo "ncu tree rc=$?"
