#!/bin/bash
# round 2, step 3: launch list + ncu --set full of the forward kernels, reference arm on the box, multi-leaf timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rng_gpu.py -x -q 2>&1 | tail -3
XQ_BENCH_SIMS=4 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_sims4.json 2> gpurun_out/r2_bench_sims4.err
echo "sims4 rc=$?"
XQ_BENCH_SIMS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2_selfplay_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
XQ_BENCH_SIMS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_kernel|fc_kernel|value_head' -s 60 -c 18 -f -o gpurun_out/prof_r2_net python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_net.log 2>&1
echo "ncu net rc=$?"
# small game counts: K = 1 against the multi-leaf mode (same total leaves per step)
for cfg in "1024 1" "1024 4" "256 16"; do
  set -- $cfg
  XQ_BENCH_GAMES=$1 XQ_BENCH_LEAVES=$2 XQ_BENCH_SIMS=800 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_g$1_k$2.json 2> gpurun_out/r2_bench_g$1_k$2.err
  echo "games $1 leaves $2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_g$1_k$2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['forward_ms_isolated'])"
done
# the reference arm on this box's host cores (short)
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err
echo "ref arm rc=$?"; cut -c1-1200 gpurun_out/r2_ref_arm.json
