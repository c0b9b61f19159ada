set -x
timeout 400 python -m pytest tests/test_tnet_gpu.py tests/test_train_gpu.py -q --tb=short > gpurun_out/r2_tnet_g.log 2>&1; tail -5 gpurun_out/r2_tnet_g.log
timeout 300 python bench.py --workload train --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_train_bench_hand.err | tee gpurun_out/r2_train_bench_hand.json | cut -c1-330
export XQ_TRAIN_GRAPH=0
B="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tg_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_tg_conv $B > gpurun_out/ncu_tg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:twg_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_twg_conv $B > gpurun_out/ncu_twg.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:tn_bn -s 0 -c 2 -f -o gpurun_out/prof_r2_bn_fwd $B > gpurun_out/ncu_bnf.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:tn_bn_bwd -s 4 -c 2 -f -o gpurun_out/prof_r2_bn_bwd $B > gpurun_out/ncu_bnb.log 2>&1
ls -la gpurun_out/*.ncu-rep
