# tests of the hand-written training step, the train bench line and the ncu launch list of two steps
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tnet_gpu.py tests/test_train_gpu.py -q --tb=short > gpurun_out/r2_tnet_g.log 2>&1; tail -4 gpurun_out/r2_tnet_g.log
timeout 300 python bench.py --workload train --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_train_bench_hand.err | tee gpurun_out/r2_train_bench_hand.json | cut -c1-330
XQ_TRAIN_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_train_ncu.log 2>&1
tail -1 gpurun_out/r2_train_ncu.log | cut -c1-150
