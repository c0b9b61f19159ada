#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
timeout 600 python profiles/tools/prof_train_dp.py > gpurun_out/r2_prof_train_1gpu.txt 2>&1
echo "1gpu rc=$?"; grep "== world" gpurun_out/r2_prof_train_1gpu.txt
timeout 600 $TR profiles/tools/prof_train_dp.py > gpurun_out/r2_prof_train_2gpu.txt 2>&1
echo "2gpu rc=$?"; grep "== world" gpurun_out/r2_prof_train_2gpu.txt
