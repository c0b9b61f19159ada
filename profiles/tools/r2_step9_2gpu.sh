#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -x > gpurun_out/r2_pytest_train.log 2>&1
echo "train tests rc=$?"; tail -12 gpurun_out/r2_pytest_train.log | cut -c1-300
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -q > gpurun_out/r2_pytest_mgpu.log 2>&1
echo "mgpu tests rc=$?"; tail -15 gpurun_out/r2_pytest_mgpu.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
timeout 600 python profiles/tools/prof_train_dp.py > gpurun_out/r2_prof_train_1gpu.txt 2>&1
echo "1gpu rc=$?"; grep "== world" gpurun_out/r2_prof_train_1gpu.txt
timeout 600 $TR profiles/tools/prof_train_dp.py > gpurun_out/r2_prof_train_2gpu.txt 2>&1
echo "2gpu rc=$?"; grep "== world" gpurun_out/r2_prof_train_2gpu.txt
