#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
  for f in 1 0; do XQ_NET_FORK=$f timeout 120 python profiles/tools/fwd_ab.py; done
done 2>&1 | tee gpurun_out/r2_fwd_ab.txt
XQ_NET_2CTA=0 timeout 120 python profiles/tools/fwd_ab.py 2>&1 | tee -a gpurun_out/r2_fwd_ab.txt
