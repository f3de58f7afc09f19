#!/bin/bash
# ncu --set full of the small SGEMMs of one train step (eager, so that every launch is a kernel of its own)
mkdir -p gpurun_out
IWAE_NO_GRAPH=1 timeout 300 python scripts/train_probe.py 3 > gpurun_out/train_probe.log 2>&1 && \
IWAE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f32_kernel -s 172 -c 20 -f -o gpurun_out/gemm \
    python scripts/train_probe.py 3 > gpurun_out/ncu_gemm.log 2>&1 ; echo "rc=$?"
