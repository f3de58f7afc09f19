#!/bin/bash
mkdir -p gpurun_out
IWAE_CHAIN=1 timeout 300 python scripts/train_probe.py 5 > gpurun_out/train_probe.log 2>&1 && \
IWAE_CHAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_chain -s 8 -c 4 -f -o gpurun_out/chain \
    python scripts/train_probe.py 5 > gpurun_out/ncu_chain.log 2>&1 ; echo "rc=$?"
