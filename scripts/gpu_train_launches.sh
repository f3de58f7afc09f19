#!/bin/bash
# Launch list (per-kernel durations) of a few train steps; plain run first.
mkdir -p gpurun_out
IWAE_CHAIN=1 timeout 300 python scripts/train_probe.py 20 > gpurun_out/train_probe.log 2>&1 && \
IWAE_CHAIN=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/launches_train.csv \
    python scripts/train_probe.py 20 > gpurun_out/ncu_launches_train.log 2>&1 ; echo "rc=$?"
