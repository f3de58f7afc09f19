#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/val_launches.py > gpurun_out/val_launches.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_val.csv \
    python scripts/val_launches.py > gpurun_out/ncu_launches_val.log 2>&1 ; echo "rc=$?"
