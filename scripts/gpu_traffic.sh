#!/bin/bash
# DRAM traffic of the fused kernel on the bench's own launch (10,000 images x k=5000): one ncu capture.
mkdir -p gpurun_out
timeout 300 python scripts/prof_eval.py --images 10000 --reps 2 > gpurun_out/plain_traffic.log 2>&1 && \
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:iwae_eval_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_full.csv \
    python scripts/prof_eval.py --images 10000 --reps 2 > gpurun_out/ncu_traffic.log 2>&1 ; echo "rc=$?"; cat gpurun_out/traffic_full.csv | tail -n 9; cat gpurun_out/plain_traffic.log
