#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/plain_prof.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:iwae_eval_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval \
    python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?" ; tail -n 3 gpurun_out/ncu_full.log ; cat gpurun_out/plain_prof.log
