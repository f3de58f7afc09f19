#!/bin/bash
mkdir -p gpurun_out
IWAE_TC_TRACE=gpurun_out/trace.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/trace_run.log; wc -l gpurun_out/trace.txt
timeout 300 python scripts/prof_eval.py --images 592 --reps 3
