#!/bin/bash
# kernel-only timing (no trace), then the one-CTA-per-SM and the normal timeline
mkdir -p gpurun_out
timeout 300 python scripts/prof_eval.py --images 2368 --reps 5 2>&1 | tail -n 2
IWAE_TC_GRID=148 IWAE_TC_TRACE=gpurun_out/trace_1cta.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace1_run.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/trace1_run.log
IWAE_TC_TRACE=gpurun_out/trace.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/trace_run.log
