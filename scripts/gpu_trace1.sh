#!/bin/bash
# timeline of one CTA with ONE CTA per SM (no inter-CTA contention) next to the normal two-per-SM run
mkdir -p gpurun_out
IWAE_TC_GRID=148 IWAE_TC_TRACE=gpurun_out/trace_1cta.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace1_run.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/trace1_run.log
IWAE_TC_TRACE=gpurun_out/trace.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/trace_run.log
