#!/bin/bash
# Timelines of the default build with two and with one resident CTA per SM (per-16-column-group events included).
mkdir -p gpurun_out
IWAE_TC_TRACE=gpurun_out/trace.txt timeout 200 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "rc=$?"
IWAE_TC_GRID=148 IWAE_TC_TRACE=gpurun_out/trace_1cta.txt timeout 200 python scripts/prof_eval.py --images 592 --reps 1 >> gpurun_out/trace_run.log 2>&1; echo "rc=$?"
python scripts/trace_summary.py gpurun_out/trace.txt > gpurun_out/timeline.txt 2>&1
python scripts/trace_summary.py gpurun_out/trace_1cta.txt > gpurun_out/timeline_1cta.txt 2>&1
head -n 1 gpurun_out/timeline.txt gpurun_out/timeline_1cta.txt
timeout 200 python scripts/prof_eval.py --images 2368 --reps 4 2>&1 | tail -n 2
