#!/bin/bash
# ONE compute-sanitizer tool per gpurun call (B200_PROFILING.md) on a small fused evaluation:
#   gpurun -- 'bash scripts/gpu_sanitize.sh memcheck'   |  racecheck | synccheck
# The plain command must have exited 0 first (same call, directly before).
TOOL=${1:-memcheck}
mkdir -p gpurun_out
CMD="python scripts/prof_eval.py --images ${IMAGES:-6} --k ${K:-300} --reps 1"
timeout 300 $CMD > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 --log-file gpurun_out/sanitizer_$TOOL.log $CMD > gpurun_out/sanitize_run_$TOOL.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/sanitize_plain.log; tail -n 12 gpurun_out/sanitizer_$TOOL.log
