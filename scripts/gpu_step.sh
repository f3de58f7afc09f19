#!/bin/bash
# One experiment step of the fused kernel: eval parity tests, the Philox restatement test, kernel timing, timeline.
mkdir -p gpurun_out
echo "== eval tests" ; timeout 900 python -m pytest tests/test_gpu_eval.py "tests/test_gpu_parity.py::test_philox_noise_matches_restatement" -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_eval.log 2>&1 ; echo "rc=$?" ; tail -n 6 gpurun_out/test_eval.log
echo "== timing" ; timeout 300 python scripts/prof_eval.py --images 2368 --reps 4 2>&1 | tail -n 3
if [ -n "$TRACE" ]; then IWAE_TC_TRACE=gpurun_out/trace.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "trace rc=$?"; fi
