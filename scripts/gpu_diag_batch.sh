#!/bin/bash
# One GPU call: time the default build and every variant library (scripts/build_variants.sh) on the fused evaluation,
# the eval parity tests on the candidate variant ($CAND), and its cycle timelines with two and with one resident CTA per SM.
mkdir -p gpurun_out
CAND=${CAND:-conv}
V=$PWD/iwae_b200/lib/variants/libiwae_$CAND.so
echo "== parity ($CAND)"; IWAE_B200_LIB=$V timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_eval_$CAND.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/test_eval_$CAND.log
for so in iwae_b200/lib/libiwae_b200.so iwae_b200/lib/variants/*.so; do
  echo "== $so"
  IWAE_B200_LIB=$PWD/$so timeout 200 python scripts/prof_eval.py --images ${IMAGES:-2368} --reps 4 2>&1 | tail -n 2
done
echo "== one CTA per SM (grid 148)"
for so in iwae_b200/lib/libiwae_b200.so $V; do
  IWAE_TC_GRID=148 IWAE_B200_LIB=$PWD/$so timeout 200 python scripts/prof_eval.py --images 1184 --reps 3 2>&1 | tail -n 1
done
echo "== traces"
IWAE_B200_LIB=$V IWAE_TC_TRACE=gpurun_out/trace_${CAND}.txt timeout 200 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "rc=$?"
IWAE_B200_LIB=$V IWAE_TC_GRID=148 IWAE_TC_TRACE=gpurun_out/trace_${CAND}_1cta.txt timeout 200 python scripts/prof_eval.py --images 592 --reps 1 >> gpurun_out/trace_run.log 2>&1; echo "rc=$?"
python scripts/trace_summary.py gpurun_out/trace_${CAND}.txt > gpurun_out/timeline_${CAND}.txt 2>&1
python scripts/trace_summary.py gpurun_out/trace_${CAND}_1cta.txt > gpurun_out/timeline_${CAND}_1cta.txt 2>&1
head -n 1 gpurun_out/timeline_${CAND}.txt gpurun_out/timeline_${CAND}_1cta.txt
