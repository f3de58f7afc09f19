#!/bin/bash
# A/B of the default build: eval parity tests, then kernel time on 2,368 images for each environment setting in $ENVS
# (semicolon-separated, e.g. "IWAE_TC_EARLY2=0;IWAE_TC_NO_EARLY=1") and for every variant library.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_hardening.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_eval.log 2>&1; echo "parity rc=$?"; tail -n 2 gpurun_out/test_eval.log
echo "== default"; timeout 200 python scripts/prof_eval.py --images ${IMAGES:-2368} --reps 4 2>&1 | tail -n 2
IFS=';' read -ra E <<< "$ENVS"
for e in "${E[@]}"; do [ -n "$e" ] || continue; echo "== $e"; env $e timeout 200 python scripts/prof_eval.py --images ${IMAGES:-2368} --reps 4 2>&1 | tail -n 1; done
for so in iwae_b200/lib/variants/*.so; do [ -f "$so" ] || continue; echo "== $so"; IWAE_B200_LIB=$PWD/$so timeout 200 python scripts/prof_eval.py --images ${IMAGES:-2368} --reps 4 2>&1 | tail -n 1; done
echo "== 1 layer"; timeout 200 python scripts/prof_eval.py --images ${IMAGES:-2368} --reps 3 --layers 1 2>&1 | tail -n 1
