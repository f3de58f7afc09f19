#!/bin/bash
# Time every variant library under iwae_b200/lib/variants (and the default build) on the fused evaluation.
for so in iwae_b200/lib/libiwae_b200.so iwae_b200/lib/variants/*.so; do
  [ -f "$so" ] || continue
  echo "== $so"
  IWAE_B200_LIB=$PWD/$so timeout 300 python scripts/prof_eval.py --images ${IMAGES:-2368} --reps 4 2>&1 | tail -n 2
done
