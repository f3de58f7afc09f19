#!/bin/bash
# A/B of kernel flags on the full-size bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -n 2
for f in 0 1; do
  IWAE_TC_FLAGS=$f timeout 600 python bench.py --no-train --no-cpu-baseline --steps 3 > gpurun_out/bench_f$f.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_f$f.json')); print('flags=$f value %.4g  ms/step %.2f  tensor frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
