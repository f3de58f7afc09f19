#!/bin/bash
# Quick iteration: eval parity tests, full-size bench (no CPU legs), timeline trace.
mkdir -p gpurun_out
echo "== eval tests" ; timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q -s --timeout 300 -p no:cacheprovider > gpurun_out/test_eval.log 2>&1 ; echo "rc=$?" ; grep -E "iwae2 scale=1.0|k=5000|passed|failed|Error" gpurun_out/test_eval.log | tail -n 8
echo "== full bench" ; timeout 900 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ; echo "rc=$?" ; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('value %.4g  ms/step %.2f  tensor frac %.3f  e2e %.4g  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']))"; tail -n 3 gpurun_out/bench_full.err
IWAE_TC_TRACE=gpurun_out/trace.txt timeout 300 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "trace rc=$?"
