#!/bin/bash
# Iteration call: eval tests, full bench, one ncu --set full capture of the fused kernel.
mkdir -p gpurun_out
echo "== eval tests" ; timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q -s --timeout 300 -p no:cacheprovider > gpurun_out/test_eval.log 2>&1 ; echo "rc=$?" ; grep -E "tc parity|k=5000|passed|failed|Error" gpurun_out/test_eval.log | tail -n 30
echo "== full bench" ; timeout 900 python bench.py --no-train --cpu-seconds 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ; echo "rc=$?" ; tail -c 2600 gpurun_out/bench_full.json ; tail -n 3 gpurun_out/bench_full.err
echo "== ncu full capture of the fused kernel"
timeout 300 python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/plain_prof.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:iwae_eval_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval \
    python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?" ; tail -n 3 gpurun_out/ncu_full.log ; cat gpurun_out/plain_prof.log
