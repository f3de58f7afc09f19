#!/bin/bash
# One gpurun call: full bench, then the ncu launch list and one --set full capture of the fused kernel.
mkdir -p gpurun_out
echo "== full bench" ; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ; echo "rc=$?" ; tail -c 2500 gpurun_out/bench_full.json ; tail -n 3 gpurun_out/bench_full.err
echo "== ncu launch list"
timeout 300 python bench.py --images 1000 --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --images 1000 --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 ; echo "rc=$?"
echo "== ncu full capture of the fused kernel"
timeout 300 python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/plain_prof.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:iwae_eval_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval \
    python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?" ; tail -n 5 gpurun_out/ncu_full.log ; cat gpurun_out/plain_prof.log
