#!/bin/bash
# Final evidence, part B: ncu launch lists (eval, train), one ncu --set full capture of the fused kernel, its DRAM / L2
# traffic on the full launch, the cycle timeline (two and one resident CTA per SM), the val_step probe.
mkdir -p gpurun_out
echo "== ncu launch list (eval)"
timeout 300 python bench.py --images 1000 --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --images 1000 --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 ; echo "rc=$?"
echo "== ncu launch list (train)"
timeout 300 python scripts/train_probe.py 20 > gpurun_out/train_probe.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/launches_train.csv \
    python scripts/train_probe.py 20 > gpurun_out/ncu_launches_train.log 2>&1 ; echo "rc=$?" ; cat gpurun_out/train_probe.log
echo "== ncu full capture"
timeout 300 python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iwae_eval_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval \
    python scripts/prof_eval.py --images 592 --reps 2 > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?" ; cat gpurun_out/plain_prof.log
echo "== timelines"
IWAE_TC_TRACE=gpurun_out/trace.txt timeout 200 python scripts/prof_eval.py --images 592 --reps 1 > gpurun_out/trace_run.log 2>&1; echo "rc=$?"
IWAE_TC_GRID=148 IWAE_TC_TRACE=gpurun_out/trace_1cta.txt timeout 200 python scripts/prof_eval.py --images 592 --reps 1 >> gpurun_out/trace_run.log 2>&1; echo "rc=$?"
echo "== ncu DRAM / L2 traffic of the full 10,000-image launch"
bash scripts/gpu_traffic.sh > gpurun_out/traffic_run.log 2>&1; tail -n 4 gpurun_out/traffic_run.log
echo "== val_step probe"; timeout 300 python scripts/val_probe.py > gpurun_out/val_probe.log 2>&1; cat gpurun_out/val_probe.log
