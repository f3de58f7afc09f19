#!/bin/bash
# Tests, smoke, the full bench line and the train launch list (no ncu capture of the eval kernel: unchanged).
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1 ; echo "rc=$?" ; tail -n 4 gpurun_out/test_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "rc=$?" ; tail -n 2 gpurun_out/smoke.log
echo "== full bench" ; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ; echo "rc=$?" ; tail -n 2 gpurun_out/bench_full.err
echo "== ncu launch list (train)"
timeout 300 python scripts/train_probe.py 20 > gpurun_out/train_probe.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/launches_train.csv \
    python scripts/train_probe.py 20 > gpurun_out/ncu_launches_train.log 2>&1 ; echo "rc=$?" ; cat gpurun_out/train_probe.log
echo "== val_step probe"; timeout 300 python scripts/val_probe.py > gpurun_out/val_probe.log 2>&1; cat gpurun_out/val_probe.log
