#!/bin/bash
# One gpurun call: GPU tests (fp32 parity first, then the tensor-core eval), smoke, a short bench.
# Everything is logged under gpurun_out/ ; no step aborts the following ones.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== parity (fp32 path)" ; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_parity.log 2>&1 ; echo "rc=$?" ; tail -n 15 gpurun_out/test_parity.log
echo "== eval (tcgen05 path)" ; timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q -s --timeout 300 -p no:cacheprovider > gpurun_out/test_eval.log 2>&1 ; echo "rc=$?" ; tail -n 25 gpurun_out/test_eval.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "rc=$?" ; tail -n 8 gpurun_out/smoke.log
echo "== bench (reduced: 1000 images)" ; timeout 600 python bench.py --images 1000 --steps 2 --warmup 1 --cpu-seconds 5 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err ; echo "rc=$?" ; tail -c 3000 gpurun_out/bench_small.json ; tail -n 5 gpurun_out/bench_small.err
