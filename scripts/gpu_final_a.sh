#!/bin/bash
# Final evidence, part A: the whole GPU test suite, smoke, the default bench line and the reference arm.
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1 ; echo "rc=$?" ; tail -n 4 gpurun_out/test_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "rc=$?" ; tail -n 2 gpurun_out/smoke.log
echo "== full bench" ; timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ; echo "rc=$?" ; tail -n 2 gpurun_out/bench_full.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print('value %.4g  ms/step %.2f  tensor frac %.3f  e2e %.4g  train %.0f  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['train']['value'], d['clocks']))"
echo "== reference arm" ; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; echo "rc=$?"
echo "== 1-layer bench" ; timeout 300 python bench.py --layers 1 --no-cpu-baseline --no-train > gpurun_out/bench_1layer.json 2> gpurun_out/bench_1layer.err ; echo "rc=$?"
