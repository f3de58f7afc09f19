#!/bin/bash
# Scaling run on one multi-GPU box: N = 1, 2, 4, 8 (as many as visible), exactly as the driver launches it.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus visible: $NG"
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scale_n$N.json') if l.startswith('{')][-1])
    print('  value %.4g  ms/step %.2f  e2e %.4g  llh %.5f  launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['test_llh'], d['gpu_launches']))
except Exception as e:
    print('  parse failed', e); print(open('gpurun_out/scale_n$N.err').read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -n 1 | cut -c1-300
