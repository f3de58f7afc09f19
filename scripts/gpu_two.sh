#!/bin/bash
# Two-GPU validation: the 2-rank GPU tests and the bench at N = 2 launched like the driver does.
mkdir -p gpurun_out
echo "== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -n 4
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
      bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/scale_n2.json') if l.startswith('{')][-1])
print('value %.4g ms/step %.2f e2e %.4g llh %.5f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['test_llh']))
print('train', d['train']['value'], d['train'].get('allreduce_us'), d['train']['launches_per_step'])
PY
