#!/bin/bash
# tensor-core encoder pre-pass: its parity test, the eval parity tests, kernel / step time with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "encoder" -s > gpurun_out/test_enc.log 2>&1; echo "encoder test rc=$?"; grep -E "tc encoder|passed|failed|Error|error" gpurun_out/test_enc.log | tail -n 14
timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_hardening.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_eval.log 2>&1; echo "eval tests rc=$?"; tail -n 2 gpurun_out/test_eval.log
for e in 1 0; do
  echo "== IWAE_TC_ENCODER=$e"
  IWAE_TC_ENCODER=$e timeout 300 python bench.py --no-train --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  kernel %.3f  e2e ms %.3f  llh %.6f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['test_llh']))"
  IWAE_TC_ENCODER=$e timeout 300 python scripts/val_probe.py 2>&1 | grep auto
done
