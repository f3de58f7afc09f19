#!/bin/bash
# Train-step experiment: gradient parity + boundary tests, then the step time with / without the fused Dense chains.
mkdir -p gpurun_out
echo "== tests" ; IWAE_CHAIN=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_hardening.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_train.log 2>&1 ; echo "rc=$?" ; tail -n 8 gpurun_out/test_train.log
echo "== chain" ; IWAE_CHAIN=1 timeout 300 python scripts/train_probe.py 300 2>&1 | tail -n 4
echo "== no chain" ; timeout 300 python scripts/train_probe.py 300 2>&1 | tail -n 4
