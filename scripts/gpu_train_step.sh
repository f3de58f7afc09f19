#!/bin/bash
# Train-step check: gradient parity / boundary / hardening / experiment tests, then the step time (graph of kernels, megakernel).
mkdir -p gpurun_out
echo "== tests" ; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_tools.py tests/test_gpu_experiments.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_train.log 2>&1 ; echo "rc=$?" ; tail -n 3 gpurun_out/test_train.log
echo "== multi-kernel" ; timeout 120 python scripts/train_probe.py 300 2>&1 | tail -n 4
echo "== mega" ; IWAE_MEGA=1 timeout 120 python scripts/train_probe.py 300 2>&1 | tail -n 4
