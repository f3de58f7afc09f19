#!/bin/bash
# Train-step experiment: gradient parity + boundary tests, then the step time of the multi-kernel graph and of the megakernel.
mkdir -p gpurun_out
echo "== tests" ; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_hardening.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/test_train.log 2>&1 ; echo "rc=$?" ; tail -n 3 gpurun_out/test_train.log
echo "== multi-kernel" ; timeout 120 python scripts/train_probe.py 300 2>&1 | tail -n 4
echo "== mega" ; IWAE_MEGA=1 timeout 120 python scripts/train_probe.py 300 2>&1 | tail -n 4
timeout 200 python scripts/mega_trace.py > gpurun_out/mega_trace.log 2>&1
