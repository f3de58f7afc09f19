#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_eval.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_val.log 2>&1 ; echo "rc=$?" ; tail -n 4 gpurun_out/test_val.log
timeout 300 python scripts/val_probe.py 2>&1 | tail -n 5
