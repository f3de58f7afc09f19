#!/bin/bash
# Train step with the backward recomputing the logits (default) against stored logits (IWAE_RECOMPUTE=0): parity tests
# of the training path, then steps/s of both (2,000 graph-replayed steps each, three repetitions).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tools.py tests/test_gpu_experiments.py tests/test_gpu_boundary.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/test_train.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/test_train.log
IWAE_RECOMPUTE=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "grad or traj or train" > gpurun_out/test_train_stored.log 2>&1; echo "tests (stored) rc=$?"; tail -n 1 gpurun_out/test_train_stored.log
for rep in 1 2 3; do
  for r in 1 0; do echo -n "IWAE_RECOMPUTE=$r: "; IWAE_RECOMPUTE=$r timeout 200 python scripts/train_probe.py 2000 | tail -n 1; done
done
