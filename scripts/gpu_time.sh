#!/bin/bash
# kernel-only timing of the fused evaluation (2368 images x k=5000, 5 reps; prints LLH and ms)
timeout 300 python scripts/prof_eval.py --images 2368 --reps 5 2>&1 | tail -n 3
