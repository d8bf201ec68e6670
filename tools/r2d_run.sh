#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== timings"
for n in 256 2048 4096; do
  echo "-- streams $n"; timeout 120 python tools/prof_run.py tx wbfm $n 0.5 6 2>&1 | tail -1
done
timeout 120 python tools/prof_run.py tx fm 4096 0.5 6 2>&1 | tail -1
CHAINS="tx_wbfm" bash tools/gpu_profile.sh r2d > gpurun_out/r2d_profile.log 2>&1
