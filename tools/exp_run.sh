#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time variant builds against the product build; every run under `timeout`
run() { HRD_LIB=$1 timeout 60 python tools/prof_run.py $2 $3 4096 0.5 8 2>&1 | tail -1; }
echo "product:"; run "" tx wbfm
echo "stages 1-4 on every fourth step only (upper bound of batching them):"; run build/exp/libhrd_b200_256.so tx wbfm
