#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time the variant builds (see exp_build.sh); every run under `timeout`
for n in 16 128 144; do echo "exp $n:"; HRD_LIB=build/exp/libhrd_b200_$n.so timeout 60 python tools/prof_run.py tx wbfm 4096 0.5 6 2>&1 | tail -1; done
