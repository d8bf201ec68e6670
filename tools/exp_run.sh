#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time variant builds against the product build; every run under `timeout`
run() { HRD_LIB=$1 timeout 60 python tools/prof_run.py $2 $3 4096 0.5 8 2>&1 | tail -1; }
echo "product (L2 ahead 4):"; for c in "rx am" "rx fm" "rx wbfm"; do run "" $c; done
for a in 0 2; do echo "L2 ahead $a:"; for c in "rx am" "rx fm" "rx wbfm"; do run build/exp/libhrd_b200_A$a.so $c; done; done
