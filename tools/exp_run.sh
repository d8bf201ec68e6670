#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time variant builds against the product build; every run under `timeout`
run() { HRD_LIB=$1 timeout 60 python tools/prof_run.py $2 $3 4096 0.5 8 2>&1 | tail -1; }
echo "product (8 CTAs/SM, 64 regs):"; for c in "rx am" "rx fm" "rx lsb"; do run "" $c; done
for a in 9 10; do echo "min CTAs $a:"; for c in "rx am" "rx fm" "rx lsb"; do run build/exp/libhrd_b200_C$a.so $c; done; done
