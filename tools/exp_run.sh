#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time variant builds against the product build; every run under `timeout`
run() { HRD_LIB=$1 timeout 60 python tools/prof_run.py $2 $3 4096 0.5 8 2>&1 | tail -1; }
echo "product (L2 prefetch 4 KiB ahead):"; for c in "rx am" "rx fm"; do run "" $c; done
for a in 2 8; do echo "L2 prefetch $a KiB ahead:"; for c in "rx am" "rx fm"; do run build/exp/libhrd_b200_A$a.so $c; done; done
