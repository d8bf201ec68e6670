#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time variant builds against the product build; every run under `timeout`
run() { HRD_LIB=$1 timeout 60 python tools/prof_run.py $2 $3 4096 0.5 8 2>&1 | tail -1; }
echo "product:"; for c in "rx am" "rx fm" "rx wbfm"; do run "" $c; done
echo "RX_DEPTH=1:"; for c in "rx am" "rx fm"; do run build/exp/libhrd_b200_RX1.so $c; done
echo "WB_DEPTH=1:"; run build/exp/libhrd_b200_WB1.so rx wbfm
