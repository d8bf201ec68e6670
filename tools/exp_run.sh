#!/bin/bash
# tools/exp_run.sh -- on the GPU box: time variant builds (tools/exp_build.sh, or any -D build under build/exp/)
# against the product build; every run under `timeout`.
#   bash tools/exp_run.sh "tx fm" "tx lsb" -- build/exp/libhrd_b200_A.so build/exp/libhrd_b200_B.so
# Results of variant builds with parts switched off (HRD_EXP) are WRONG by design: timings only.
chains=(); libs=("")
while [ $# -gt 0 ] && [ "$1" != "--" ]; do chains+=("$1"); shift; done
[ "$1" == "--" ] && shift
libs+=("$@")
for lib in "${libs[@]}"; do
    echo "== ${lib:-product (hackrfdiags_b200/libhrd_b200.so)}"
    for c in "${chains[@]}"; do
        HRD_LIB=${lib:+$PWD/$lib} timeout 60 python tools/prof_run.py $c 4096 0.5 8 2>&1 | tail -1
    done
done
