#!/bin/bash
# tools/exp_build.sh N... -- variant builds with parts of a kernel switched off (-DHRD_EXP=N), for
# timing experiments only: build/exp/libhrd_b200_N.so, loaded with HRD_LIB=... (results are WRONG by design)
mkdir -p build/exp
cd hackrfdiags_b200/csrc
for n in "$@"; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
        -DHRD_EXP=$n -shared -o ../../build/exp/libhrd_b200_$n.so hrd_api.cu hrd_rx.cu hrd_tx.cu hrd_squelch.cu hrd_adapt.cc hrd_shard.cc &
done
wait
