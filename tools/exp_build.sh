#!/bin/bash
# tools/exp_build.sh NAME "FLAGS" [NAME "FLAGS" ...] -- variant builds for timing experiments:
# build/exp/libhrd_b200_NAME.so compiled with the extra FLAGS (e.g. "-DHRD_RX_WB_NBUF=3"), loaded with HRD_LIB=...
# (variants that switch parts of a kernel off with -DHRD_EXP=N give WRONG results by design: timings only)
mkdir -p build/exp
cd hackrfdiags_b200/csrc
while [ $# -ge 2 ]; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
        $2 -shared -o ../../build/exp/libhrd_b200_$1.so hrd_api.cu hrd_rx.cu hrd_tx.cu hrd_squelch.cu hrd_adapt.cc hrd_shard.cc &
    shift 2
done
wait
