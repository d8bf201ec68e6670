#!/bin/bash
# tools/final_run.sh TAG [CHAINS] -- on the GPU box: the bench line of the build and the ncu captures that go with it
# (tools/gpu_profile.sh); afterwards, here: python tools/make_profiles.py TAG
TAG=${1:-r1}
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
CHAINS="${2:-rx_mix rx_iir rx_fm rx_wbfm tx_am tx_fm tx_lsb tx_wbfm}" bash tools/gpu_profile.sh $TAG > gpurun_out/${TAG}_profile.log 2>&1
tail -3 gpurun_out/${TAG}_bench.err
