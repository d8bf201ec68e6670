set -x
timeout 400 python bench.py > gpurun_out/r1w_bench.json 2> gpurun_out/r1w_bench.err; echo bench rc=$?
CHAINS="tx_am tx_fm tx_lsb" bash tools/gpu_profile.sh r1w > gpurun_out/r1w_profile.log 2>&1
tail -3 gpurun_out/r1w_bench.err
