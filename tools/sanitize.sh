#!/bin/bash
# tools/sanitize.sh -- on the GPU box: compute-sanitizer over the GPU parity tests (every kernel kind, both entries,
# tiled and streaming calls, the WBFM retry / re-run paths, squelch, adapters, the sharded entry; the full-size
# property tests are left out).  Logs into gpurun_out/san_<tool>.log; every run under `timeout`.
OUT=gpurun_out
mkdir -p $OUT
run() { # tool seconds pytest-args...
    local tool=$1 secs=$2
    shift 2
    timeout $secs compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
        python -m pytest "$@" -m gpu -q > $OUT/san_$tool.log 2>&1
    echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/san_$tool.log | tr '\n' ' ')"
}
run memcheck 300 tests/test_gpu_parity_small.py tests/test_gpu_squelch.py tests/test_gpu_signals.py tests/test_gpu_adapters.py tests/test_gpu_sharded.py
run racecheck 300 tests/test_gpu_parity_small.py tests/test_gpu_squelch.py -k "wbfm or mixed or squelch or tile"
run synccheck 150 tests/test_gpu_parity_small.py tests/test_gpu_signals.py -k "wbfm or mixed or tx or signals"
run initcheck 200 tests/test_gpu_parity_small.py tests/test_gpu_squelch.py
