#!/bin/bash
# tools/quick_modes.sh [chains...] -- on the GPU box: un-profiled timing of single chains (prof_run.py), 4096 x 0.5 s
for c in ${@:-rx_wbfm tx_wbfm tx_fm tx_lsb rx_fm rx_am}; do
    kind=${c%%_*}; mode=${c#*_}
    python tools/prof_run.py $kind $mode 4096 0.5 10 2>&1 | tail -1
done
