#!/bin/bash
# tools/gpu_profile.sh TAG -- run on the B200 box (under gpurun).  Writes into gpurun_out/:
#   TAG_launches.csv      launch list of bench.py --quick (our kernels only), gpu__time_duration
#   TAG_<chain>.ncu-rep   one `--set full` capture per chain (second launch: warm caches/tables)
#   TAG_<chain>.log       prof_run.py's own (un-profiled-quality) output, for reference only
# Numbers printed under ncu are never bench values.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
KF='regex:rx_kernel|rx_dc_iir|rx_wbfm|tx_kernel|tx_wbfm|tx_idle'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 60 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --quick > $OUT/${TAG}_launches_bench.log 2>&1

cap() { # name kernel-regex launch-skip args...
    local name=$1 kre=$2 skip=$3
    shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c 1 -f \
        -o $OUT/${TAG}_$name python tools/prof_run.py "$@" > $OUT/${TAG}_$name.log 2>&1
}
for chain in ${CHAINS:-rx_mix rx_iir rx_fm rx_wbfm tx_am tx_fm tx_lsb tx_wbfm}; do
    case $chain in
    rx_mix)  cap rx_mix  'rx_kernel'  2 rx mix 1024 1.0 3 ;;
    rx_iir)  cap rx_iir  'rx_dc_iir'  2 rx mix 1024 1.0 3 ;;
    rx_fm)   cap rx_fm   'rx_kernel'  2 rx fm 4096 0.5 3 ;;
    rx_wbfm) cap rx_wbfm 'rx_wbfm_kernel'  2 rx wbfm 3996 0.25 3 ;; # 27 x 148 streams: one untiled launch per call
    tx_am)   cap tx_am   'tx_kernel'  2 tx am 4096 0.25 3 ;;
    tx_fm)   cap tx_fm   'tx_kernel'  2 tx fm 4096 0.25 3 ;;
    tx_lsb)  cap tx_lsb  'tx_kernel'  2 tx lsb 4096 0.25 3 ;;
    tx_wbfm) cap tx_wbfm 'tx_wbfm'    2 tx wbfm 4096 0.25 3 ;;
    esac
done
ls -la $OUT | tail -20
