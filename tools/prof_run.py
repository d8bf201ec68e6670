#!/usr/bin/env python
"""Tiny driver for ncu: launches one chain a few times, device-resident, no timing claims.
   python tools/prof_run.py rx am 4096 0.25 [reps]      python tools/prof_run.py tx fm 4096 0.25"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hackrfdiags_b200 import capi  # noqa: E402

kind, mode_name, n, secs = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
mode = {"am": 1, "fm": 2, "wbfm": 3, "lsb": 4, "usb": 5, "mix": 0, "mix5": 0}[mode_name]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n_samples = int(secs * bench.FS) // 8192 * 8192
if kind == "rx":
    if mode_name == "mix":  # BASELINE config 2: half AM, quarter LSB, quarter USB
        groups = [(1, n // 2), (4, n // 4), (5, n - n // 2 - n // 4)]
    elif mode_name == "mix5":  # the bench headline (config 5): AM, NBFM, WBFM, LSB, USB
        from hackrfdiags_b200 import shard
        groups = shard.mode_groups(shard.shard_modes(shard.mixed_mode_plan(n, bench.MIX), 1, 0))
    else:
        groups = [(mode, n)]
    r, keep = bench.bench_rx_modes(torch, capi, dev, groups, n_samples, reps, 1, seed=5)
    ms = r["ms"]
    print(f"tile kernel {r['kernel_ms']:.3f} ms, tail {r['tail_ms']:.3f} ms, wbfm fallbacks {r['wbfm_fallbacks']} serial {r['wbfm_serial']} in {reps + 1} calls")
else:
    ms = bench.bench_tx_mode(torch, capi, dev, mode, n, n_samples // 256, reps, 1, seed=5)
torch.cuda.synchronize()
sps = n * n_samples / (ms * 1e-3)
print(f"{kind} {mode_name} streams={n} secs={secs}: {ms:.3f} ms/launch, {sps / 1e6:.0f} MS/s "
      f"({sps * 2.0078125 / 1e9:.0f} GB/s algorithmic)")
