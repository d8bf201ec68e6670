#!/usr/bin/env python
"""The squelched AM call of bench.py (run_next_rows) a few times, for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/sq_prof.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hackrfdiags_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_samples = int(0.5 * bench.FS) // 131072 * 131072
stream = torch.cuda.current_stream().cuda_stream
b, iq, pcm, _ = bench.make_rx_batch(torch, capi, dev, [(1, n_streams)], n_samples, seed=17)
for k in range(n_samples // 131072):
    if (k // 2) % 2 == 1:
        iq[:, k * 262144:(k + 1) * 262144].div_(32, rounding_mode="floor")
b.set_param(capi.PARAM_SQUELCH_THRESHOLD, -40.0)
call = lambda: b.rx_device(iq.data_ptr(), iq.shape[1], iq.stride(0), pcm.data_ptr(), pcm.stride(0), capi.ENTRY_2048K, stream)
ms, _ = bench.time_calls(torch, [call], 4, 2)
print(f"squelched AM, {n_streams} streams x {n_samples / bench.FS:.3f} s: {ms:.3f} ms per call, open fraction {b.squelch_report()[1].mean():.3f}")
