#!/usr/bin/env python
"""Record event traces of the UNMODIFIED BasebandDataProcessor PCM ring (oracle/_ref, dev container only) into
tests/golden/ring_traces.npz, so that tests/test_adapters_cpu.py can pin hrd_pcm_ring_* where /root/reference is absent."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from test_adapters_cpu import make_events, run_reference  # noqa: E402

out = {"n": np.int64(6)}
for k in range(6):
    ev = make_events(100 + k, n=300)
    o, st = run_reference(ev)
    out[f"events_{k}"] = np.array(ev)
    out[f"out_{k}"] = o
    out[f"stats_{k}"] = st
np.savez_compressed(os.path.join(HERE, "ring_traces.npz"), **out)
print("wrote ring_traces.npz", {k: v.shape for k, v in out.items() if k.startswith("out")})
