#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference, compiled by oracle/Makefile
into oracle/_ref/libhrd_ref.so (the reference repo itself stores no expected outputs for
this path, SURVEY.md section 8c, so these vectors are what pins the oracle and the CUDA
path when /root/reference is not around, e.g. on the GPU box).

    python tests/golden/make_golden.py          # needs /root/reference (dev container only)

Every vector stores the INPUT bytes next to the reference's output, so the fixtures do not
depend on numpy's random generator staying stable.  Reference calls made per vector:

  rx2048k_<mode>   IqDataProcessor::setDemodulatorMode + acceptIqData, two calls (state carries)
  rx256k_<mode>    <X>Demodulator::acceptIqData, two calls
  fe               IqDataProcessor::reduceSampleRate + upconvertByFsOver4
  tx_<mode>        <X>Modulator::acceptData, two calls (64 + 32 PCM samples)
  squelch_<mode>   IqDataProcessor::setSignalDetectThreshold(-40) + acceptIqData in 12 calls of 8192 bytes whose
                   level crosses the threshold; PCM plus what the reference's magnitude / state callbacks reported
  sig_<head>       signals/<head> < pcm | signals/interpolateSignal (the programs themselves, oracle/_ref/)
  tables           quantised taps as the constructors built them, sha256 of the NCO tables
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from cpu_checkers import Ref, TAPS, build_checkers, DEMOD_OF_MODE  # noqa: E402
from hackrfdiags_b200 import synth  # noqa: E402

MODES = {"am": 1, "fm": 2, "wbfm": 3, "lsb": 4, "usb": 5}
CONFIG = 90  # seed family reserved for the golden vectors


def main():
    build_checkers()
    ref = Ref()
    out = {}

    # ---- Rx, 2.048 MS/s entry: a signal stream and a full-range noise stream per mode ----
    for name, mode in MODES.items():
        for tag, edge, n in (("sig", None, 24576), ("noise", "noise", 8192)):
            iq = synth.rx_stream(mode, n, stream=mode, config=CONFIG, edge=edge)
            h = ref.rx_new()
            ref.rx_set_mode(h, mode)
            cut = 2 * (n // 2 // 256 * 256)
            pcm = np.concatenate([ref.rx_accept_2048k(h, iq[:cut]), ref.rx_accept_2048k(h, iq[cut:])])
            ref.rx_free(h)
            assert pcm.size == n // 256
            out[f"rx2048k_{name}_{tag}_iq"] = iq
            out[f"rx2048k_{name}_{tag}_cut"] = np.int64(cut)
            out[f"rx2048k_{name}_{tag}_pcm"] = pcm

    # ---- Rx, 256 kS/s entry, default and non-default gain --------------------------------
    for name, mode in MODES.items():
        n = 4096
        iq = synth.rx_stream(mode, n, stream=10 + mode, config=CONFIG, entry="256k")
        for tag, gain in (("g0", None), ("g1", 1234.5)):
            h = ref.rx_new()
            ref.rx_set_mode(h, mode)
            if gain is not None:
                ref.rx_set_gain(h, DEMOD_OF_MODE[mode], gain)
            pcm = np.concatenate([ref.rx_accept_256k(h, iq[:2 * 1024]), ref.rx_accept_256k(h, iq[2 * 1024:])])
            ref.rx_free(h)
            assert pcm.size == n // 32
            out[f"rx256k_{name}_{tag}_pcm"] = pcm
        out[f"rx256k_{name}_iq"] = iq

    # ---- front end alone (the 256 kS/s stream the reference can dump over UDP) -----------
    for tag, edge in (("noise", "noise"), ("max", "max"), ("min", "min")):
        iq = synth.rx_stream(2, 4096, stream=20, config=CONFIG, edge=edge)
        h = ref.rx_new()
        out[f"fe_{tag}_iq"] = iq
        out[f"fe_{tag}_out"] = ref.rx_front_end(h, iq)
        ref.rx_free(h)

    # ---- Tx ----------------------------------------------------------------------------
    for name, mode in MODES.items():
        for kind in ("sine", "noise"):
            pcm = synth.tx_stream(96, stream=30 + mode, config=CONFIG, kind=kind)
            h = ref.tx_new()
            iq = np.concatenate([ref.tx_accept(h, mode, pcm[:64]), ref.tx_accept(h, mode, pcm[64:])])
            ref.tx_free(h)
            out[f"tx_{name}_{kind}_pcm"] = pcm
            out[f"tx_{name}_{kind}_iq"] = iq

    # ---- squelch gate: short reference calls whose level crosses the threshold -------------
    for name in ("am", "fm", "usb"):
        mode, blk = MODES[name], 8192
        rng = np.random.default_rng(synth.stream_seed(CONFIG, 40 + mode))
        levels = [90, 2, 2, 50, 50, 1, 20, 120, 0, 0, 40, 3]
        n = 12 * blk // 2
        t = np.arange(n) / 2.048e6
        amp = np.repeat(np.array(levels, dtype=np.float64), blk // 2)
        z = amp * np.exp(2j * np.pi * (-64000.0 * t + 0.4 * np.sin(2 * np.pi * 1000 * t)))
        iq = np.empty(2 * n, dtype=np.int8)
        iq[0::2] = np.clip(np.rint(z.real + rng.normal(0, 1, n)), -128, 127).astype(np.int8)
        iq[1::2] = np.clip(np.rint(z.imag + rng.normal(0, 1, n)), -128, 127).astype(np.int8)
        pcm, mags, opens = ref.run_rx_squelch(mode, iq, -40, 16, block=blk)
        assert 0 < opens.sum() < 12
        out[f"squelch_{name}_iq"] = iq
        out[f"squelch_{name}_pcm"] = pcm
        out[f"squelch_{name}_mag"] = mags
        out[f"squelch_{name}_open"] = opens

    # ---- the stand-alone tools of signals/: heads piped into interpolateSignal, as generateBaseband.sh does ----
    import subprocess
    refdir = os.path.join(ROOT, "oracle", "_ref")

    def pipe(data, *programs):
        buf = data.tobytes()
        for prog in programs:
            buf = subprocess.run([os.path.join(refdir, prog)], input=buf, capture_output=True, check=True).stdout
        return np.frombuffer(buf, dtype=np.int8).copy()

    pcm = synth.tx_stream(96, stream=51, config=CONFIG, kind="sine")
    pcm2 = synth.tx_stream(96, stream=52, config=CONFIG, kind="noise")
    pairs = np.empty(192, dtype=np.int16)
    pairs[0::2], pairs[1::2] = pcm, pcm2
    out["sig_pcm"] = pcm2
    out["sig_pairs"] = pairs
    out["sig_iq8k_iq"] = pipe(pairs, "interpolateSignal")
    for head in ("dsb", "am", "pm", "fm"):
        out[f"sig_{head}_iq"] = pipe(pcm2, f"sig_{head}", "interpolateSignal")

    # ---- tables -------------------------------------------------------------------------
    for i, name in enumerate(TAPS):
        out[f"taps_{name}"] = ref.taps(i)
    s, c = ref.nco_tables()
    out["nco_sin_sha256"] = np.frombuffer(hashlib.sha256(s.tobytes()).digest(), dtype=np.uint8)
    out["nco_cos_sha256"] = np.frombuffer(hashlib.sha256(c.tobytes()).digest(), dtype=np.uint8)
    out["nco_sin_probe"] = s[::257].copy()
    out["nco_cos_probe"] = c[::257].copy()

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
