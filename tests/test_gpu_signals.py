"""The signals/ tool chain on the GPU: interpolateSignal's eight-stage
interpolator with its own stage-1 taps behind raw I,Q pairs or the dsb / am / pm prototype heads, through the
C ABI, against the oracle (pinned to the reference's own programs in tests/test_signals_cpu.py)."""
import numpy as np
import pytest

from cpu_checkers import Oracle
from hackrfdiags_b200 import capi, synth

pytestmark = pytest.mark.gpu
HEAD_OF_MODE = {capi.MODE_IQ8K: 0, capi.MODE_DSB: 1, capi.MODE_AM_PROTO: 2, capi.MODE_PM: 3, capi.MODE_FM_PROTO: 4}


def test_mixed_batch_of_heads_and_modulators_streaming():
    """Eight streams: the four signals/ heads twice, next to two reference modulators, three calls of ragged
    lengths (tiles, halos and carried state), rows of 2n int16 because the batch holds I,Q-pair streams."""
    oracle = Oracle()
    modes = [capi.MODE_IQ8K, capi.MODE_DSB, capi.MODE_AM_PROTO, capi.MODE_PM, capi.MODE_PM, capi.MODE_IQ8K,
             capi.MODE_DSB, capi.MODE_AM_PROTO, capi.MODE_LSB, capi.MODE_AM, capi.MODE_FM_PROTO, capi.MODE_FM,
             capi.MODE_FM_PROTO]
    n_total, cuts = 800, [0, 320, 352, 800]
    kinds = ["sine", "noise", "square", None]
    pcm = [synth.tx_stream(n_total, stream=i, config=9, kind=kinds[i % 4]) for i in range(len(modes))]
    other = [synth.tx_stream(n_total, stream=100 + i, config=9, kind="noise") for i in range(len(modes))]
    b = capi.Batch(len(modes), capi.TX, 0)
    for i, m in enumerate(modes):
        b.set_mode(m, i)
    got = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        n = hi - lo
        rows = np.zeros((len(modes), 2 * n), dtype=np.int16)
        for i, m in enumerate(modes):
            if m == capi.MODE_IQ8K:
                rows[i, 0::2], rows[i, 1::2] = pcm[i][lo:hi], other[i][lo:hi]
            else:
                rows[i, :n] = pcm[i][lo:hi]
        got.append(b.tx(rows, n=n))
    got = np.concatenate(got, axis=1)
    for i, m in enumerate(modes):
        if m in HEAD_OF_MODE:
            data = pcm[i]
            if m == capi.MODE_IQ8K:
                data = np.empty(2 * n_total, dtype=np.int16)
                data[0::2], data[1::2] = pcm[i], other[i]
            want = oracle.run_tx_signals(HEAD_OF_MODE[m], data)
        else:
            want = oracle.run_tx(m, pcm[i])
        err = np.abs(got[i].astype(np.int32) - want.astype(np.int32)).max()
        assert err == 0, f"stream {i} mode {m}: max abs err {err}"  # PM / FM heads too: libm's cosf / sinf bit for bit


def test_many_streams_tiled():
    """4096 PM/DSB streams x 0.25 s: the tile chooser cuts the call; identical streams give identical bytes and a
    sample of them matches the oracle."""
    oracle = Oracle()
    n_streams, n = 512, 2000
    distinct = [synth.tx_stream(n, stream=i, config=10) for i in range(4)]
    rows = np.stack([distinct[i % 4] for i in range(n_streams)])
    b = capi.Batch(n_streams, capi.TX, 0)
    for i in range(n_streams):
        b.set_mode(capi.MODE_DSB if (i // 4) % 2 else capi.MODE_AM_PROTO, i)
    got = b.tx(rows)
    for i in (0, 1, 6, 7, 509, 511):
        head = 1 if (i // 4) % 2 else 2
        assert np.array_equal(got[i], oracle.run_tx_signals(head, rows[i])), i
    assert np.array_equal(got[0], got[8]) and np.array_equal(got[5], got[13])


def test_reset_and_mode_validation():
    b = capi.Batch(1, capi.TX, 0)
    b.set_mode(capi.MODE_DSB)
    pcm = synth.tx_stream(64, stream=1, config=11, kind="noise")[None, :].copy()
    first = b.tx(pcm)
    b.tx(pcm)
    b.reset(capi.UNIT_SIGNALS)
    assert np.array_equal(b.tx(pcm), first)
    rx = capi.Batch(1, capi.RX, 0)
    with pytest.raises(capi.HrdError):
        rx.set_mode(capi.MODE_PM)
    b.set_mode(capi.MODE_IQ8K)
    with pytest.raises(capi.HrdError):
        b.tx(np.zeros((1, 65), dtype=np.int16), n=64)  # 2n int16 do not fit the row


def test_packed_half_tail_across_its_range_limit():
    """Stages 6-8 run as fp16 pairs while the stage-6 inputs stay within +-995 and as integers beyond (a vote per
    warp iteration, hrd_tx.cu).  Raw I,Q tones whose amplitude ramps through the level where that happens -- on one
    rail only, on both, in phase and in quadrature, with a DC offset -- so that single iterations, single lanes and
    whole streams sit on either side of the limit and switch back and forth: bit for bit against the oracle."""
    oracle = Oracle()
    n = 1536
    t = np.arange(n)
    ramp = np.linspace(27000.0, 32767.0, n)
    rows, wants = [], []
    cases = [(ramp, 0.0, 0.031, 0.0), (32767.0 - (ramp - 27000.0), 0.25, 0.013, 0.0), (ramp, 0.5, 0.047, 1500.0),
             (np.full(n, 31200.0), 0.0, 0.002, 0.0), (np.full(n, 31900.0), 0.125, 0.11, -700.0), (ramp * 0.5, 0.0, 0.2, 0.0)]
    for amp, quad, f, dc in cases:
        i = np.clip(np.round(amp * np.cos(2 * np.pi * f * t) + dc), -32768, 32767).astype(np.int16)
        q = np.clip(np.round(0.3 * amp * np.sin(2 * np.pi * (f * t + quad))), -32768, 32767).astype(np.int16)
        data = np.empty(2 * n, dtype=np.int16)
        data[0::2], data[1::2] = i, q
        rows.append(data)
        wants.append(oracle.run_tx_signals(HEAD_OF_MODE[capi.MODE_IQ8K], data))
    b = capi.Batch(len(rows), capi.TX, 0)
    b.set_mode(capi.MODE_IQ8K)
    got = np.concatenate([b.tx(np.stack([r[:2 * 700] for r in rows]), n=700), b.tx(np.stack([r[2 * 700:] for r in rows]), n=n - 700)], axis=1)
    for k, want in enumerate(wants):
        d = np.abs(got[k].astype(np.int32) - want.astype(np.int32))
        assert d.max() == 0, f"case {k}: max abs err {d.max()}, {(d != 0).sum()} of {d.size} differ"
