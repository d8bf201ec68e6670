"""Pin the C restatement (oracle/hrd_oracle.c) against the compiled, unmodified
reference (oracle/_ref/libhrd_ref.so): bit-exact on every mode, both entries,
tables, streaming in odd block sizes, resets and parameter setters."""
import numpy as np
import pytest

from cpu_checkers import AM, FM, WBFM, LSB, USB, DEMOD_OF_MODE, MOD_OF_MODE, TAPS
from hackrfdiags_b200 import synth

RX_MODES = [AM, FM, WBFM, LSB, USB]


def test_quantised_taps_match_reference(oracle, ref):
    for i, name in enumerate(TAPS):
        a, b = oracle.taps(i), ref.taps(i)
        assert a.shape == b.shape and np.array_equal(a, b), name
    # the quirks SURVEY.md section 7.1 calls out
    assert oracle.taps(TAPS.index("ssb_delay"))[-1] == -32768
    assert list(oracle.taps(TAPS.index("fe1"))) == [8206, 16384, 8206]
    assert list(oracle.taps(TAPS.index("fe2"))) == [8249, 16384, 8249]
    assert list(oracle.taps(TAPS.index("fe3"))) == [8424, 16384, 8424]


def test_nco_tables_match_reference(oracle, ref):
    so, co = oracle.nco_tables()
    sr, cr = ref.nco_tables()
    assert np.array_equal(so.view(np.uint32), sr.view(np.uint32))
    assert np.array_equal(co.view(np.uint32), cr.view(np.uint32))


@pytest.mark.parametrize("edge", [None, "noise", "min", "max", "alt", "zero"])
def test_front_end(oracle, ref, edge):
    iq = synth.rx_stream(FM, 131072, stream=3, edge=edge)
    ho, hr = oracle.rx_new(), ref.rx_new()
    for blk in range(2):  # two calls: state carries over
        a = oracle.rx_front_end(ho, iq[blk * 131072:(blk + 1) * 131072])
        b = ref.rx_front_end(hr, iq[blk * 131072:(blk + 1) * 131072])
        assert a.size == 16384 and np.array_equal(a, b)
    oracle.rx_free(ho)
    ref.rx_free(hr)


@pytest.mark.parametrize("mode", RX_MODES)
@pytest.mark.parametrize("edge", [None, "noise", "min", "max", "alt", "zero"])
def test_rx_2048k(oracle, ref, mode, edge):
    n = 3 * 131072 + 4096  # three blocks and a short one
    iq = synth.rx_stream(mode, n, stream=5, edge=edge)
    a = oracle.run_rx(mode, iq)
    b = ref.run_rx(mode, iq)
    assert a.size == n // 256
    assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", RX_MODES)
@pytest.mark.parametrize("edge", [None, "noise", "min"])
def test_rx_256k(oracle, ref, mode, edge):
    n = 5 * 16384
    iq = synth.rx_stream(mode, n, stream=7, entry="256k", edge=edge)
    a = oracle.run_rx(mode, iq, entry="256k")
    b = ref.run_rx(mode, iq, entry="256k")
    assert a.size == n // 32
    assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_block_size_invariance_and_gain(oracle, ref, mode):
    n = 2 * 131072
    iq = synth.rx_stream(mode, n, stream=11)
    for gain in (None, 30.0, 30000.0):
        a = oracle.run_rx(mode, iq, gain=gain, block=262144)
        b = ref.run_rx(mode, iq, gain=gain, block=262144)
        c = oracle.run_rx(mode, iq, gain=gain, block=2048)
        assert np.array_equal(a, b)
        assert np.array_equal(a, c)


@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_reset_and_mode_switch(oracle, ref, mode):
    iq = synth.rx_stream(mode, 3 * 131072, stream=13)
    outs = []
    for lib in (oracle, ref):
        h = lib.rx_new()
        lib.rx_set_mode(h, mode)
        o = [lib.rx_accept_2048k(h, iq[:262144])]
        lib.rx_reset_demod(h, DEMOD_OF_MODE[mode])
        o.append(lib.rx_accept_2048k(h, iq[262144:524288]))
        other = FM if mode != FM else AM
        lib.rx_set_mode(h, other)
        o.append(lib.rx_accept_2048k(h, iq[524288:]))
        lib.rx_set_mode(h, mode)
        o.append(lib.rx_accept_2048k(h, iq[:262144]))
        lib.rx_free(h)
        outs.append(np.concatenate(o))
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("mode", RX_MODES)
@pytest.mark.parametrize("kind", synth.TX_CLASSES)
def test_tx(oracle, ref, mode, kind):
    pcm = synth.tx_stream(1024 + 37, stream=2, kind=kind)
    a = oracle.run_tx(mode, pcm)
    b = ref.run_tx(mode, pcm)
    if mode == FM:
        # Nco::run calls libm sinf/cosf: same libm here, so still bit-exact
        assert np.array_equal(a, b)
    else:
        assert np.array_equal(a, b)


def test_tx_parameters_and_reset(oracle, ref):
    pcm = synth.tx_stream(700, stream=1, kind="speechlike")
    for lib_pair in [(oracle, ref)]:
        res = []
        for lib in lib_pair:
            h = lib.tx_new()
            lib.tx_set_am_index(h, 0.3)
            lib.tx_set_am_index(h, 1.5)       # rejected
            lib.tx_set_fm_deviation(h, 1000)
            lib.tx_set_fm_deviation(h, 9000)  # accepted: the guard looks at the old value
            lib.tx_set_fm_deviation(h, 100)   # rejected: old value 9000 is out of range
            lib.tx_set_wbfm_deviation(h, 112000)
            o = []
            for mode in (AM, FM, WBFM, LSB, USB):
                o.append(lib.tx_accept(h, mode, pcm[:300]))
                lib.tx_reset_mod(h, MOD_OF_MODE[mode])
                o.append(lib.tx_accept(h, mode, pcm[300:]))
            lib.tx_free(h)
            res.append(np.concatenate(o))
        assert np.array_equal(res[0], res[1])


def test_tx_count_raw_like_full_scale(oracle, ref):
    # full-range ramp including -32768 and 32767, as signals/count.raw spans
    pcm = np.concatenate([np.arange(-32768, 32768, 257), np.arange(32767, -32769, -513)]).astype(np.int16)
    for mode in (AM, FM, WBFM, LSB, USB):
        assert np.array_equal(oracle.run_tx(mode, pcm), ref.run_tx(mode, pcm))
