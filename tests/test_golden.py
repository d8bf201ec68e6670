"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py ran it in the
dev container through oracle/_ref/libhrd_ref.so).  They travel with the repo, so they pin

  * the CPU oracle  (not gpu): every vector bit for bit;
  * the CUDA path   (gpu)    : every vector bit for bit (FM Tx included),
                               called through the C ABI exactly as the reference calls were made.
"""
import hashlib
import os

import numpy as np
import pytest

from cpu_checkers import DEMOD_OF_MODE, TAPS

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {"am": 1, "fm": 2, "wbfm": 3, "lsb": 4, "usb": 5}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))


# ------------------------------------------------------------------------------------------
# oracle vs golden (CPU)
# ------------------------------------------------------------------------------------------
SIG_HEADS = {"iq8k": 0, "dsb": 1, "am": 2, "pm": 3, "fm": 4}


@pytest.mark.parametrize("head", list(SIG_HEADS))
def test_oracle_signals_chain(oracle, golden, head):
    """signals/ tool chain: <head> | interpolateSignal, two calls (state carries)."""
    data = golden["sig_pairs"] if head == "iq8k" else golden["sig_pcm"]
    got = oracle.run_tx_signals(SIG_HEADS[head], data, chunks=[64, 32])
    assert np.array_equal(got, golden[f"sig_{head}_iq"])


@pytest.mark.parametrize("name", ["am", "fm", "usb"])
def test_oracle_squelch(oracle, golden, name):
    """Squelch gate: 12 reference calls of 8192 bytes, threshold -40 dBFS."""
    pcm, mags, opens = oracle.run_rx_squelch(MODES[name], golden[f"squelch_{name}_iq"], -40, 16, block=8192)
    assert np.array_equal(mags, golden[f"squelch_{name}_mag"])
    assert np.array_equal(opens, golden[f"squelch_{name}_open"])
    assert np.array_equal(pcm, golden[f"squelch_{name}_pcm"])


@pytest.mark.parametrize("name", list(MODES))
@pytest.mark.parametrize("tag", ["sig", "noise"])
def test_oracle_rx_2048k(oracle, golden, name, tag):
    iq, cut = golden[f"rx2048k_{name}_{tag}_iq"], int(golden[f"rx2048k_{name}_{tag}_cut"])
    h = oracle.rx_new()
    oracle.rx_set_mode(h, MODES[name])
    got = np.concatenate([oracle.rx_accept_2048k(h, iq[:cut]), oracle.rx_accept_2048k(h, iq[cut:])])
    oracle.rx_free(h)
    assert np.array_equal(got, golden[f"rx2048k_{name}_{tag}_pcm"])


@pytest.mark.parametrize("name", list(MODES))
@pytest.mark.parametrize("tag,gain", [("g0", None), ("g1", 1234.5)])
def test_oracle_rx_256k(oracle, golden, name, tag, gain):
    iq = golden[f"rx256k_{name}_iq"]
    h = oracle.rx_new()
    oracle.rx_set_mode(h, MODES[name])
    if gain is not None:
        oracle.rx_set_gain(h, DEMOD_OF_MODE[MODES[name]], gain)
    got = np.concatenate([oracle.rx_accept_256k(h, iq[:2048]), oracle.rx_accept_256k(h, iq[2048:])])
    oracle.rx_free(h)
    assert np.array_equal(got, golden[f"rx256k_{name}_{tag}_pcm"])


@pytest.mark.parametrize("tag", ["noise", "max", "min"])
def test_oracle_front_end(oracle, golden, tag):
    h = oracle.rx_new()
    got = oracle.rx_front_end(h, golden[f"fe_{tag}_iq"])
    oracle.rx_free(h)
    assert np.array_equal(got, golden[f"fe_{tag}_out"])


@pytest.mark.parametrize("name", list(MODES))
@pytest.mark.parametrize("kind", ["sine", "noise"])
def test_oracle_tx(oracle, golden, name, kind):
    pcm = golden[f"tx_{name}_{kind}_pcm"]
    h = oracle.tx_new()
    got = np.concatenate([oracle.tx_accept(h, MODES[name], pcm[:64]), oracle.tx_accept(h, MODES[name], pcm[64:])])
    oracle.tx_free(h)
    assert np.array_equal(got, golden[f"tx_{name}_{kind}_iq"])


def test_oracle_tables(oracle, golden):
    for i, name in enumerate(TAPS):
        assert np.array_equal(oracle.taps(i), golden[f"taps_{name}"]), name
    s, c = oracle.nco_tables()
    assert hashlib.sha256(s.tobytes()).digest() == golden["nco_sin_sha256"].tobytes()
    assert hashlib.sha256(c.tobytes()).digest() == golden["nco_cos_sha256"].tobytes()


# ------------------------------------------------------------------------------------------
# CUDA path vs golden (B200), through the C ABI
# ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MODES))
@pytest.mark.parametrize("tag", ["sig", "noise"])
def test_cuda_rx_2048k(golden, name, tag):
    from hackrfdiags_b200 import capi
    iq, cut = golden[f"rx2048k_{name}_{tag}_iq"], int(golden[f"rx2048k_{name}_{tag}_cut"])
    b = capi.Batch(1, capi.RX)
    b.set_mode(MODES[name])
    got = np.concatenate([b.rx(iq[None, :cut].copy()), b.rx(iq[None, cut:].copy())], axis=1)[0]
    assert np.array_equal(got, golden[f"rx2048k_{name}_{tag}_pcm"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MODES))
@pytest.mark.parametrize("tag,gain", [("g0", None), ("g1", 1234.5)])
def test_cuda_rx_256k(golden, name, tag, gain):
    from hackrfdiags_b200 import capi
    iq = golden[f"rx256k_{name}_iq"]
    b = capi.Batch(1, capi.RX)
    b.set_mode(MODES[name])
    if gain is not None:
        b.set_param([capi.PARAM_AM_GAIN, capi.PARAM_FM_GAIN, capi.PARAM_WBFM_GAIN, capi.PARAM_SSB_GAIN]
                    [DEMOD_OF_MODE[MODES[name]]], gain)
    got = np.concatenate([b.rx(iq[None, :2048].copy(), entry=capi.ENTRY_256K),
                          b.rx(iq[None, 2048:].copy(), entry=capi.ENTRY_256K)], axis=1)[0]
    assert np.array_equal(got, golden[f"rx256k_{name}_{tag}_pcm"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["noise", "max", "min"])
def test_cuda_front_end(golden, tag):
    from hackrfdiags_b200 import capi
    b = capi.Batch(1, capi.RX)
    got = b.rx_front_end(golden[f"fe_{tag}_iq"][None, :].copy())[0]
    assert np.array_equal(got, golden[f"fe_{tag}_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MODES))
@pytest.mark.parametrize("kind", ["sine", "noise"])
def test_cuda_tx(golden, name, kind):
    from hackrfdiags_b200 import capi
    pcm = golden[f"tx_{name}_{kind}_pcm"]
    b = capi.Batch(1, capi.TX)
    b.set_mode(MODES[name])
    got = np.concatenate([b.tx(pcm[None, :64].copy()), b.tx(pcm[None, 64:].copy())], axis=1)[0]
    want = golden[f"tx_{name}_{kind}_iq"]
    err = np.abs(got.astype(np.int32) - want.astype(np.int32)).max()
    assert err <= (1 if name == "fm" else 0), f"max abs err {err}"


@pytest.mark.gpu
def test_cuda_tables(golden):
    from hackrfdiags_b200 import capi
    for i, name in enumerate(TAPS):
        assert np.array_equal(capi.get_taps(i), golden[f"taps_{name}"]), name
    b = capi.Batch(1, capi.TX)
    assert hashlib.sha256(b.get_table(1).tobytes()).digest() == golden["nco_sin_sha256"].tobytes()
    assert hashlib.sha256(b.get_table(2).tobytes()).digest() == golden["nco_cos_sha256"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["am", "fm", "usb"])
def test_cuda_squelch(golden, name):
    from hackrfdiags_b200 import capi
    b = capi.Batch(1, capi.RX, 0)
    b.set_mode(MODES[name])
    b.set_param(capi.PARAM_SQUELCH_THRESHOLD, -40)
    b.set_option(capi.OPT_RX_SQUELCH_BLOCK, 8192)
    got = b.rx(golden[f"squelch_{name}_iq"][None, :].copy())
    mags, opens = b.squelch_report()
    want = golden[f"squelch_{name}_pcm"]
    assert np.array_equal(mags[0], golden[f"squelch_{name}_mag"])
    assert np.array_equal(opens[0], golden[f"squelch_{name}_open"])
    assert b.last_counts[0] == want.size and np.array_equal(got[0, :want.size], want)


@pytest.mark.gpu
@pytest.mark.parametrize("head", list(SIG_HEADS))
def test_cuda_signals_chain(golden, head):
    from hackrfdiags_b200 import capi
    mode = {"iq8k": capi.MODE_IQ8K, "dsb": capi.MODE_DSB, "am": capi.MODE_AM_PROTO, "pm": capi.MODE_PM,
            "fm": capi.MODE_FM_PROTO}[head]
    b = capi.Batch(1, capi.TX, 0)
    b.set_mode(mode)
    if head == "iq8k":
        pairs = golden["sig_pairs"]
        got = np.concatenate([b.tx(pairs[None, :128].copy(), n=64)[0], b.tx(pairs[None, 128:].copy(), n=32)[0]])
    else:
        pcm = golden["sig_pcm"]
        got = np.concatenate([b.tx(pcm[None, :64].copy())[0], b.tx(pcm[None, 64:].copy())[0]])
    err = np.abs(got.astype(np.int32) - golden[f"sig_{head}_iq"].astype(np.int32)).max()
    assert err <= (1 if head in ("pm", "fm") else 0), f"{head}: max abs err {err}"
