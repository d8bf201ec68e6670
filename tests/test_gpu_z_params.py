"""GPU parity with NON-DEFAULT parameters, a different setting on every stream of one batch, through the C ABI
(gains x0.1 / x10, modulation index 0.3 / 1.0, deviation 1000 / 112000; plus the setters' guards).
The same sweeps run oracle-vs-reference on the CPU in tests/test_params_cpu.py.
(File name: collected last, after the parity tests it extends.)"""
import numpy as np
import pytest

from hackrfdiags_b200 import capi, synth
from param_sweeps import GAIN_PARAM, NAMES, RX_MODES, rx_gain_sweep, tx_setter_sweeps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("entry", ["2048k", "256k"])
@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_gain_per_stream(oracle, mode, entry):
    b = capi.Batch(1, capi.RX)
    default = b.get_param(GAIN_PARAM[mode], 0)
    b.close()
    gains = rx_gain_sweep(default)
    n_streams = len(gains)
    n = 131072 + 8192 if entry == "2048k" else 16384 + 1024
    iq = synth.rx_batch(mode, n_streams, n, config=3, entry=entry)
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(mode)
    for s, g in enumerate(gains):
        if g is not None:
            b.set_param(GAIN_PARAM[mode], g, s)
    got = b.rx(iq, entry=capi.ENTRY_2048K if entry == "2048k" else capi.ENTRY_256K)
    for s, g in enumerate(gains):
        want = oracle.run_rx(mode, iq[s], entry=entry, gain=g)
        bad = int((got[s] != want).sum())
        assert bad == 0, f"{NAMES[mode]} {entry} stream {s} gain {g}: {bad} of {want.size} PCM samples differ"


@pytest.mark.parametrize("mode", [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM])
def test_tx_setters_per_stream(oracle, mode):
    """Every stream gets its own sequence of setter calls (accepted and rejected ones: the guards of
    AmModulator.cc:329-339, FmModulator.cc:336-346, WbFmModulator.cc:310-328 are part of the interface)."""
    param = {capi.MODE_AM: capi.PARAM_AM_INDEX, capi.MODE_FM: capi.PARAM_FM_DEV, capi.MODE_WBFM: capi.PARAM_WBFM_DEV}[mode]
    sweeps = tx_setter_sweeps(mode)
    n_streams, n = len(sweeps), 32 * 6 + 5
    pcm = synth.tx_batch(n_streams, n, config=3)
    b = capi.Batch(n_streams, capi.TX)
    b.set_mode(mode)
    for s, calls in enumerate(sweeps):
        for v in calls:
            b.set_param(param, v, s)
    got = b.tx(pcm)
    tol = 0  # FM included (hrd_device.cuh glibc_sincosf)
    for s, calls in enumerate(sweeps):
        h = oracle.tx_new()
        setter = {capi.MODE_AM: oracle.tx_set_am_index, capi.MODE_FM: oracle.tx_set_fm_deviation,
                  capi.MODE_WBFM: oracle.tx_set_wbfm_deviation}[mode]
        for v in calls:
            setter(h, v)
        want = oracle.tx_accept(h, mode, pcm[s])
        oracle.tx_free(h)
        err = int(np.abs(got[s].astype(np.int32) - want.astype(np.int32)).max())
        assert err <= tol, f"{NAMES[mode]} stream {s} setters {calls}: max abs err {err}"
