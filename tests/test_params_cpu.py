"""The parameter sweeps of tests/test_gpu_z_params.py, oracle against the compiled unmodified reference (CPU only): the
oracle the GPU test trusts is pinned on exactly these settings."""
import numpy as np
import pytest

from hackrfdiags_b200 import capi, synth
from param_sweeps import DEFAULT_GAIN, NAMES, RX_MODES, rx_gain_sweep, tx_setter_sweeps


@pytest.mark.parametrize("entry", ["2048k", "256k"])
@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_gain_sweep_oracle_vs_reference(oracle, ref, mode, entry):
    gains = rx_gain_sweep(DEFAULT_GAIN[mode])
    n = 131072 + 8192 if entry == "2048k" else 16384 + 1024
    iq = synth.rx_batch(mode, len(gains), n, config=3, entry=entry)
    for s, g in enumerate(gains):
        a = oracle.run_rx(mode, iq[s], entry=entry, gain=g)
        b = ref.run_rx(mode, iq[s], entry=entry, gain=g)
        assert np.array_equal(a, b), f"{NAMES[mode]} {entry} stream {s} gain {g}"


@pytest.mark.parametrize("mode", [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM])
def test_tx_setter_sweeps_oracle_vs_reference(oracle, ref, mode):
    sweeps = tx_setter_sweeps(mode)
    pcm = synth.tx_batch(len(sweeps), 32 * 6 + 5, config=3)
    outs = []
    for lib in (oracle, ref):
        rows = []
        for s, calls in enumerate(sweeps):
            h = lib.tx_new()
            setter = {capi.MODE_AM: lib.tx_set_am_index, capi.MODE_FM: lib.tx_set_fm_deviation,
                      capi.MODE_WBFM: lib.tx_set_wbfm_deviation}[mode]
            for v in calls:
                setter(h, v)
            rows.append(lib.tx_accept(h, mode, pcm[s]))
            lib.tx_free(h)
        outs.append(rows)
    for s, calls in enumerate(sweeps):
        assert np.array_equal(outs[0][s], outs[1][s]), f"{NAMES[mode]} stream {s} setters {calls}"
    # the sweeps are not vacuous: different settings give different outputs
    assert any(not np.array_equal(outs[0][0], outs[0][s]) for s in range(1, len(sweeps)))


def test_sincosf_restatement_is_libm(tmp_path):
    """The FM / PM transmit heads evaluate libm's cosf / sinf on the GPU as a restatement of glibc's algorithm
    (hrd_device.cuh glibc_sincosf).  tools/verify_sincosf.c holds the same restatement in C and compares it with the
    host's libm -- the one the reference and the oracle call -- for every float below 8 in magnitude when run without
    an argument (22 s: 2.18e9 floats, 0 mismatches); here every 61st float."""
    import os
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    exe = tmp_path / "verify_sincosf"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tools", "verify_sincosf.c"), "-lm"], check=True)
    out = subprocess.run([str(exe), "61"], check=True, capture_output=True, text=True).stdout
    assert "sinf 0 mismatches, cosf 0" in out.splitlines()[1] and "sinf 0 mismatches, cosf 0" in out.splitlines()[2], out


def test_packed_half_tx_tail_is_the_integer_tail(tmp_path):
    """Stages 6-8 of the transmit interpolator run on both rails at once as fp16 pairs (hrd_tx.cu tail3_h2) wherever
    the stage-6 inputs stay within +-995 (checked per warp iteration, integer form otherwise).
    tools/verify_tx_tail_h2.c emulates the fp16 arithmetic exactly and compares all 3 964 081 input pairs of that
    range with the reference's integer arithmetic."""
    import os
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    exe = tmp_path / "verify_tx_tail_h2"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tools", "verify_tx_tail_h2.c"), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "3964081 (x, xm) pairs, 0 mismatches" in out, out
    assert subprocess.run([str(exe), "1000"], capture_output=True, text=True).returncode != 0  # the range is tight
