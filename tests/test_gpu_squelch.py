"""Squelch gate + signal magnitude on the GPU, through the C ABI, against the
oracle (itself pinned to the unmodified reference classes in tests/test_squelch_cpu.py): per-block
magnitudes and decisions, the PCM that comes out (bit-exact, shorter by the gated blocks), the tracker
state across calls, mixed modes and thresholds in one batch."""
import numpy as np
import pytest

from cpu_checkers import Oracle
from hackrfdiags_b200 import capi, synth

pytestmark = pytest.mark.gpu

MODES = [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB]
BLOCK = 262144


def _batch(modes, thresholds, gains_db):
    b = capi.Batch(len(modes), capi.RX, 0)
    for i, (m, t, g) in enumerate(zip(modes, thresholds, gains_db)):
        b.set_mode(m, i)
        b.set_param(capi.PARAM_SQUELCH_THRESHOLD, t, i)
        b.set_param(capi.PARAM_RX_GAIN_DB, g, i)
    return b


def test_squelch_matches_oracle_mixed_batch():
    oracle = Oracle()
    modes = MODES + [capi.MODE_NONE, capi.MODE_FM, capi.MODE_AM]
    thresholds = [-40, -30, -45, -25, -40, -40, -200, 0]
    gains = [16, 0, 16, 16, 40, 16, 16, 16]
    n_blocks = 12
    iq = np.stack([synth.rx_bursty_stream(m if m else 1, n_blocks, stream=i) for i, m in enumerate(modes)])
    b = _batch(modes, thresholds, gains)
    got = b.rx(iq)
    mags, allowed = b.squelch_report()
    assert mags.shape == (len(modes), n_blocks)
    for i, m in enumerate(modes):
        want_pcm, want_mag, want_open = oracle.run_rx_squelch(m, iq[i], thresholds[i], gains[i])
        assert np.array_equal(mags[i], want_mag), f"stream {i}: magnitudes"
        assert np.array_equal(allowed[i], want_open), f"stream {i}: decisions"
        assert b.last_counts[i] == want_pcm.size, f"stream {i}: {b.last_counts[i]} vs {want_pcm.size} PCM samples"
        assert np.array_equal(got[i, :want_pcm.size], want_pcm), f"stream {i} (mode {m}): PCM differs"
    assert 0 < allowed[0].sum() < n_blocks, "the test input never crossed the threshold"


def test_squelch_state_across_calls_and_short_last_block():
    """The tracker and the demodulator state persist: the same input in calls of 5, 1 and 6.x blocks."""
    oracle = Oracle()
    mode, thr = capi.MODE_FM, -38
    iq = synth.rx_bursty_stream(mode, 13, stream=2)[: 12 * BLOCK + 65536]
    want_pcm, want_mag, want_open = oracle.run_rx_squelch(mode, iq, thr)
    b = _batch([mode], [thr], [16])
    parts, opens = [], []
    for lo, hi in ((0, 5 * BLOCK), (5 * BLOCK, 6 * BLOCK), (6 * BLOCK, iq.size)):
        pcm = b.rx(iq[None, lo:hi].copy())
        parts.append(pcm[0, :b.last_counts[0]])
        opens.append(b.squelch_report()[1][0])
    assert np.array_equal(np.concatenate(opens), want_open)
    assert np.array_equal(np.concatenate(parts), want_pcm)


def test_default_threshold_takes_the_fused_path_and_option_forces_the_report():
    mode = capi.MODE_AM
    iq = np.stack([synth.rx_bursty_stream(mode, 4, stream=s) for s in range(3)])
    a = _batch([mode] * 3, [-200] * 3, [16] * 3)
    fused = a.rx(iq)
    assert a.squelch_report()[0].shape[1] == 0
    b = _batch([mode] * 3, [-200] * 3, [16] * 3)
    b.set_option(capi.OPT_RX_SQUELCH, 1)
    blockwise = b.rx(iq)
    mags, allowed = b.squelch_report()
    assert allowed.all() and mags.shape == (3, 4)
    assert np.array_equal(fused, blockwise), "block-by-block demodulation differs from the fused pass"
    want = Oracle().run_rx_squelch(mode, iq[1], -200)
    assert np.array_equal(mags[1], want[1]) and np.array_equal(blockwise[1], want[0])


def test_threshold_raised_after_fused_calls_keeps_the_tail():
    """Squelch::run runs on every block of the reference, also while no threshold can close the gate: the tracker
    sits in Tracking after loud blocks through the FUSED path, so when the threshold is raised the first quiet
    block is the tail (ENDOFSIGNAL) and still passes; the ones after it are closed."""
    oracle = Oracle()
    mode = capi.MODE_AM
    loud = synth.rx_stream(mode, 2 * 131072, stream=1)
    quiet = (synth.rx_stream(mode, 3 * 131072, stream=2).astype(np.int16) // 32).astype(np.int8)
    h = oracle.rx_new()
    try:
        oracle.rx_set_mode(h, mode)
        want1 = oracle.rx_accept_2048k(h, loud)           # default threshold: every block present
        oracle.lib.hro_rx_set_squelch_threshold(h, -40)
        want2 = oracle.rx_accept_2048k(h, quiet)          # tail block passes, two closed
    finally:
        oracle.rx_free(h)
    assert want1.size == 1024 and want2.size == 512
    b = capi.Batch(1, capi.RX, 0)
    b.set_mode(mode)
    got1 = b.rx(loud[None].copy())
    assert b.squelch_report()[0].shape[1] == 0            # the fused path ran
    b.set_param(capi.PARAM_SQUELCH_THRESHOLD, -40)
    got2 = b.rx(quiet[None].copy())
    assert np.array_equal(got1[0], want1)
    assert list(b.squelch_report()[1][0]) == [1, 0, 0]
    assert b.last_counts[0] == 512 and np.array_equal(got2[0, :512], want2)


def test_squelch_small_blocks_and_many_boundaries_per_iteration():
    """Blocks of 512 bytes (32 samples at 256 kS/s): four block ends inside every warp iteration of the gate kernel."""
    oracle = Oracle()
    mode, thr = capi.MODE_FM, -38
    iq = synth.rx_bursty_stream(mode, 2, stream=5)[: 40 * 512 * 16]
    # a level that flips every few hundred samples, so that neighbouring tiny blocks decide differently
    env = (np.arange(iq.size // 2) // 700) % 2
    iq = (iq.astype(np.int16) // (1 + 31 * np.repeat(env, 2))).astype(np.int8)
    want_pcm, want_mag, want_open = oracle.run_rx_squelch(mode, iq, thr, block=512)
    b = _batch([mode], [thr], [16])
    b.set_option(capi.OPT_RX_SQUELCH_BLOCK, 512)
    got = b.rx(iq[None].copy())
    mags, allowed = b.squelch_report()
    assert np.array_equal(mags[0], want_mag) and np.array_equal(allowed[0], want_open)
    assert 0 < want_open.sum() < want_open.size
    assert b.last_counts[0] == want_pcm.size and np.array_equal(got[0, :want_pcm.size], want_pcm)


def test_squelch_edge_inputs_device_memory():
    """Edge classes (full-range noise, constant -128 / +127, alternating, zero) with device pointers."""
    torch = pytest.importorskip("torch")
    oracle = Oracle()
    edges = ["noise", "min", "max", "alt", "zero"]
    n = 3 * 131072
    iq = np.stack([synth.rx_stream(capi.MODE_WBFM, n, stream=i, edge=e) for i, e in enumerate(edges)])
    thr = [-20, -5, -5, -10, -60]
    b = _batch([capi.MODE_WBFM] * 5, thr, [16] * 5)
    d_iq = torch.from_numpy(iq).cuda()
    d_pcm = torch.zeros((5, n // 256), dtype=torch.int16, device="cuda")
    b.rx_device(d_iq.data_ptr(), iq.shape[1], iq.shape[1], d_pcm.data_ptr(), d_pcm.shape[1])
    torch.cuda.synchronize()
    pcm = d_pcm.cpu().numpy()
    mags, allowed = b.squelch_report()
    for i in range(5):
        want_pcm, want_mag, want_open = oracle.run_rx_squelch(capi.MODE_WBFM, iq[i], thr[i])
        assert np.array_equal(mags[i], want_mag) and np.array_equal(allowed[i], want_open), edges[i]
        assert np.array_equal(pcm[i, :want_pcm.size], want_pcm), edges[i]


def test_fs4_rotation_on_its_own():
    """hrd_rx_fs4_rotate = IqDataProcessor::upconvertByFsOver4 / downconvertByFsOver4 (IqDataProcessor.cc:771-815,
    715-759), int8 negation wrapping (-(-128) = -128) included."""
    import ctypes as C
    rng = np.random.default_rng(8)
    iq = rng.integers(-128, 128, size=8 * 1000, dtype=np.int64).astype(np.int8)
    iq[:16] = -128
    z = iq.reshape(-1, 4, 2).astype(np.int16)  # groups of four (I, Q) samples

    def neg(v):
        return (-v).astype(np.int8)  # wraps -(-128) to -128 like the int8_t assignment in the reference

    up = z.copy().astype(np.int8)
    up[:, 1, 0], up[:, 1, 1] = neg(z[:, 1, 1]), z[:, 1, 0].astype(np.int8)
    up[:, 2, 0], up[:, 2, 1] = neg(z[:, 2, 0]), neg(z[:, 2, 1])
    up[:, 3, 0], up[:, 3, 1] = z[:, 3, 1].astype(np.int8), neg(z[:, 3, 0])
    b = capi.Batch(1, capi.RX, 0)
    buf = iq.copy()
    assert b.lib.hrd_rx_fs4_rotate(b.h, buf.ctypes.data, buf.size, 1, capi.MEM_HOST, None) == 0
    assert np.array_equal(buf, up.reshape(-1))
    # down undoes up wherever no -128 is involved; and the front end is reduce + up
    ok = (np.abs(z) < 128).all(axis=(1, 2))
    assert b.lib.hrd_rx_fs4_rotate(b.h, buf.ctypes.data, buf.size, 0, capi.MEM_HOST, None) == 0
    assert np.array_equal(buf.reshape(-1, 8)[ok], iq.reshape(-1, 8)[ok])
