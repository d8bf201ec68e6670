"""GPU parity, small sizes: CUDA path (through the C ABI) vs the CPU oracle, bit for bit."""
import numpy as np
import pytest

from hackrfdiags_b200 import capi, synth

pytestmark = pytest.mark.gpu

RX_MODES = [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB]
NAMES = {1: "am", 2: "fm", 3: "wbfm", 4: "lsb", 5: "usb"}


def _diff(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()) if d.size else 0, int((d != 0).sum())


def test_tables_match_oracle(oracle):
    b = capi.Batch(1, capi.RX)
    assert np.array_equal(b.get_table(0).view(np.uint32), oracle.atan2_table().ravel().view(np.uint32))
    s, c = oracle.nco_tables()
    assert np.array_equal(b.get_table(1).view(np.uint32), s.view(np.uint32))
    assert np.array_equal(b.get_table(2).view(np.uint32), c.view(np.uint32))
    for i in range(13):
        assert np.array_equal(capi.get_taps(i), oracle.taps(i))


def test_front_end_bit_exact(oracle):
    n_streams, n = 12, 131072 + 4096
    iq = synth.rx_batch(capi.MODE_FM, n_streams, n)
    b = capi.Batch(n_streams, capi.RX)
    got = [b.rx_front_end(iq[:, :2 * 131072]), b.rx_front_end(iq[:, 2 * 131072:])]
    for s in range(n_streams):
        h = oracle.rx_new()
        w0 = oracle.rx_front_end(h, iq[s, :2 * 131072])
        w1 = oracle.rx_front_end(h, iq[s, 2 * 131072:])
        oracle.rx_free(h)
        assert np.array_equal(got[0][s], w0), f"stream {s} call 0"
        assert np.array_equal(got[1][s], w1), f"stream {s} call 1"


@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_2048k_bit_exact(oracle, mode):
    n_streams, n = 16, 2 * 131072 + 8192 + 256
    iq = synth.rx_batch(mode, n_streams, n)
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(mode)
    got = b.rx(iq)
    assert got.shape == (n_streams, n // 256)
    for s in range(n_streams):
        want = oracle.run_rx(mode, iq[s])
        mx, cnt = _diff(got[s], want)
        assert mx == 0, f"{NAMES[mode]} stream {s}: max abs err {mx}, {cnt} mismatches of {want.size}"


@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_256k_bit_exact(oracle, mode):
    n_streams, n = 8, 3 * 16384 + 32 * 5
    iq = synth.rx_batch(mode, n_streams, n, entry="256k")
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(mode)
    got = b.rx(iq, entry=capi.ENTRY_256K)
    for s in range(n_streams):
        want = oracle.run_rx(mode, iq[s], entry="256k")
        mx, cnt = _diff(got[s], want)
        assert mx == 0, f"{NAMES[mode]} stream {s}: max abs err {mx}, {cnt} mismatches"


@pytest.mark.parametrize("mode", RX_MODES)
def test_rx_streaming_state(oracle, mode):
    """Many calls of odd sizes == one long call == the oracle."""
    n_streams = 6
    sizes = [256, 512, 131072, 256 * 33, 8192, 256 * 7]
    iq = synth.rx_batch(mode, n_streams, sum(sizes), config=1)
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(mode)
    parts, off = [], 0
    for sz in sizes:
        parts.append(b.rx(np.ascontiguousarray(iq[:, 2 * off:2 * (off + sz)])))
        off += sz
    got = np.concatenate(parts, axis=1)
    for s in range(n_streams):
        want = oracle.run_rx(mode, iq[s])
        mx, cnt = _diff(got[s], want)
        assert mx == 0, f"{NAMES[mode]} stream {s}: max abs err {mx}, {cnt} mismatches"


@pytest.mark.parametrize("mode", RX_MODES)
def test_tx(oracle, mode):
    n_streams, n = 10, 32 * 5 + 17
    pcm = synth.tx_batch(n_streams, n)
    b = capi.Batch(n_streams, capi.TX)
    b.set_mode(mode)
    got = b.tx(pcm)
    tol = 0  # FM included: Nco::run's libm cosf / sinf are restated bit for bit (hrd_device.cuh glibc_sincosf)
    for s in range(n_streams):
        want = oracle.run_tx(mode, pcm[s])
        mx, cnt = _diff(got[s], want)
        assert mx <= tol, f"{NAMES[mode]} stream {s}: max abs err {mx}, {cnt} mismatches of {want.size}"


@pytest.mark.parametrize("mode", RX_MODES)
def test_tx_streaming_state(oracle, mode):
    n_streams = 5
    sizes = [1, 31, 32, 33, 512, 100]
    pcm = synth.tx_batch(n_streams, sum(sizes), config=2)
    b = capi.Batch(n_streams, capi.TX)
    b.set_mode(mode)
    parts, off = [], 0
    for sz in sizes:
        parts.append(b.tx(np.ascontiguousarray(pcm[:, off:off + sz])))
        off += sz
    got = np.concatenate(parts, axis=1)
    tol = 0
    for s in range(n_streams):
        want = oracle.run_tx(mode, pcm[s])
        mx, cnt = _diff(got[s], want)
        assert mx <= tol, f"{NAMES[mode]} stream {s}: max abs err {mx}, {cnt} mismatches"


@pytest.mark.parametrize("mode", [capi.MODE_AM, capi.MODE_FM, capi.MODE_LSB, capi.MODE_USB])
@pytest.mark.parametrize("tile_samples", [32, 64, 96, 160])
def test_tx_time_tiles(oracle, mode, tile_samples):
    """Tx tiles of 1..5 batches (halo 32 or 64 PCM samples), ragged last tile, three calls in a row; FM reads the
    NCO phases of the serial pre-pass.  Same bits as one tile."""
    n_streams = 7
    sizes = [32 * 11 + 5, 32 * 4, 77]
    pcm = synth.tx_batch(n_streams, sum(sizes), config=12)
    b = capi.Batch(n_streams, capi.TX)
    b.set_mode(mode)
    b.set_option(capi.OPT_TX_TILE_SAMPLES, tile_samples)
    parts, off = [], 0
    for sz in sizes:
        parts.append(b.tx(np.ascontiguousarray(pcm[:, off:off + sz])))
        off += sz
    got = np.concatenate(parts, axis=1)
    tol = 0
    for s in range(n_streams):
        want = oracle.run_tx(mode, pcm[s])
        mx, cnt = _diff(got[s], want)
        assert mx <= tol, f"{NAMES[mode]} tile={tile_samples} stream {s}: max abs err {mx}, {cnt} mismatches"


def test_tx_mixed_modes_one_batch(oracle):
    """All modulators plus the idle carrier in one batch, two calls; automatic tiling."""
    modes = [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB, capi.MODE_NONE] * 2
    n = 32 * 40 + 9
    pcm = synth.tx_batch(len(modes), n, config=13)
    b = capi.Batch(len(modes), capi.TX)
    for i, m in enumerate(modes):
        b.set_mode(m, i)
    got = np.concatenate([b.tx(np.ascontiguousarray(pcm[:, :700])), b.tx(np.ascontiguousarray(pcm[:, 700:]))], axis=1)
    for i, m in enumerate(modes):
        if m == capi.MODE_NONE:
            assert (got[i] == 64).all()  # BasebandDataProcessor.cc:689-694
            continue
        want = oracle.run_tx(m, pcm[i])
        err = np.abs(got[i].astype(np.int32) - want.astype(np.int32)).max()
        assert err == 0, f"stream {i} mode {m}: max abs err {err}"


# ---- time tiling: every tile size must give what one tile (= the serial order) gives -----
@pytest.mark.parametrize("mode", [capi.MODE_AM, capi.MODE_FM, capi.MODE_LSB, capi.MODE_USB])
@pytest.mark.parametrize("tile_batches", [1, 2, 3, 5])
def test_rx_time_tiles_bit_exact(oracle, mode, tile_batches):
    """Tiles of 1..5 batches (halo 1-2 batches each), ragged last tile, two calls in a row."""
    n_streams = 9
    sizes = [8192 * 7 + 256 * 3, 8192 * 4, 256 * 5]
    iq = synth.rx_batch(mode, n_streams, sum(sizes), config=3)
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(mode)
    b.set_option(capi.OPT_RX_TILE_BATCHES, tile_batches)
    parts, off = [], 0
    for sz in sizes:
        parts.append(b.rx(np.ascontiguousarray(iq[:, 2 * off:2 * (off + sz)])))
        off += sz
    got = np.concatenate(parts, axis=1)
    for s in range(n_streams):
        want = oracle.run_rx(mode, iq[s])
        mx, cnt = _diff(got[s], want)
        assert mx == 0, f"{NAMES[mode]} tile={tile_batches} stream {s}: max abs err {mx}, {cnt} mismatches"


@pytest.mark.parametrize("tile_batches", [2, 3])
def test_rx_front_end_time_tiles(oracle, tile_batches):
    n_streams, n = 6, 8192 * 9 + 512
    iq = synth.rx_batch(capi.MODE_FM, n_streams, n, config=4)
    b = capi.Batch(n_streams, capi.RX)
    b.set_option(capi.OPT_RX_TILE_BATCHES, tile_batches)
    got = b.rx_front_end(iq)
    for s in range(n_streams):
        h = oracle.rx_new()
        want = oracle.rx_front_end(h, iq[s])
        oracle.rx_free(h)
        assert np.array_equal(got[s], want), f"stream {s}"


@pytest.mark.parametrize("tile_batches", [2, 3, 5])
def test_rx_wbfm_time_tiles_bit_exact(oracle, tile_batches):
    """WBFM tiles by verified speculation: every tile's warmed-up recurrence value is checked against the
    true one on the device, so the output is bit-exact whatever the tile size (and no re-run was needed)."""
    n_streams = 8
    sizes = [8192 * 12, 8192 * 7 + 256 * 3, 8192 * 5]
    iq = synth.rx_batch(capi.MODE_WBFM, n_streams, sum(sizes), config=5)
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(capi.MODE_WBFM)
    assert b.get_option(capi.OPT_RX_WBFM_TILING) == 1  # on by default
    b.set_option(capi.OPT_RX_TILE_BATCHES, tile_batches)
    parts, off = [], 0
    for sz in sizes:
        parts.append(b.rx(np.ascontiguousarray(iq[:, 2 * off:2 * (off + sz)])))
        off += sz
    got = np.concatenate(parts, axis=1)
    for s in range(n_streams):
        want = oracle.run_rx(capi.MODE_WBFM, iq[s])
        mx, cnt = _diff(got[s], want)
        assert mx == 0, f"tile={tile_batches} stream {s}: max abs err {mx}, {cnt} mismatches"
    # the signal streams verify; only the constant-input edge streams (exactly-zero discriminator output, a
    # denormal tail no warm-up reproduces) may need the untiled re-run
    assert b.wbfm_fallback_count() <= 4 * len(sizes)


@pytest.mark.parametrize("force", [1, 2])
def test_rx_wbfm_failed_verification_reruns_exactly(oracle, force):
    """Force the verification to fail: the tiled retry (whose guess is wrong for a signal that is not constant, so
    its own verification sends the stream on; force = 2 fails that verification outright) and the serial re-run
    from the untouched state must give the same bits, call after call (state carry-over through the re-run path)."""
    n_streams, sizes = 5, [8192 * 9, 8192 * 6]
    iq = synth.rx_batch(capi.MODE_WBFM, n_streams, sum(sizes), config=9, with_edges=False)
    b = capi.Batch(n_streams, capi.RX)
    b.set_mode(capi.MODE_WBFM)
    b.set_option(capi.OPT_RX_TILE_BATCHES, 2)
    b.set_option(capi.OPT_DEBUG_WBFM_FORCE_RERUN, force)
    parts, off = [], 0
    for sz in sizes:
        parts.append(b.rx(np.ascontiguousarray(iq[:, 2 * off:2 * (off + sz)])))
        off += sz
    got = np.concatenate(parts, axis=1)
    for s in range(n_streams):
        assert np.array_equal(got[s], oracle.run_rx(capi.MODE_WBFM, iq[s])), f"stream {s}"
    assert b.wbfm_fallback_count() == 2 * n_streams
    assert b.wbfm_serial_count() == 2 * n_streams  # retried in tiles, failed again (a wrong guess, or force = 2), walked serially
    b.set_option(capi.OPT_RX_WBFM_TILING, 0)  # never tile: nothing to verify, nothing to re-run
    b.rx(np.ascontiguousarray(iq[:, :2 * 8192 * 4]))
    assert b.wbfm_fallback_count() == 2 * n_streams


def test_rx_wbfm_silent_and_constant_inputs_stay_tiled(oracle):
    """Inputs whose discriminator output is constant or exactly zero (rails, zeros, the Fs/2 pattern the front end
    removes) leave the de-emphasis recurrence on a limit cycle or on a vanishing value.  Tiles handle them: a
    warmed-up value that is nothing is replaced by the call's vanishing start value, vanishing values count as equal.
    A stream that falls silent BETWEEN calls has no such start value: its first verification fails, the tiled retry
    (every tile from the true value at the first check point) passes, nothing is walked serially.  All bit-exact."""
    sizes = [8192 * 10, 8192 * 14]
    n = sum(sizes)
    rows = [synth.rx_stream(capi.MODE_WBFM, n, stream=0, config=31),
            synth.rx_stream(capi.MODE_WBFM, n, stream=1, config=31, edge="min"),
            synth.rx_stream(capi.MODE_WBFM, n, stream=2, config=31, edge="max"),
            synth.rx_stream(capi.MODE_WBFM, n, stream=3, config=31, edge="zero"),
            synth.rx_stream(capi.MODE_WBFM, n, stream=4, config=31, edge="alt")]
    half = synth.rx_stream(capi.MODE_WBFM, n, stream=5, config=31, edge="max").copy()
    half[n:] = rows[0][n:]  # constant, then a signal (in the middle of the second call)
    silent = rows[0].copy()
    silent[2 * sizes[0]:] = rows[4][2 * sizes[0]:]  # a signal in the first call, the Fs/2 pattern in the second
    iq = np.stack(rows + [half, silent])
    b = capi.Batch(len(iq), capi.RX)
    b.set_mode(capi.MODE_WBFM)
    for s in range(len(iq)):
        b.set_param(capi.PARAM_WBFM_GAIN, 300.0 + 77.0 * s, s)
    b.set_option(capi.OPT_RX_TILE_BATCHES, 3)
    parts, off, counts = [], 0, []
    for sz in sizes:
        parts.append(b.rx(np.ascontiguousarray(iq[:, 2 * off:2 * (off + sz)])))
        counts.append((b.wbfm_fallback_count(), b.wbfm_serial_count()))
        off += sz
    got = np.concatenate(parts, axis=1)
    for s in range(len(iq)):
        assert np.array_equal(got[s], oracle.run_rx(capi.MODE_WBFM, iq[s], gain=300.0 + 77.0 * s)), f"stream {s}"
    assert counts[1][0] > counts[0][0], counts   # the stream that fell silent between the calls was retried ...
    assert counts[1][1] <= 1, counts             # ... and (at most the half-constant stream) nothing walked serially


def test_rx_mixed_modes_one_batch(oracle):
    """All five modes plus NONE in one batch and one call; auto tiling."""
    modes = [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB, capi.MODE_NONE] * 2
    n = 8192 * 6
    iq = np.stack([synth.rx_stream(m if m else capi.MODE_AM, n, stream=i, config=6) for i, m in enumerate(modes)])
    b = capi.Batch(len(modes), capi.RX)
    for i, m in enumerate(modes):
        b.set_mode(m, i)
    got = [b.rx(np.ascontiguousarray(iq[:, :2 * 8192 * 4])), b.rx(np.ascontiguousarray(iq[:, 2 * 8192 * 4:]))]
    got = np.concatenate(got, axis=1)
    for i, m in enumerate(modes):
        if m == capi.MODE_NONE:
            assert b.last_counts[i] == 0
            continue
        want = oracle.run_rx(m, iq[i])
        assert np.array_equal(got[i], want), f"stream {i} mode {m}"
