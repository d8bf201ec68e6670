"""Parity at BASELINE.json's full sizes (configs 2, 3, 4), through the C ABI on device buffers.

The oracle cannot chew through 4-17 GB in a test, so full size is covered by size-independent properties:
  * EVERY distinct stream of the batch (32 per mode, the last five of them the edge classes: full-range noise,
    constant -128, constant +127, alternating +-127, zero), whole length, bit for bit against the CPU oracle
    (FM Tx included);
  * streams fed identical inputs give identical outputs wherever they sit in the batch (independence): every
    repeat of every distinct row, compared on the device;
  * automatic time tiling == one tile per stream (the serial order);
  * one long call == the same signal in several calls (state carry-over);
Inputs are generated on the device (bench.py's generators: 32 distinct rows per mode, tiled)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS = 2_048_000


@pytest.fixture(scope="module")
def env():
    import torch
    import bench
    from hackrfdiags_b200 import capi
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    return torch, bench, capi, dev


def _rx_call(capi, b, iq, pcm, lo=0, hi=None, stream=0):
    hi = iq.shape[1] if hi is None else hi
    b.rx_device(iq.data_ptr() + lo, hi - lo, iq.stride(0), pcm.data_ptr() + (lo // 512) * 2, pcm.stride(0),
                capi.ENTRY_2048K, stream)


def _distinct_rows(layout):
    """the first occurrence of every distinct row of a batch built by bench.make_rx_batch"""
    return [at + k for at, nd, _ in layout.values() for k in range(nd)]


def _check_sample_vs_oracle(torch, oracle, iq, pcm, modes, picks):
    worst = 0
    for s in picks:
        want = oracle.run_rx(modes[s], iq[s].cpu().numpy())
        got = pcm[s].cpu().numpy()
        assert got.shape == want.shape
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        worst = max(worst, int(d.max()))
        assert worst == 0, f"stream {s} (mode {modes[s]}): max abs err {d.max()}, {(d != 0).sum()} of {d.size} differ"
    return worst


def test_config1_single_stream_nbfm_10s(env, oracle):
    """BASELINE configs[0]: one NBFM stream, 10 s at 2.048 MS/s, fed in the reference's 262144-byte blocks
    (IqDataProcessor::acceptIqData block size) through the host-memory entry; 79 872 + 128 PCM samples."""
    torch, bench, capi, dev = env
    from hackrfdiags_b200 import synth
    n = 10 * FS
    iq = synth.rx_stream(capi.MODE_FM, n, stream=0, config=1)
    b = capi.Batch(1, capi.RX, 0)
    b.set_mode(capi.MODE_FM)
    parts = [b.rx(iq[None, off:off + 262144].copy()) for off in range(0, iq.size, 262144)]
    got = np.concatenate(parts, axis=1)[0]
    want = oracle.run_rx(capi.MODE_FM, iq)
    assert got.size == want.size == n // 256
    assert np.array_equal(got, want), f"{(got != want).sum()} of {want.size} PCM samples differ"


def test_config2_am_ssb_1024_streams(env, oracle):
    torch, bench, capi, dev = env
    n_samples = FS  # 1 s per stream: 4.19 GB of IQ
    groups = [(capi.MODE_AM, 512), (capi.MODE_LSB, 256), (capi.MODE_USB, 256)]
    b, iq, pcm, layout = bench.make_rx_batch(torch, capi, dev, groups, n_samples, seed=21)
    modes = [m for m, c in groups for _ in range(c)]
    _rx_call(capi, b, iq, pcm)
    torch.cuda.synchronize()
    # (1) every distinct stream (edge classes included) against the oracle, whole second; plus the last rows
    _check_sample_vs_oracle(torch, oracle, iq, pcm, modes, _distinct_rows(layout) + [511, 767, 1023])
    # (2) independence: every repeat of every distinct row
    assert bench.repeats_identical(torch, pcm, layout) == 0
    assert not torch.equal(pcm[512:544], pcm[768:800])  # LSB vs USB of mirrored signals differ
    # (3) automatic tiling == a single tile per stream
    b2 = capi.Batch(1024, capi.RX, 0)
    for s, m in enumerate(modes):
        b2.set_mode(m, s)
    b2.set_option(capi.OPT_RX_TILE_BATCHES, n_samples // 8192)
    pcm2 = torch.zeros_like(pcm)
    _rx_call(capi, b2, iq, pcm2)
    torch.cuda.synchronize()
    assert torch.equal(pcm, pcm2)
    # (4) one call == four calls
    b3 = capi.Batch(1024, capi.RX, 0)
    for s, m in enumerate(modes):
        b3.set_mode(m, s)
    pcm3 = torch.zeros_like(pcm)
    q = iq.shape[1] // 4
    for k in range(4):
        _rx_call(capi, b3, iq, pcm3, k * q, (k + 1) * q)
    torch.cuda.synchronize()
    assert torch.equal(pcm, pcm3)


def test_config3_wbfm_4096_streams(env, oracle):
    torch, bench, capi, dev = env
    n_samples = FS  # 1 s per stream: 16.8 GB of IQ
    b, iq, pcm, layout = bench.make_rx_batch(torch, capi, dev, [(capi.MODE_WBFM, 4096)], n_samples, seed=22)
    _rx_call(capi, b, iq, pcm)
    torch.cuda.synchronize()
    _check_sample_vs_oracle(torch, oracle, iq, pcm, [capi.MODE_WBFM] * 4096, _distinct_rows(layout) + [2048, 4095])
    assert bench.repeats_identical(torch, pcm, layout) == 0
    # one call == three uneven calls (the de-emphasis recurrence and the FIR histories carry over)
    b2 = capi.Batch(4096, capi.RX, 0)
    b2.set_mode(capi.MODE_WBFM)
    pcm2 = torch.zeros_like(pcm)
    cuts = [0, 8192 * 2 * 7, 8192 * 2 * 40, iq.shape[1]]
    for k in range(3):
        _rx_call(capi, b2, iq, pcm2, cuts[k], cuts[k + 1])
    torch.cuda.synchronize()
    assert torch.equal(pcm, pcm2)


def test_wbfm_whole_waves_plus_a_small_remainder(env, oracle):
    """WBFM launches are planned in two parts (hrd_api.cu plan_wbfm): the streams that fill whole waves of CTAs run
    untiled, the rest finely tiled in one more wave.  One wave plus 5 streams, and two waves plus 1, three calls in a
    row (state carried through both parts): every distinct row against the oracle, all repeats identical."""
    torch, bench, capi, dev = env
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    for n in (sms * 27 + 5, 2 * sms * 27 + 1):
        n_samples = 8192 * 12  # 12 batches per stream: 0.8 / 1.6 GB of IQ
        b, iq, pcm, layout = bench.make_rx_batch(torch, capi, dev, [(capi.MODE_WBFM, n)], n_samples, seed=29)
        cuts = [0, 8192 * 2 * 5, 8192 * 2 * 9, iq.shape[1]]
        for lo, hi in zip(cuts, cuts[1:]):
            _rx_call(capi, b, iq, pcm, lo, hi)
        torch.cuda.synchronize()
        _check_sample_vs_oracle(torch, oracle, iq, pcm, [capi.MODE_WBFM] * n, _distinct_rows(layout) + [sms * 27 - 1, sms * 27, n - 1])
        assert bench.repeats_identical(torch, pcm, layout) == 0
        assert b.wbfm_serial_count() == 0
        del b, iq, pcm
        torch.cuda.empty_cache()


def test_config5_mixed_modes_4096_streams(env, oracle):
    """BASELINE configs[4] at the bench's headline point: 4096 mixed-mode streams (1/4 AM, NBFM, WBFM, 1/8 LSB, USB)
    in one batch, 0.5 s each; every distinct stream of every mode against the oracle, every repeat on the device."""
    torch, bench, capi, dev = env
    from hackrfdiags_b200 import shard
    n_samples = FS // 2 // 8192 * 8192
    plan = shard.mixed_mode_plan(4096, bench.MIX)
    groups = shard.mode_groups(shard.shard_modes(plan, 1, 0))
    b, iq, pcm, layout = bench.make_rx_batch(torch, capi, dev, groups, n_samples, seed=24)
    modes = [m for m, c in groups for _ in range(c)]
    _rx_call(capi, b, iq, pcm)
    torch.cuda.synchronize()
    _check_sample_vs_oracle(torch, oracle, iq, pcm, modes, _distinct_rows(layout))
    assert bench.repeats_identical(torch, pcm, layout) == 0


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_config4_tx_4096_streams(env, oracle, mode):
    torch, bench, capi, dev = env
    n_pcm = 8000  # 1 s per stream -> 16.8 GB of IQ out
    distinct = bench.make_tx_pcm_device(torch, 32, n_pcm, dev, seed=23)
    pcm = bench.tile_rows(torch, distinct, 4096)
    iq = torch.empty((4096, n_pcm * 512), dtype=torch.int8, device=dev)
    b = capi.Batch(4096, capi.TX, 0)
    b.set_mode(mode)
    b.tx_device(pcm.data_ptr(), n_pcm, pcm.stride(0), iq.data_ptr(), iq.stride(0), 0)
    torch.cuda.synchronize()
    tol = 0  # FM too: the Nco::run head computes libm's cosf / sinf bit for bit (hrd_device.cuh glibc_sincosf)
    for s in list(range(32)) + [4095]:  # every distinct row (sines to -32768, noise, AM tone, silence, square wave)
        want = oracle.run_tx(mode, pcm[s].cpu().numpy())
        got = iq[s].cpu().numpy()
        err = np.abs(got.astype(np.int32) - want.astype(np.int32)).max()
        assert err <= tol, f"mode {mode} stream {s}: max abs err {err}"
    assert bench.repeats_identical(torch, iq, {mode: (0, 32, 4096)}) == 0
    # one call == two calls (interpolator histories and NCO phase carry over)
    b2 = capi.Batch(4096, capi.TX, 0)
    b2.set_mode(mode)
    iq2 = torch.empty_like(iq)
    cut = 777
    b2.tx_device(pcm.data_ptr(), cut, pcm.stride(0), iq2.data_ptr(), iq2.stride(0), 0)
    b2.tx_device(pcm.data_ptr() + 2 * cut, n_pcm - cut, pcm.stride(0), iq2.data_ptr() + cut * 512, iq2.stride(0), 0)
    torch.cuda.synchronize()
    assert torch.equal(iq, iq2)
    del iq, iq2
    torch.cuda.empty_cache()


def test_squelched_batch_1024_streams(env, oracle):
    """SURVEY 8f row 1 at batch size: 1024 mixed-mode streams x 8 transfer blocks, levels that cross a -40 dBFS
    threshold, per-stream thresholds and receive gains.  A sample of streams against the oracle (PCM, magnitudes,
    decisions); streams fed identical inputs and parameters agree wherever they sit; the counts add up."""
    torch, bench, capi, dev = env
    from hackrfdiags_b200 import synth
    n_streams, n_blocks = 1024, 8
    modes = [(capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB, capi.MODE_NONE)[s % 6] for s in range(n_streams)]
    thr = [(-40, -30, -200, -45)[(s // 6) % 4] for s in range(n_streams)]
    gain = [(16, 0, 24)[(s // 24) % 3] for s in range(n_streams)]
    distinct = [synth.rx_bursty_stream(capi.MODE_FM, n_blocks, stream=k) for k in range(8)]
    host = np.stack([distinct[(s // 72) % 8] for s in range(n_streams)])  # 72 = lcm of the parameter periods
    iq = torch.from_numpy(host).to(dev)
    pcm = torch.zeros((n_streams, n_blocks * 512), dtype=torch.int16, device=dev)
    b = capi.Batch(n_streams, capi.RX, 0)
    for s in range(n_streams):
        b.set_mode(modes[s], s)
        b.set_param(capi.PARAM_SQUELCH_THRESHOLD, thr[s], s)
        b.set_param(capi.PARAM_RX_GAIN_DB, gain[s], s)
    counts = np.zeros(n_streams, dtype=np.uint32)
    rc = b.lib.hrd_rx_process(b.h, iq.data_ptr(), iq.shape[1], iq.stride(0), capi.ENTRY_2048K, pcm.data_ptr(), pcm.stride(0),
                              counts.ctypes.data, capi.MEM_DEVICE, None)
    assert rc == 0
    torch.cuda.synchronize()
    got = pcm.cpu().numpy()
    mags, allowed = b.squelch_report()
    assert mags.shape == (n_streams, n_blocks)
    for s in range(n_streams):
        want_count = 0 if modes[s] == capi.MODE_NONE else 512 * int(allowed[s].sum())
        assert counts[s] == want_count, s
    # streams with the same input and parameters: 72 apart, same block of eight distinct inputs every 576
    for s in (0, 5, 17, 100):
        t = s + 576
        assert np.array_equal(allowed[s], allowed[t]) and np.array_equal(got[s], got[t]), (s, t)
    gated = 0
    for s in (0, 1, 2, 3, 4, 7, 30, 77, 500, 1023):
        want_pcm, want_mag, want_open = oracle.run_rx_squelch(modes[s], host[s], thr[s], gain[s])
        assert np.array_equal(mags[s], want_mag) and np.array_equal(allowed[s], want_open), s
        if modes[s] != capi.MODE_NONE:
            assert np.array_equal(got[s, :counts[s]], want_pcm), f"stream {s} mode {modes[s]}"
        gated += int((want_open == 0).sum())
    assert gated > 0
