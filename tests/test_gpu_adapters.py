"""The adapters end to end on the GPU: blocks pushed per stream into the DataConsumer-style
queue come out of hrd_rx_from_queue as the PCM the oracle gives for the same blocks; blocks written into the PCM rings
go out through hrd_tx_from_ring as the IQ the oracle's modulators give for the block sequence the ring policy picked."""
import ctypes as C

import numpy as np
import pytest

from cpu_checkers import Oracle
from hackrfdiags_b200 import capi, synth
from test_adapters_cpu import _lib

pytestmark = pytest.mark.gpu


def test_rx_rounds_from_the_queue():
    lib, oracle = _lib(), Oracle()
    lib.hrd_rx_from_queue.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    modes = [capi.MODE_AM, capi.MODE_FM, capi.MODE_USB]
    n_blocks = 3
    iq = [synth.rx_stream(m, n_blocks * 131072, stream=i, config=14) for i, m in enumerate(modes)]
    b = capi.Batch(3, capi.RX, 0)
    for i, m in enumerate(modes):
        b.set_mode(m, i)
    q = C.c_void_p()
    assert lib.hrd_iq_queue_create(3, C.byref(q)) == 0
    pcm = np.zeros((3, 512), dtype=np.int16)
    counts = np.zeros(3, dtype=np.uint32)
    got = [[] for _ in modes]
    order = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (2, 1), (2, 2), (0, 2), (1, 2)]  # producers run at their own pace
    for s, k in order:
        blk = iq[s][k * 262144:(k + 1) * 262144]
        assert lib.hrd_iq_queue_push(q, s, k, blk.ctypes.data, blk.size) == 0
        while True:
            rc = lib.hrd_rx_from_queue(b.h, q, pcm.ctypes.data, 512, counts.ctypes.data)
            assert rc >= 0
            if rc == 0:
                break
            for i in range(3):
                got[i].append(pcm[i, :counts[i]].copy())
    for i, m in enumerate(modes):
        assert np.array_equal(np.concatenate(got[i]), oracle.run_rx(m, iq[i])), f"stream {i}"
    lib.hrd_iq_queue_destroy(q)


def test_tx_blocks_from_the_rings():
    lib, oracle = _lib(), Oracle()
    lib.hrd_tx_from_ring.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    modes = [capi.MODE_AM, capi.MODE_LSB]
    b = capi.Batch(2, capi.TX, 0)
    for i, m in enumerate(modes):
        b.set_mode(m, i)
    ring = C.c_void_p()
    assert lib.hrd_pcm_ring_create(2, C.byref(ring)) == 0
    shadow = C.c_void_p()  # a second ring fed the same events tells which blocks the policy sends
    assert lib.hrd_pcm_ring_create(2, C.byref(shadow)) == 0
    for r in (ring, shadow):
        lib.hrd_pcm_ring_start(r, -1, 1)
    pcm = [synth.tx_stream(40 * 512, stream=i, config=15) for i in range(2)]
    iq = np.zeros((2, 262144), dtype=np.int8)
    rows = np.zeros((2, 512), dtype=np.int16)
    sent, out = [[], []], [[], []]
    rng = np.random.default_rng(3)
    k = [0, 0]
    for _ in range(60):
        for s in range(2):
            for _w in range(int(rng.integers(0, 3))):  # 0..2 blocks arrive per transfer: the ring repeats and drops
                if k[s] < 40:
                    blk = pcm[s][k[s] * 512:(k[s] + 1) * 512]
                    for r in (ring, shadow):
                        lib.hrd_pcm_ring_write(r, s, blk.ctypes.data, 512)
                    k[s] += 1
        lib.hrd_pcm_ring_read_all(shadow, rows.ctypes.data, 512, None)
        assert lib.hrd_tx_from_ring(b.h, ring, iq.ctypes.data, 262144, capi.MEM_HOST, None) == 0
        for s in range(2):
            sent[s].append(rows[s].copy())
            out[s].append(iq[s].copy())
    for s, m in enumerate(modes):
        want = oracle.run_tx(m, np.concatenate(sent[s]))
        assert np.array_equal(np.concatenate(out[s]), want), f"stream {s}"
    st = (C.c_uint32 * 4)()
    lib.hrd_pcm_ring_stats(ring, 0, st)
    assert st[2] + st[3] > 0, "the rate matching never acted"
    lib.hrd_pcm_ring_destroy(ring)
    lib.hrd_pcm_ring_destroy(shadow)
