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


def test_rx_pipe_1024_streams_with_a_producer_thread():
    """The adapters at rate: 1024 streams x 32 transfer blocks (2.1 s of signal each) pushed by a
    producer thread into the page-locked pool while the consumer keeps three rounds in flight (H2D of round k+1
    beside the kernels of round k and the PCM copy of round k-1).  Every stream's PCM, every round, against the oracle."""
    import threading
    lib, oracle = capi.load(), Oracle()
    vp = C.c_void_p
    lib.hrd_rx_pipe_create.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    lib.hrd_rx_pipe_destroy.argtypes = [vp]
    lib.hrd_rx_pipe_submit.argtypes = [vp]
    lib.hrd_rx_pipe_collect.argtypes = [vp, C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.c_size_t), C.POINTER(C.POINTER(C.c_uint32))]
    lib.hrd_rx_pipe_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.hrd_iq_queue_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hrd_iq_queue_push.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_uint32]
    lib.hrd_iq_queue_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32)]
    lib.hrd_iq_queue_destroy.argtypes = [vp]
    n, rounds, distinct = 1024, 32, 32
    modes = [(capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB)[i % 5] for i in range(distinct)]
    rows = [synth.rx_stream(m, rounds * 131072, stream=i, config=16) for i, m in enumerate(modes)]
    want = np.stack([oracle.run_rx(m, r) for m, r in zip(modes, rows)])  # [distinct, rounds * 512]
    b = capi.Batch(n, capi.RX, 0)
    for s in range(n):
        b.set_mode(modes[s % distinct], s)
    q, pipe = vp(), vp()
    assert lib.hrd_iq_queue_create(n, C.byref(q)) == 0
    assert lib.hrd_rx_pipe_create(b.h, q, 3, C.byref(pipe)) == 0
    failed = []

    def producer():
        st = (C.c_uint32 * 3)()
        for k in range(rounds):
            while True:  # the reference has no overflow check: the producer must not lap the consumer
                lib.hrd_iq_queue_stats(q, n - 1, st)
                if st[0] < 8:
                    break
            for s in range(n):
                blk = rows[s % distinct][k * 262144:(k + 1) * 262144]
                if lib.hrd_iq_queue_push(q, s, k, blk.ctypes.data, blk.size) != 0:
                    failed.append((k, s))

    t = threading.Thread(target=producer)
    t.start()
    pcm_p, stride, cnt_p = C.POINTER(C.c_int16)(), C.c_size_t(), C.POINTER(C.c_uint32)()
    done = 0
    expect = np.tile(want, (n // distinct, 1))  # row s -> distinct row s % distinct
    while done < rounds:
        while lib.hrd_rx_pipe_submit(pipe) == 1:
            pass
        rc = lib.hrd_rx_pipe_collect(pipe, C.byref(pcm_p), C.byref(stride), C.byref(cnt_p))
        assert rc >= 0
        if rc == 0:
            continue
        pcm = np.ctypeslib.as_array(pcm_p, shape=(n, stride.value))
        counts = np.ctypeslib.as_array(cnt_p, shape=(n,))
        assert (counts == 512).all()
        assert np.array_equal(pcm[:, :512], expect[:, done * 512:(done + 1) * 512]), f"round {done}"
        done += 1
    t.join()
    assert not failed
    stats = (C.c_uint64 * 2)()
    lib.hrd_rx_pipe_stats(pipe, stats)
    assert stats[0] == rounds
    assert stats[1] <= 2 * rounds, f"{stats[1]} copies for {rounds} rounds: the streams moved in step, so about one per round"
    lib.hrd_rx_pipe_destroy(pipe)
    lib.hrd_iq_queue_destroy(q)


def test_tx_pipe_matches_the_plain_path():
    """hrd_tx_pipe_*: the same block sequence through the pipelined path and through hrd_tx_from_ring."""
    lib = capi.load()
    vp = C.c_void_p
    lib.hrd_tx_pipe_create.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    lib.hrd_tx_pipe_destroy.argtypes = [vp]
    lib.hrd_tx_pipe_submit.argtypes = [vp]
    lib.hrd_tx_pipe_collect.argtypes = [vp, C.POINTER(C.POINTER(C.c_int8)), C.POINTER(C.c_size_t)]
    lib.hrd_tx_from_ring.argtypes = [vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.hrd_pcm_ring_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hrd_pcm_ring_start.argtypes = [vp, C.c_int, C.c_int]
    lib.hrd_pcm_ring_write.argtypes = [vp, C.c_int, vp, C.c_uint32]
    lib.hrd_pcm_ring_destroy.argtypes = [vp]
    n, rounds = 8, 12
    modes = [1, 2, 3, 4, 5, 1, 3, 2]
    pcm = [synth.tx_stream(rounds * 512, stream=i, config=17) for i in range(n)]
    outs = []
    for piped in (False, True):
        b = capi.Batch(n, capi.TX, 0)
        for i, m in enumerate(modes):
            b.set_mode(m, i)
        ring = vp()
        assert lib.hrd_pcm_ring_create(n, C.byref(ring)) == 0
        lib.hrd_pcm_ring_start(ring, -1, 1)
        for k in range(8):  # the reader starts eight blocks behind the writer
            for s in range(n):
                blk = pcm[s][k * 512:(k + 1) * 512]
                lib.hrd_pcm_ring_write(ring, s, blk.ctypes.data, 512)
        got = []
        if piped:
            pipe = vp()
            assert lib.hrd_tx_pipe_create(b.h, ring, 2, C.byref(pipe)) == 0
            iq_p, stride = C.POINTER(C.c_int8)(), C.c_size_t()
            for k in range(8, rounds):
                for s in range(n):
                    blk = pcm[s][k * 512:(k + 1) * 512]
                    lib.hrd_pcm_ring_write(ring, s, blk.ctypes.data, 512)
                assert lib.hrd_tx_pipe_submit(pipe) == 1
                if k % 2:  # two rounds in flight
                    for _ in range(2):
                        assert lib.hrd_tx_pipe_collect(pipe, C.byref(iq_p), C.byref(stride)) == 1
                        got.append(np.ctypeslib.as_array(iq_p, shape=(n, stride.value)).copy())
            lib.hrd_tx_pipe_destroy(pipe)
        else:
            iq = np.zeros((n, 262144), dtype=np.int8)
            for k in range(8, rounds):
                for s in range(n):
                    blk = pcm[s][k * 512:(k + 1) * 512]
                    lib.hrd_pcm_ring_write(ring, s, blk.ctypes.data, 512)
                assert lib.hrd_tx_from_ring(b.h, ring, iq.ctypes.data, 262144, capi.MEM_HOST, None) == 0
                got.append(iq.copy())
        outs.append(np.stack(got))
        lib.hrd_pcm_ring_destroy(ring)
    assert outs[0].shape == outs[1].shape and np.array_equal(outs[0], outs[1])
