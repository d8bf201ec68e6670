"""Batched ingest / egress adapters (hackrfdiags_b200/csrc/hrd_adapt.cc): host-side code,
so it runs without a GPU.  The PCM ring is checked event by event against the UNMODIFIED BasebandDataProcessor
(its private ring methods reached through oracle/ref_driver.cc) and against event traces that reference produced
(tests/golden/ring_traces.npz, for boxes without /root/reference); the block queue against DataConsumer's rules."""
import ctypes as C
import os

import numpy as np
import pytest

from cpu_checkers import Ref, have_ref
from hackrfdiags_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
TRACES = os.path.join(HERE, "golden", "ring_traces.npz")
needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference)")


def _lib():
    lib = capi.load()
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    lib.hrd_pcm_ring_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hrd_pcm_ring_destroy.argtypes = [vp]
    lib.hrd_pcm_ring_start.argtypes = [vp, C.c_int, C.c_int]
    lib.hrd_pcm_ring_write.argtypes = [vp, C.c_int, vp, C.c_uint32]
    lib.hrd_pcm_ring_read_all.argtypes = [vp, vp, C.c_size_t, i32p]
    lib.hrd_pcm_ring_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32)]
    lib.hrd_iq_queue_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hrd_iq_queue_destroy.argtypes = [vp]
    lib.hrd_iq_queue_push.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_uint32]
    lib.hrd_iq_queue_pop_all.argtypes = [vp, vp, C.c_size_t, vp, vp]
    lib.hrd_iq_queue_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32)]
    return lib


def make_events(seed, n=400):
    """W = the reader thread wrote a block, R = a transmit callback, S/T = start / stop streaming.  The write rate
    drifts above and below the read rate so that the ring drops and repeats blocks."""
    rng = np.random.default_rng(seed)
    ev = ["W"] * int(rng.integers(0, 12)) + ["S"]
    p_write = 0.5
    for k in range(n):
        if k % 60 == 0:
            p_write = float(rng.choice([0.35, 0.5, 0.65, 0.8, 0.2]))
        ev.append("W" if rng.random() < p_write else "R")
        if rng.random() < 0.01:
            ev.append("T" if rng.random() < 0.5 else "S")
    return "".join(ev)


def run_ours(lib, events, n_streams=1, stream=0):
    ring = C.c_void_p()
    assert lib.hrd_pcm_ring_create(n_streams, C.byref(ring)) == 0
    rows = np.zeros((n_streams, 512), dtype=np.int16)
    slots = np.zeros(n_streams, dtype=np.int32)
    out, seq = [], 0
    for e in events:
        if e == "W":
            blk = np.full(512, seq % 30000 + 1, dtype=np.int16)
            seq += 1
            assert lib.hrd_pcm_ring_write(ring, stream, blk.ctypes.data, 512) == 0
        elif e == "R":
            assert lib.hrd_pcm_ring_read_all(ring, rows.ctypes.data, 512, slots.ctypes.data_as(C.POINTER(C.c_int32))) == 0
            out.append((int(slots[stream]), int(rows[stream, 0]), int(rows[stream, 511])))
        else:
            lib.hrd_pcm_ring_start(ring, stream, 1 if e == "S" else 0)
    st = (C.c_uint32 * 4)()
    lib.hrd_pcm_ring_stats(ring, stream, st)
    lib.hrd_pcm_ring_destroy(ring)
    return np.array(out, dtype=np.int32).reshape(-1, 3), np.array(list(st), dtype=np.uint32)


def run_reference(events):
    ref = Ref().lib
    ref.ref_bbp_new.restype = C.c_void_p
    ref.ref_bbp_free.argtypes = [C.c_void_p]
    ref.ref_bbp_run.argtypes = [C.c_void_p, C.c_int]
    ref.ref_bbp_write.argtypes = [C.c_void_p, C.c_void_p]
    ref.ref_bbp_read.argtypes = [C.c_void_p, C.c_void_p]
    ref.ref_bbp_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    h = ref.ref_bbp_new()
    blk_out = np.zeros(512, dtype=np.int16)
    out, seq = [], 0
    for e in events:
        if e == "W":
            blk = np.full(512, seq % 30000 + 1, dtype=np.int16)
            seq += 1
            ref.ref_bbp_write(h, blk.ctypes.data)
        elif e == "R":
            slot = ref.ref_bbp_read(h, blk_out.ctypes.data)
            out.append((slot, int(blk_out[0]), int(blk_out[511])))
        else:
            ref.ref_bbp_run(h, 1 if e == "S" else 0)
    st = (C.c_uint32 * 4)()
    ref.ref_bbp_stats(h, st)
    ref.ref_bbp_free(h)
    return np.array(out, dtype=np.int32).reshape(-1, 3), np.array(list(st), dtype=np.uint32)


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_pcm_ring_matches_basebanddataprocessor(seed):
    events = make_events(seed)
    got, got_stats = run_ours(_lib(), events)
    want, want_stats = run_reference(events)
    assert np.array_equal(got, want), "slot / block sequence differs"
    assert np.array_equal(got_stats, want_stats), "produced / consumed / dropped / added differ"
    assert want_stats[2] + want_stats[3] > 0, "the trace never exercised the rate matching"


def test_pcm_ring_matches_recorded_reference_traces():
    t = np.load(TRACES)
    lib = _lib()
    for k in range(int(t["n"])):
        events = str(t[f"events_{k}"])
        got, got_stats = run_ours(lib, events, n_streams=3, stream=k % 3)  # other streams stay idle: zero blocks
        assert np.array_equal(got, t[f"out_{k}"]) and np.array_equal(got_stats, t[f"stats_{k}"]), k


def test_iq_queue_follows_dataconsumer_rules():
    """DataConsumer.cc:219-261: clamp to 262144, count short blocks, 16 slots round robin WITHOUT an overflow check
    (the 17th unconsumed block overwrites the first), FIFO order; a round needs a block of every stream."""
    lib = _lib()
    q = C.c_void_p()
    assert lib.hrd_iq_queue_create(2, C.byref(q)) == 0
    rows = np.zeros((2, 262144), dtype=np.int8)
    nbytes = np.zeros(2, dtype=np.uint32)
    stamps = np.zeros(2, dtype=np.uint32)

    def push(s, tag, n=262144):
        blk = np.full(max(n, 1), tag, dtype=np.int8)
        assert lib.hrd_iq_queue_push(q, s, 1000 + tag, blk.ctypes.data, n) == 0

    def pop():
        return lib.hrd_iq_queue_pop_all(q, rows.ctypes.data, 262144, nbytes.ctypes.data, stamps.ctypes.data)

    push(0, 1)
    assert pop() == 0, "stream 1 has nothing queued: no round"
    push(1, 2, n=1000)      # a short block
    push(0, 3, n=300000)    # clamped
    assert pop() == 1 and rows[0, 0] == 1 and rows[1, 0] == 2 and list(nbytes) == [262144, 1000] and list(stamps) == [1001, 1002]
    st = (C.c_uint32 * 3)()
    lib.hrd_iq_queue_stats(q, 1, st)
    assert list(st) == [0, 1, 1002]
    lib.hrd_iq_queue_stats(q, 0, st)
    assert list(st)[:2] == [1, 0]
    # overflow: 17 more blocks on stream 0 (one is already queued in slot 1): slot 1 is overwritten by tag 4 + 15
    for k in range(17):
        push(0, 4 + k)
    push(1, 50)
    assert pop() == 1 and rows[0, 0] == 4 + 15, "the overwritten slot is delivered with its NEW content, as in the reference"
    lib.hrd_iq_queue_destroy(q)


def test_iq_queue_push_rows_is_a_push_per_stream():
    lib = _lib()
    lib.hrd_iq_queue_push_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32]
    q = C.c_void_p()
    assert lib.hrd_iq_queue_create(5, C.byref(q)) == 0
    blocks = (np.arange(3 * 4096, dtype=np.int64) % 251).astype(np.int8).reshape(3, 4096)
    assert lib.hrd_iq_queue_push_rows(q, 1, 3, 77, blocks.ctypes.data, blocks.strides[0], 4096) == 0   # streams 1..3
    assert lib.hrd_iq_queue_push_rows(q, 3, 3, 78, blocks.ctypes.data, blocks.strides[0], 4096) == -1  # HRD_EINVAL: past the end
    one = np.full(4096, 9, dtype=np.int8)
    for s in (0, 4):
        assert lib.hrd_iq_queue_push(q, s, 77, one.ctypes.data, 4096) == 0
    rows = np.zeros((5, 262144), dtype=np.int8)
    nbytes = np.zeros(5, dtype=np.uint32)
    assert lib.hrd_iq_queue_pop_all(q, rows.ctypes.data, 262144, nbytes.ctypes.data, None) == 1
    assert list(nbytes) == [4096] * 5
    for i in range(3):
        assert np.array_equal(rows[1 + i, :4096], blocks[i])
    lib.hrd_iq_queue_destroy(q)


def test_library_exports_the_adapter_symbols():
    lib = capi.load()
    for name in ("hrd_pcm_ring_create", "hrd_pcm_ring_destroy", "hrd_pcm_ring_start", "hrd_pcm_ring_write", "hrd_pcm_ring_read_all",
                 "hrd_pcm_ring_stats", "hrd_tx_from_ring", "hrd_iq_queue_create", "hrd_iq_queue_destroy", "hrd_iq_queue_push", "hrd_iq_queue_push_rows",
                 "hrd_iq_queue_pop_all", "hrd_iq_queue_stats", "hrd_rx_from_queue"):
        assert hasattr(lib, name), name
