"""The multi-GPU host entry (hackrfdiags_b200/csrc/hrd_shard.cc): one job of N streams dealt to
G shards -- contiguous ranges, one worker thread and one batch per shard, nothing crossing GPUs -- gives what one
batch gives.  On a one-GPU box the shards share device 0 (the partition, the global stream numbers and the
threading are what is under test); with two or more GPUs the same job also runs across devices 0 and 1 from this
one process, WBFM included (its kernels opt in to ~200 KB of shared memory PER DEVICE)."""
import ctypes as C

import numpy as np
import pytest

from cpu_checkers import Oracle
from hackrfdiags_b200 import capi, shard, synth

pytestmark = pytest.mark.gpu

RX_MODES = [capi.MODE_AM, capi.MODE_FM, capi.MODE_WBFM, capi.MODE_LSB, capi.MODE_USB]


def _lib():
    lib = capi.load()
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.hrd_sharded_create.argtypes = [C.POINTER(i), i, i, i, C.POINTER(vp)]
    lib.hrd_sharded_destroy.argtypes = [vp]
    lib.hrd_sharded_count.argtypes = [vp]
    lib.hrd_sharded_shard.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(vp)]
    lib.hrd_sharded_last_error.argtypes = [vp]
    lib.hrd_sharded_last_error.restype = C.c_char_p
    lib.hrd_sharded_set_mode.argtypes = [vp, i, i]
    lib.hrd_sharded_set_param.argtypes = [vp, i, i, C.c_float]
    lib.hrd_sharded_reset.argtypes = [vp, i, i]
    lib.hrd_sharded_rx_process.argtypes = [vp, vp, sz, sz, i, vp, sz, vp]
    lib.hrd_sharded_tx_process.argtypes = [vp, vp, sz, sz, vp, sz]
    return lib


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [g % have for g in range(n)]


def _run_rx(lib, devices, modes, gains, iq):
    n = len(modes)
    h = C.c_void_p()
    dev = (C.c_int * len(devices))(*devices)
    assert lib.hrd_sharded_create(dev, len(devices), n, capi.RX, C.byref(h)) == 0, capi.load().hrd_last_error()
    for s in range(n):
        assert lib.hrd_sharded_set_mode(h, s, modes[s]) == 0
        assert lib.hrd_sharded_set_param(h, s, capi.PARAM_AM_GAIN + {1: 0, 2: 1, 3: 2, 4: 3, 5: 3}[modes[s]], gains[s]) == 0
    out = []
    cut = iq.shape[1] // 2 // 512 * 512
    for lo, hi in ((0, cut), (cut, iq.shape[1])):  # two calls: the state lives in the shards' batches
        part = np.ascontiguousarray(iq[:, lo:hi])
        pcm = np.zeros((n, (hi - lo) // 512), dtype=np.int16)
        counts = np.zeros(n, dtype=np.uint32)
        rc = lib.hrd_sharded_rx_process(h, part.ctypes.data, part.shape[1], part.strides[0], capi.ENTRY_2048K, pcm.ctypes.data,
                                        pcm.shape[1], counts.ctypes.data)
        assert rc == 0, lib.hrd_sharded_last_error(h)
        assert (counts == pcm.shape[1]).all()
        out.append(pcm)
    # the ranges are shard.py's
    for g in range(lib.hrd_sharded_count(h)):
        d, lo, hi = C.c_int(), C.c_int(), C.c_int()
        assert lib.hrd_sharded_shard(h, g, C.byref(d), C.byref(lo), C.byref(hi), None) == 0
        assert (lo.value, hi.value) == shard.shard_range(n, len(devices), g) and d.value == devices[g]
    lib.hrd_sharded_destroy(h)
    return np.concatenate(out, axis=1)


@pytest.mark.parametrize("n_shards", [1, 2, 3])
def test_sharded_rx_equals_the_oracle(n_shards):
    lib, oracle = _lib(), Oracle()
    n = 11
    modes = [RX_MODES[s % 5] for s in range(n)]
    gains = [300.0 if s % 3 else 120.0 for s in range(n)]
    iq = np.stack([synth.rx_stream(m, 3 * 131072, stream=s, config=18) for s, m in enumerate(modes)])
    got = _run_rx(lib, _devices(n_shards), modes, gains, iq)
    for s, m in enumerate(modes):
        want = oracle.run_rx(m, iq[s], gain=gains[s])
        assert np.array_equal(got[s], want), f"stream {s} (mode {m}) with {n_shards} shards"


def test_sharded_tx_equals_one_batch():
    lib = _lib()
    n = 9
    modes = [(1, 2, 3, 4, 5)[s % 5] for s in range(n)]
    pcm = np.stack([synth.tx_stream(700, stream=s, config=19) for s in range(n)])
    one = capi.Batch(n, capi.TX, 0)
    for s, m in enumerate(modes):
        one.set_mode(m, s)
    want = one.tx(pcm)
    h = C.c_void_p()
    devices = _devices(2)
    dev = (C.c_int * 2)(*devices)
    assert lib.hrd_sharded_create(dev, 2, n, capi.TX, C.byref(h)) == 0
    for s, m in enumerate(modes):
        assert lib.hrd_sharded_set_mode(h, s, m) == 0
    iq = np.zeros((n, 700 * 512), dtype=np.int8)
    assert lib.hrd_sharded_tx_process(h, pcm.ctypes.data, 700, pcm.strides[0] // 2, iq.ctypes.data, iq.strides[0]) == 0, \
        lib.hrd_sharded_last_error(h)
    lib.hrd_sharded_destroy(h)
    assert np.array_equal(iq, want)


def test_two_devices_from_one_process_wbfm():
    """Batches on devices 0 and 1 in ONE process, WBFM both ways: the big-shared-memory opt-in is per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    oracle = Oracle()
    iq = np.stack([synth.rx_stream(capi.MODE_WBFM, 2 * 131072, stream=s, config=20) for s in range(3)])
    pcm = np.stack([synth.tx_stream(300, stream=s, config=21) for s in range(3)])
    for device in (0, 1, 0):
        rx = capi.Batch(3, capi.RX, device)
        rx.set_mode(capi.MODE_WBFM)
        got = rx.rx(iq)
        tx = capi.Batch(3, capi.TX, device)
        tx.set_mode(capi.MODE_WBFM)
        out = tx.tx(pcm)
        for s in range(3):
            assert np.array_equal(got[s], oracle.run_rx(capi.MODE_WBFM, iq[s])), f"rx device {device} stream {s}"
            assert np.array_equal(out[s], oracle.run_tx(capi.MODE_WBFM, pcm[s])), f"tx device {device} stream {s}"


def test_pipe_on_the_second_device():
    """A pipelined ingest on device 1 driven from a thread whose current device is 0: the pipe's streams, buffers
    and events must live where the batch lives."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    lib, oracle = capi.load(), Oracle()
    vp = C.c_void_p
    lib.hrd_rx_pipe_create.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    lib.hrd_rx_pipe_destroy.argtypes = [vp]
    lib.hrd_rx_pipe_submit.argtypes = [vp]
    lib.hrd_rx_pipe_collect.argtypes = [vp, C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.c_size_t), C.POINTER(C.POINTER(C.c_uint32))]
    lib.hrd_iq_queue_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hrd_iq_queue_push.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_uint32]
    lib.hrd_iq_queue_destroy.argtypes = [vp]
    lib.hrd_get_device.argtypes = [vp, C.POINTER(C.c_int)]
    n, rounds = 5, 3
    rows = [synth.rx_stream(RX_MODES[s], rounds * 131072, stream=s, config=22) for s in range(n)]
    b = capi.Batch(n, capi.RX, 1)
    d = C.c_int(-1)
    assert lib.hrd_get_device(b.h, C.byref(d)) == 0 and d.value == 1
    for s in range(n):
        b.set_mode(RX_MODES[s], s)
    q, pipe = vp(), vp()
    assert lib.hrd_iq_queue_create(n, C.byref(q)) == 0
    assert lib.hrd_rx_pipe_create(b.h, q, 2, C.byref(pipe)) == 0
    got = [[] for _ in range(n)]
    pcm_p, stride, cnt_p = C.POINTER(C.c_int16)(), C.c_size_t(), C.POINTER(C.c_uint32)()
    for k in range(rounds):
        for s in range(n):
            blk = rows[s][k * 262144:(k + 1) * 262144]
            assert lib.hrd_iq_queue_push(q, s, k, blk.ctypes.data, blk.size) == 0
        assert lib.hrd_rx_pipe_submit(pipe) == 1
        assert lib.hrd_rx_pipe_collect(pipe, C.byref(pcm_p), C.byref(stride), C.byref(cnt_p)) == 1
        pcm = np.ctypeslib.as_array(pcm_p, shape=(n, stride.value))
        for s in range(n):
            got[s].append(pcm[s, :512].copy())
    lib.hrd_rx_pipe_destroy(pipe)
    lib.hrd_iq_queue_destroy(q)
    for s in range(n):
        assert np.array_equal(np.concatenate(got[s]), oracle.run_rx(RX_MODES[s], rows[s])), f"stream {s}"
