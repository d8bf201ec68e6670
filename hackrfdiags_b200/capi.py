"""ctypes binding of include/hrd.h (libhrd_b200.so) -- test and bench plumbing only.

The product is the shared library; this module just loads it and marshals numpy arrays or
raw device pointers (e.g. ``torch.Tensor.data_ptr()``) across the C ABI.  There is no
fallback: if the library is missing, or there is no sm_100 device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HRD_LIB: timing experiments load a variant build (tools/exp_build.sh); the product is libhrd_b200.so
LIB_PATH = os.environ.get("HRD_LIB") or os.path.join(_HERE, "libhrd_b200.so")

RX, TX = 0, 1
MODE_NONE, MODE_AM, MODE_FM, MODE_WBFM, MODE_LSB, MODE_USB = range(6)
MODE_IQ8K, MODE_DSB, MODE_PM, MODE_AM_PROTO, MODE_FM_PROTO = 6, 7, 8, 9, 10  # Tx only: the tool chain of signals/
(PARAM_AM_GAIN, PARAM_FM_GAIN, PARAM_WBFM_GAIN, PARAM_SSB_GAIN,
 PARAM_AM_INDEX, PARAM_FM_DEV, PARAM_WBFM_DEV, PARAM_SQUELCH_THRESHOLD, PARAM_RX_GAIN_DB) = range(9)
UNIT_AM, UNIT_FM, UNIT_WBFM, UNIT_SSB, UNIT_FRONT_END, UNIT_ALL, UNIT_SIGNALS = range(7)
ENTRY_2048K, ENTRY_256K = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
ALL_STREAMS = -1
(OPT_RX_TILE_BATCHES, OPT_RX_WBFM_TILING, OPT_TX_TILE_SAMPLES, OPT_PROFILE, OPT_DEBUG_WBFM_FORCE_RERUN, OPT_RX_SQUELCH,
 OPT_RX_SQUELCH_BLOCK, OPT_RX_SERIAL, OPT_RX_WBFM_PACK) = range(9)

# every symbol include/hrd.h declares (tests check that the library exports all of them)
EXPORTS = [
    "hrd_abi_version", "hrd_create", "hrd_destroy", "hrd_last_error", "hrd_set_mode", "hrd_get_mode",
    "hrd_set_param", "hrd_get_param", "hrd_reset", "hrd_set_option", "hrd_get_option", "hrd_rx_process", "hrd_rx_front_end", "hrd_rx_squelch_report", "hrd_rx_fs4_rotate", "hrd_tx_process",
    "hrd_pcm_ring_create", "hrd_pcm_ring_destroy", "hrd_pcm_ring_start", "hrd_pcm_ring_write", "hrd_pcm_ring_read_all",
    "hrd_pcm_ring_stats", "hrd_tx_from_ring", "hrd_iq_queue_create", "hrd_iq_queue_destroy", "hrd_iq_queue_push", "hrd_iq_queue_push_rows",
    "hrd_iq_queue_pop_all", "hrd_iq_queue_stats", "hrd_rx_from_queue",
    "hrd_rx_pipe_create", "hrd_rx_pipe_destroy", "hrd_rx_pipe_submit", "hrd_rx_pipe_collect", "hrd_rx_pipe_stats",
    "hrd_tx_pipe_create", "hrd_tx_pipe_destroy", "hrd_tx_pipe_submit", "hrd_tx_pipe_collect",
    "hrd_sharded_create", "hrd_sharded_destroy", "hrd_sharded_count", "hrd_sharded_shard", "hrd_sharded_last_error",
    "hrd_sharded_set_mode", "hrd_sharded_set_param", "hrd_sharded_reset", "hrd_sharded_set_option",
    "hrd_sharded_rx_process", "hrd_sharded_tx_process",
    "hrd_synchronize", "hrd_get_device", "hrd_launch_count", "hrd_wbfm_fallback_count", "hrd_wbfm_serial_count", "hrd_kernel_ms", "hrd_get_table", "hrd_get_taps", "hrd_state_bytes_per_stream",
]


class HrdError(RuntimeError):
    pass


_lib = None


def load():
    """Load libhrd_b200.so; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HrdError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    lib.hrd_abi_version.restype = i
    lib.hrd_last_error.restype = C.c_char_p
    lib.hrd_create.argtypes = [i, i, i, C.POINTER(vp)]
    lib.hrd_destroy.argtypes = [vp]
    lib.hrd_set_mode.argtypes = [vp, i, i]
    lib.hrd_get_mode.argtypes = [vp, i, C.POINTER(i)]
    lib.hrd_set_param.argtypes = [vp, i, i, C.c_float]
    lib.hrd_get_param.argtypes = [vp, i, i, C.POINTER(C.c_float)]
    lib.hrd_reset.argtypes = [vp, i, i]
    lib.hrd_set_option.argtypes = [vp, i, i]
    lib.hrd_get_option.argtypes = [vp, i, C.POINTER(i)]
    lib.hrd_rx_process.argtypes = [vp, vp, sz, sz, i, vp, sz, vp, i, vp]
    lib.hrd_rx_front_end.argtypes = [vp, vp, sz, sz, vp, sz, i, vp]
    lib.hrd_tx_process.argtypes = [vp, vp, sz, sz, vp, sz, i, vp]
    lib.hrd_rx_squelch_report.argtypes = [vp, vp, vp, sz, C.POINTER(C.c_uint32)]
    lib.hrd_rx_fs4_rotate.argtypes = [vp, vp, sz, i, i, vp]
    lib.hrd_synchronize.argtypes = [vp]
    lib.hrd_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.hrd_kernel_ms.argtypes = [vp, i, i, C.POINTER(C.c_float)]
    lib.hrd_wbfm_fallback_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.hrd_wbfm_serial_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.hrd_get_table.argtypes = [vp, i, vp, sz]
    lib.hrd_get_taps.argtypes = [i, vp, i]
    lib.hrd_state_bytes_per_stream.argtypes = [i]
    lib.hrd_state_bytes_per_stream.restype = sz
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise HrdError(f"hrd error {rc}: {load().hrd_last_error().decode()}")


def get_taps(which: int) -> np.ndarray:
    out = np.zeros(64, dtype=np.int16)
    n = load().hrd_get_taps(which, out.ctypes.data, 64)
    if n < 0:
        raise HrdError(load().hrd_last_error().decode())
    return out[:n].copy()


class Batch:
    """One hrd_batch_t: n_streams independent reference object graphs on one GPU."""

    def __init__(self, n_streams: int, kind: int, device: int = 0):
        self.lib = load()
        self.n = int(n_streams)
        self.kind = kind
        self.h = C.c_void_p()
        _check(self.lib.hrd_create(int(device), self.n, int(kind), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.hrd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- control ----
    def set_mode(self, mode: int, stream: int = ALL_STREAMS):
        _check(self.lib.hrd_set_mode(self.h, int(stream), int(mode)))

    def get_mode(self, stream: int) -> int:
        m = C.c_int()
        _check(self.lib.hrd_get_mode(self.h, int(stream), C.byref(m)))
        return m.value

    def set_param(self, param: int, value: float, stream: int = ALL_STREAMS):
        _check(self.lib.hrd_set_param(self.h, int(stream), int(param), float(value)))

    def get_param(self, param: int, stream: int) -> float:
        v = C.c_float()
        _check(self.lib.hrd_get_param(self.h, int(stream), int(param), C.byref(v)))
        return v.value

    def reset(self, unit: int = UNIT_ALL, stream: int = ALL_STREAMS):
        _check(self.lib.hrd_reset(self.h, int(stream), int(unit)))

    def set_option(self, option: int, value: int):
        _check(self.lib.hrd_set_option(self.h, int(option), int(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int()
        _check(self.lib.hrd_get_option(self.h, int(option), C.byref(v)))
        return v.value

    def synchronize(self):
        _check(self.lib.hrd_synchronize(self.h))

    def launch_count(self) -> int:
        c = C.c_uint64()
        _check(self.lib.hrd_launch_count(self.h, C.byref(c)))
        return c.value

    def wbfm_fallback_count(self) -> int:
        c = C.c_uint64()
        _check(self.lib.hrd_wbfm_fallback_count(self.h, C.byref(c)))
        return c.value

    def wbfm_serial_count(self) -> int:
        c = C.c_uint64()
        _check(self.lib.hrd_wbfm_serial_count(self.h, C.byref(c)))
        return c.value

    def kernel_ms(self, which: int = 0, age: int = 0) -> float:
        v = C.c_float()
        _check(self.lib.hrd_kernel_ms(self.h, int(which), int(age), C.byref(v)))
        return v.value

    def get_table(self, which: int) -> np.ndarray:
        n = 65536 if which == 0 else 16384
        out = np.zeros(n, dtype=np.float32)
        _check(self.lib.hrd_get_table(self.h, which, out.ctypes.data, n))
        return out

    # ---- host-memory calls (numpy) ----
    def rx(self, iq: np.ndarray, entry: int = ENTRY_2048K) -> np.ndarray:
        """iq[n_streams, bytes] int8 (C-contiguous rows) -> pcm[n_streams, n_pcm] int16."""
        assert iq.dtype == np.int8 and iq.ndim == 2 and iq.shape[0] == self.n and iq.strides[1] == 1
        nbytes = iq.shape[1]
        npcm = nbytes // (512 if entry == ENTRY_2048K else 64)
        pcm = np.zeros((self.n, max(npcm, 1)), dtype=np.int16)
        counts = np.zeros(self.n, dtype=np.uint32)
        _check(self.lib.hrd_rx_process(self.h, iq.ctypes.data, nbytes, iq.strides[0], entry, pcm.ctypes.data,
                                       pcm.shape[1], counts.ctypes.data, MEM_HOST, None))
        self.last_counts = counts
        return pcm[:, :npcm]

    def squelch_report(self):
        """(magnitudes[n_streams, n_blocks] uint32, allowed[n_streams, n_blocks] uint8) of the latest squelched rx call."""
        nb = C.c_uint32(0)
        _check(self.lib.hrd_rx_squelch_report(self.h, None, None, 0, C.byref(nb)))
        mags = np.zeros((self.n, max(nb.value, 1)), dtype=np.uint32)
        allowed = np.zeros((self.n, max(nb.value, 1)), dtype=np.uint8)
        _check(self.lib.hrd_rx_squelch_report(self.h, mags.ctypes.data, allowed.ctypes.data, mags.shape[1], C.byref(nb)))
        return mags[:, :nb.value], allowed[:, :nb.value]

    def rx_front_end(self, iq: np.ndarray) -> np.ndarray:
        assert iq.dtype == np.int8 and iq.ndim == 2 and iq.shape[0] == self.n and iq.strides[1] == 1
        nbytes = iq.shape[1]
        out = np.zeros((self.n, max(nbytes // 8, 1)), dtype=np.int8)
        _check(self.lib.hrd_rx_front_end(self.h, iq.ctypes.data, nbytes, iq.strides[0], out.ctypes.data,
                                         out.shape[1], MEM_HOST, None))
        return out[:, :nbytes // 8]

    def tx(self, pcm: np.ndarray, n: int | None = None) -> np.ndarray:
        """pcm[n_streams, n] int16 -> iq[n_streams, n*512] int8.  With HRD_MODE_IQ8K streams in the batch the rows
        hold 2*n int16 (I,Q pairs for those streams, PCM in the first n for the others): pass n."""
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.shape[0] == self.n and pcm.strides[1] == 2
        n = pcm.shape[1] if n is None else n
        iq = np.zeros((self.n, max(n * 512, 32)), dtype=np.int8)
        _check(self.lib.hrd_tx_process(self.h, pcm.ctypes.data, n, pcm.strides[0] // 2, iq.ctypes.data,
                                       iq.shape[1], MEM_HOST, None))
        return iq[:, :n * 512]

    # ---- device-memory calls (raw pointers, e.g. torch data_ptr()) ----
    def rx_device(self, iq_ptr: int, nbytes: int, iq_stride: int, pcm_ptr: int, pcm_stride: int,
                  entry: int = ENTRY_2048K, cuda_stream: int = 0):
        _check(self.lib.hrd_rx_process(self.h, iq_ptr, nbytes, iq_stride, entry, pcm_ptr, pcm_stride, None,
                                       MEM_DEVICE, cuda_stream or None))

    def tx_device(self, pcm_ptr: int, n: int, pcm_stride: int, iq_ptr: int, iq_stride: int, cuda_stream: int = 0):
        _check(self.lib.hrd_tx_process(self.h, pcm_ptr, n, pcm_stride, iq_ptr, iq_stride, MEM_DEVICE,
                                       cuda_stream or None))

    def rx_host_ptr(self, iq_ptr: int, nbytes: int, iq_stride: int, pcm_ptr: int, pcm_stride: int,
                    entry: int = ENTRY_2048K):
        """Host pointers (e.g. pinned torch tensors): copies happen inside the call."""
        _check(self.lib.hrd_rx_process(self.h, iq_ptr, nbytes, iq_stride, entry, pcm_ptr, pcm_stride, None,
                                       MEM_HOST, None))

    def tx_host_ptr(self, pcm_ptr: int, n: int, pcm_stride: int, iq_ptr: int, iq_stride: int):
        _check(self.lib.hrd_tx_process(self.h, pcm_ptr, n, pcm_stride, iq_ptr, iq_stride, MEM_HOST, None))
