"""ctypes bindings for the two CPU CHECKERS (test infrastructure, never the product).

* ``Oracle``  -> oracle/liboracle.so      (our C restatement, hrd_oracle.c)
* ``Ref``     -> oracle/_ref/libhrd_ref.so (the unmodified reference sources,
                 compiled by oracle/Makefile where /root/reference exists; the
                 built library travels to the GPU box, the sources do not)

Both expose the same Python surface so tests can swap them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libhrd_ref.so")

NONE, AM, FM, WBFM, LSB, USB = range(6)
MODE_NAMES = {AM: "am", FM: "fm", WBFM: "wbfm", LSB: "lsb", USB: "usb"}
DEMOD_OF_MODE = {AM: 0, FM: 1, WBFM: 2, LSB: 3, USB: 3}
MOD_OF_MODE = {AM: 0, FM: 1, WBFM: 2, LSB: 3, USB: 3}

TAPS = ["fe1", "fe2", "fe3", "am1", "am2", "am3", "fm_tuner", "fm_post", "audio40",
        "wbfm_post1", "ssb_delay", "ssb_hilbert", "tx_hb8"]

_i8p = C.POINTER(C.c_int8)
_i16p = C.POINTER(C.c_int16)
_f32p = C.POINTER(C.c_float)


def build_checkers() -> None:
    """(Re)build liboracle.so and, when the reference tree is present, _ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class _Base:
    prefix = ""
    lib = None

    # ---- rx ----
    def rx_new(self):
        return C.c_void_p(getattr(self.lib, self.prefix + "rx_new")())

    def rx_free(self, h):
        getattr(self.lib, self.prefix + "rx_free")(h)

    def rx_set_mode(self, h, mode):
        getattr(self.lib, self.prefix + "rx_set_mode")(h, int(mode))

    def rx_set_gain(self, h, demod, gain):
        getattr(self.lib, self.prefix + "rx_set_gain")(h, int(demod), C.c_float(gain))

    def rx_reset_demod(self, h, demod):
        getattr(self.lib, self.prefix + "rx_reset_demod")(h, int(demod))

    def rx_front_end(self, h, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, dtype=np.int8)
        out = np.zeros(iq.size // 8 + 16, dtype=np.int8)
        n = getattr(self.lib, self.prefix + "rx_front_end")(h, _ptr(iq, _i8p), iq.size, _ptr(out, _i8p))
        return out[:n].copy()

    def _rx_accept(self, name, h, iq, block):
        iq = np.ascontiguousarray(iq, dtype=np.int8)
        ratio = 512 if "2048k" in name else 64
        pcm = np.zeros(iq.size // ratio + 1024, dtype=np.int16)
        total = 0
        fn = getattr(self.lib, self.prefix + name)
        for off in range(0, iq.size, block):
            chunk = iq[off:off + block]
            total += fn(h, _ptr(chunk, _i8p), chunk.size, _ptr(pcm[total:], _i16p))
        return pcm[:total].copy()

    def rx_accept_2048k(self, h, iq, block=262144):
        """IqDataProcessor::acceptIqData in reference-sized blocks."""
        return self._rx_accept("rx_accept_2048k", h, iq, block)

    def rx_accept_256k(self, h, iq, block=32768):
        """<X>Demodulator::acceptIqData in reference-sized blocks."""
        return self._rx_accept("rx_accept_256k", h, iq, block)

    # ---- tx ----
    def tx_new(self):
        return C.c_void_p(getattr(self.lib, self.prefix + "tx_new")())

    def tx_free(self, h):
        getattr(self.lib, self.prefix + "tx_free")(h)

    def tx_set_am_index(self, h, v):
        getattr(self.lib, self.prefix + "tx_set_am_index")(h, C.c_float(v))

    def tx_set_fm_deviation(self, h, v):
        getattr(self.lib, self.prefix + "tx_set_fm_deviation")(h, C.c_float(v))

    def tx_set_wbfm_deviation(self, h, v):
        getattr(self.lib, self.prefix + "tx_set_wbfm_deviation")(h, C.c_float(v))

    def tx_reset_mod(self, h, mod):
        getattr(self.lib, self.prefix + "tx_reset_mod")(h, int(mod))

    def tx_accept(self, h, mode, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.zeros(pcm.size * 512, dtype=np.int8)
        n = getattr(self.lib, self.prefix + "tx_accept")(h, int(mode), _ptr(pcm, _i16p), pcm.size, _ptr(out, _i8p))
        assert n == out.size
        return out

    # ---- convenience: whole-stream runs from a fresh state ----
    def run_rx(self, mode, iq, entry="2048k", gain=None, block=None) -> np.ndarray:
        h = self.rx_new()
        try:
            self.rx_set_mode(h, mode)
            if gain is not None:
                self.rx_set_gain(h, DEMOD_OF_MODE[mode], gain)
            if entry == "2048k":
                return self.rx_accept_2048k(h, iq, block or 262144)
            return self.rx_accept_256k(h, iq, block or 32768)
        finally:
            self.rx_free(h)

    def run_tx(self, mode, pcm, am_index=None, fm_dev=None, wbfm_dev=None) -> np.ndarray:
        h = self.tx_new()
        try:
            if am_index is not None:
                self.tx_set_am_index(h, am_index)
            if fm_dev is not None:
                self.tx_set_fm_deviation(h, fm_dev)
            if wbfm_dev is not None:
                self.tx_set_wbfm_deviation(h, wbfm_dev)
            return self.tx_accept(h, mode, pcm)
        finally:
            self.tx_free(h)


def _declare(lib, prefix):
    vp = C.c_void_p
    sigs = {
        "rx_new": (vp, []),
        "rx_free": (None, [vp]),
        "rx_set_mode": (None, [vp, C.c_int]),
        "rx_set_gain": (None, [vp, C.c_int, C.c_float]),
        "rx_reset_demod": (None, [vp, C.c_int]),
        "rx_front_end": (C.c_size_t, [vp, _i8p, C.c_size_t, _i8p]),
        "rx_accept_2048k": (C.c_size_t, [vp, _i8p, C.c_size_t, _i16p]),
        "rx_accept_256k": (C.c_size_t, [vp, _i8p, C.c_size_t, _i16p]),
        "rx_set_squelch_threshold": (None, [vp, C.c_int32]),
        "tx_new": (vp, []),
        "tx_free": (None, [vp]),
        "tx_set_am_index": (None, [vp, C.c_float]),
        "tx_set_fm_deviation": (None, [vp, C.c_float]),
        "tx_set_wbfm_deviation": (None, [vp, C.c_float]),
        "tx_reset_mod": (None, [vp, C.c_int]),
        "tx_accept": (C.c_size_t, [vp, C.c_int, _i16p, C.c_size_t, _i8p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args


class Oracle(_Base):
    prefix = "hro_"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_checkers()
        self.lib = C.CDLL(ORACLE_SO)
        _declare(self.lib, self.prefix)
        self.lib.hro_taps.restype = C.c_int
        self.lib.hro_taps.argtypes = [C.c_int, _i16p, C.c_int]
        self.lib.hro_atan2_table.argtypes = [_f32p]
        self.lib.hro_nco_tables.argtypes = [_f32p, _f32p]
        self.lib.hro_tx_signals.restype = C.c_size_t
        self.lib.hro_tx_signals.argtypes = [C.c_void_p, C.c_int, _i16p, C.c_size_t, _i8p]
        self.lib.hro_rx_set_rx_gain_db.argtypes = [C.c_void_p, C.c_uint32]
        self.lib.hro_rx_signal_magnitude.restype = C.c_uint32
        self.lib.hro_rx_signal_magnitude.argtypes = [C.c_void_p]
        self.lib.hro_rx_signal_allowed.restype = C.c_int
        self.lib.hro_rx_signal_allowed.argtypes = [C.c_void_p]

    def run_tx_signals(self, head, samples, chunks=None) -> np.ndarray:
        """signals/ tool chain (hro_tx_signals): head 0 = int16 I,Q pairs, 1 dsb, 2 am, 3 pm."""
        samples = np.ascontiguousarray(samples, dtype=np.int16)
        n = samples.size // 2 if head == 0 else samples.size
        out = np.zeros(n * 512, dtype=np.int8)
        h = self.tx_new()
        try:
            at = 0
            for c in (chunks or [n]):
                part = samples[2 * at:2 * (at + c)] if head == 0 else samples[at:at + c]
                self.lib.hro_tx_signals(h, head, _ptr(part, _i16p), c, _ptr(out[at * 512:], _i8p))
                at += c
            assert at == n
        finally:
            self.tx_free(h)
        return out

    def run_rx_squelch(self, mode, iq, threshold, gain_db=16, block=262144, demod_gain=None):
        """IqDataProcessor::acceptIqData block by block with a squelch threshold: returns
        (pcm, per-block average magnitude, per-block squelch decision)."""
        iq = np.ascontiguousarray(iq, dtype=np.int8)
        h = self.rx_new()
        try:
            self.rx_set_mode(h, mode)
            if demod_gain is not None:
                self.rx_set_gain(h, {1: 0, 2: 1, 3: 2, 4: 3, 5: 3}[mode], demod_gain)
            self.lib.hro_rx_set_squelch_threshold(h, int(threshold))
            self.lib.hro_rx_set_rx_gain_db(h, int(gain_db))
            pcm = np.zeros(iq.size // 512 + 1024, dtype=np.int16)
            mags, opens, total = [], [], 0
            for off in range(0, iq.size, block):
                chunk = iq[off:off + block]
                total += self.lib.hro_rx_accept_2048k(h, _ptr(chunk, _i8p), chunk.size, _ptr(pcm[total:], _i16p))
                mags.append(self.lib.hro_rx_signal_magnitude(h))
                opens.append(self.lib.hro_rx_signal_allowed(h))
            return pcm[:total].copy(), np.array(mags, dtype=np.uint32), np.array(opens, dtype=np.uint8)
        finally:
            self.rx_free(h)

    def taps(self, which: int) -> np.ndarray:
        out = np.zeros(64, dtype=np.int16)
        n = self.lib.hro_taps(which, _ptr(out, _i16p), 64)
        return out[:n].copy()

    def atan2_table(self) -> np.ndarray:
        out = np.zeros((256, 256), dtype=np.float32)
        self.lib.hro_atan2_table(_ptr(out, _f32p))
        return out

    def nco_tables(self):
        s = np.zeros(16384, dtype=np.float32)
        c = np.zeros(16384, dtype=np.float32)
        self.lib.hro_nco_tables(_ptr(s, _f32p), _ptr(c, _f32p))
        return s, c


class Ref(_Base):
    prefix = "ref_"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        _declare(self.lib, self.prefix)
        vp = C.c_void_p
        self.lib.ref_taps.restype = C.c_int
        self.lib.ref_taps.argtypes = [vp, vp, C.c_int, _i16p, C.c_int]
        self.lib.ref_nco_tables.argtypes = [vp, _f32p, _f32p]
        self.lib.ref_set_rx_gain_db.argtypes = [C.c_uint32]
        self.lib.ref_rx_run_2048k_squelch.restype = C.c_size_t
        self.lib.ref_rx_run_2048k_squelch.argtypes = [vp, _i8p, C.c_size_t, C.c_size_t, _i16p,
                                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
        self.lib.ref_bench_rx.restype = C.c_double
        self.lib.ref_bench_rx.argtypes = [C.c_int, _i8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                          _i16p, C.c_size_t]
        self.lib.ref_bench_tx.restype = C.c_double
        self.lib.ref_bench_tx.argtypes = [C.c_int, _i16p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                          _i8p, C.c_size_t]

    def run_rx_squelch(self, mode, iq, threshold, gain_db=16, block=262144, demod_gain=None):
        """The unmodified IqDataProcessor with its own notification callbacks reporting per block."""
        iq = np.ascontiguousarray(iq, dtype=np.int8)
        h = self.rx_new()
        try:
            self.rx_set_mode(h, mode)
            if demod_gain is not None:
                self.rx_set_gain(h, {1: 0, 2: 1, 3: 2, 4: 3, 5: 3}[mode], demod_gain)
            self.lib.ref_rx_set_squelch_threshold(h, int(threshold))
            self.lib.ref_set_rx_gain_db(int(gain_db))
            n_blocks = (iq.size + block - 1) // block
            pcm = np.zeros(iq.size // 512 + 1024, dtype=np.int16)
            mags = np.zeros(n_blocks, dtype=np.uint32)
            opens = np.zeros(n_blocks, dtype=np.uint8)
            total = self.lib.ref_rx_run_2048k_squelch(h, _ptr(iq, _i8p), iq.size, block, _ptr(pcm, _i16p),
                                                      mags.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                      opens.ctypes.data_as(C.POINTER(C.c_uint8)))
            return pcm[:total].copy(), mags, opens
        finally:
            self.lib.ref_set_rx_gain_db(16)
            self.rx_free(h)

    def taps(self, which: int) -> np.ndarray:
        rx, tx = self.rx_new(), self.tx_new()
        out = np.zeros(64, dtype=np.int16)
        n = self.lib.ref_taps(rx, tx, which, _ptr(out, _i16p), 64)
        self.rx_free(rx)
        self.tx_free(tx)
        return out[:n].copy()

    def nco_tables(self):
        tx = self.tx_new()
        s = np.zeros(16384, dtype=np.float32)
        c = np.zeros(16384, dtype=np.float32)
        self.lib.ref_nco_tables(tx, _ptr(s, _f32p), _ptr(c, _f32p))
        self.tx_free(tx)
        return s, c

    def bench_rx(self, mode, iq2d: np.ndarray, n_threads: int, want_pcm=False):
        """Time the reference Rx chain over iq2d[n_streams, bytes]; returns (seconds, pcm|None)."""
        iq2d = np.ascontiguousarray(iq2d, dtype=np.int8)
        n, nb = iq2d.shape
        pcm = np.zeros((n, nb // 512), dtype=np.int16) if want_pcm else None
        dt = self.lib.ref_bench_rx(int(mode), _ptr(iq2d, _i8p), nb, nb, n, int(n_threads),
                                   _ptr(pcm, _i16p) if want_pcm else None, nb // 512)
        return dt, pcm

    def bench_tx(self, mode, pcm2d: np.ndarray, n_threads: int, want_iq=False):
        pcm2d = np.ascontiguousarray(pcm2d, dtype=np.int16)
        n, ns = pcm2d.shape
        iq = np.zeros((n, ns * 512), dtype=np.int8) if want_iq else None
        dt = self.lib.ref_bench_tx(int(mode), _ptr(pcm2d, _i16p), ns, ns, n, int(n_threads),
                                   _ptr(iq, _i8p) if want_iq else None, ns * 512)
        return dt, iq


def have_ref() -> bool:
    return os.path.exists(REF_SO)
