"""Deterministic synthetic inputs for the parity tests and the bench.

Rx: int8 interleaved I,Q.  At the 2.048 MS/s entry the carrier sits at -64 kHz
(the +Fs/4 rotation at 256 kS/s brings it to DC, IqDataProcessor.cc:937-946);
at the 256 kS/s entry it sits at DC.  Tx: int16 PCM at 8 kS/s.

Stream ``s`` of config ``c`` uses seed ``0x48524644 + 1000*c + s``.  Signal
classes cycle with the stream index; every batch of >= 8 streams contains each
edge class (full-range noise, constant -128, constant +127, alternating +-127,
all-zero) that exercises the reference's wrap-around casts.
"""
from __future__ import annotations

import numpy as np

SEED0 = 0x48524644
FS_RX = 2_048_000
FS_DEMOD = 256_000
FS_PCM = 8_000

AM, FM, WBFM, LSB, USB = 1, 2, 3, 4, 5

AMPLITUDES = (20.0, 60.0, 100.0, 127.0)
SIGMAS = (1.0, 3.0, 10.0)
EDGE_CLASSES = ("noise", "min", "max", "alt", "zero")


def stream_seed(config: int, stream: int) -> int:
    return SEED0 + 1000 * config + stream


def _baseband(mode: int, t: np.ndarray) -> np.ndarray:
    """Unit-amplitude complex envelope of the wanted signal, carrier at DC."""
    two_pi = 2.0 * np.pi
    if mode == AM:
        return (1.0 + 0.8 * np.sin(two_pi * 1000.0 * t)) / 1.8 + 0j
    if mode == FM:
        # 1 kHz tone, +-3 kHz deviation -> phase = (3000/1000) * -cos(2 pi 1000 t)
        return np.exp(1j * (3000.0 / 1000.0) * (1.0 - np.cos(two_pi * 1000.0 * t)))
    if mode == WBFM:
        ph = (45000.0 / 1000.0) * (1.0 - np.cos(two_pi * 1000.0 * t)) \
            + (30000.0 / 10000.0) * (1.0 - np.cos(two_pi * 10000.0 * t))
        return np.exp(1j * ph)
    if mode in (LSB, USB):
        sign = -1.0 if mode == LSB else 1.0
        return 0.5 * (np.exp(sign * 1j * two_pi * 700.0 * t) + np.exp(sign * 1j * two_pi * 1900.0 * t))
    raise ValueError(mode)


def rx_stream(mode: int, n_samples: int, stream: int = 0, config: int = 0,
              entry: str = "2048k", edge: str | None = None) -> np.ndarray:
    """One stream of ``n_samples`` IQ samples -> int8 array of 2*n_samples bytes."""
    rng = np.random.default_rng(stream_seed(config, stream))
    if edge == "noise":
        return rng.integers(-128, 128, size=2 * n_samples, dtype=np.int64).astype(np.int8)
    if edge == "min":
        return np.full(2 * n_samples, -128, dtype=np.int8)
    if edge == "max":
        return np.full(2 * n_samples, 127, dtype=np.int8)
    if edge == "alt":
        v = np.where(np.arange(n_samples) % 2 == 0, 127, -127).astype(np.int8)
        return np.repeat(v, 2)
    if edge == "zero":
        return np.zeros(2 * n_samples, dtype=np.int8)
    fs = FS_RX if entry == "2048k" else FS_DEMOD
    fc = -64000.0 if entry == "2048k" else 0.0
    amp = AMPLITUDES[stream % len(AMPLITUDES)]
    sigma = SIGMAS[(stream // len(AMPLITUDES)) % len(SIGMAS)]
    t = np.arange(n_samples, dtype=np.float64) / fs
    z = amp * _baseband(mode, t) * np.exp(2j * np.pi * fc * t)
    i = np.rint(z.real + rng.normal(0.0, sigma, n_samples))
    q = np.rint(z.imag + rng.normal(0.0, sigma, n_samples))
    out = np.empty(2 * n_samples, dtype=np.int8)
    out[0::2] = np.clip(i, -128, 127).astype(np.int8)
    out[1::2] = np.clip(q, -128, 127).astype(np.int8)
    return out


def rx_batch(mode: int, n_streams: int, n_samples: int, config: int = 0,
             entry: str = "2048k", with_edges: bool = True) -> np.ndarray:
    """[n_streams, 2*n_samples] int8.  The last min(5, n_streams-1) streams are edge classes."""
    out = np.empty((n_streams, 2 * n_samples), dtype=np.int8)
    n_edge = min(len(EDGE_CLASSES), max(n_streams - 1, 0)) if with_edges else 0
    for s in range(n_streams):
        k = s - (n_streams - n_edge)
        edge = EDGE_CLASSES[k] if k >= 0 else None
        out[s] = rx_stream(mode, n_samples, s, config, entry, edge)
    return out


TX_CLASSES = ("sine", "noise", "silence", "square", "speechlike")


def tx_stream(n_samples: int, stream: int = 0, config: int = 0, kind: str | None = None) -> np.ndarray:
    """One PCM stream (int16 at 8 kS/s).  Full-scale sines reach -32768."""
    rng = np.random.default_rng(stream_seed(config, stream))
    kind = kind or TX_CLASSES[stream % len(TX_CLASSES)]
    t = np.arange(n_samples, dtype=np.float64) / FS_PCM
    if kind == "sine":
        f = 300.0 + (stream * 97) % 3100
        x = np.rint(-32768.0 * np.cos(2.0 * np.pi * f * t))
    elif kind == "noise":
        return rng.integers(-32768, 32768, size=n_samples, dtype=np.int64).astype(np.int16)
    elif kind == "silence":
        return np.zeros(n_samples, dtype=np.int16)
    elif kind == "square":
        x = np.where((np.arange(n_samples) // 7) % 2 == 0, 32767, -32767)
    else:  # band-limited "speech": a few harmonics with a slow envelope
        env = 0.5 * (1.0 + np.sin(2.0 * np.pi * 3.0 * t))
        x = np.zeros(n_samples)
        for h, a in ((1, 1.0), (2, 0.6), (3, 0.4), (5, 0.2)):
            x += a * np.sin(2.0 * np.pi * 140.0 * h * t + rng.uniform(0, 6.28))
        x = np.rint(12000.0 * env * x / 2.2)
    return np.clip(x, -32768, 32767).astype(np.int16)


def tx_batch(n_streams: int, n_samples: int, config: int = 0) -> np.ndarray:
    out = np.empty((n_streams, n_samples), dtype=np.int16)
    for s in range(n_streams):
        out[s] = tx_stream(n_samples, s, config)
    return out


def rx_bursty_stream(mode: int, n_blocks: int, stream: int = 0, config: int = 7,
                     levels=(100, 2, 2, 60, 60, 1, 20, 127, 0, 0, 40, 3)) -> np.ndarray:
    """Squelch test input at the 2.048 MS/s entry: ``n_blocks`` blocks of 131072 IQ samples (one reference
    call each) whose carrier amplitude changes from block to block (cycling through ``levels``, rotated by the
    stream number), so that a squelch threshold opens, holds for its one-block tail and closes again."""
    rng = np.random.default_rng(stream_seed(config, stream))
    n_blk = 131072
    n = n_blocks * n_blk
    t = np.arange(n, dtype=np.float64) / FS_RX
    amp = np.repeat(np.array([levels[(b + stream) % len(levels)] for b in range(n_blocks)], dtype=np.float64), n_blk)
    z = amp * _baseband(mode, t) * np.exp(2j * np.pi * -64000.0 * t)
    out = np.empty(2 * n, dtype=np.int8)
    out[0::2] = np.clip(np.rint(z.real + rng.normal(0.0, 1.0, n)), -128, 127).astype(np.int8)
    out[1::2] = np.clip(np.rint(z.imag + rng.normal(0.0, 1.0, n)), -128, 127).astype(np.int8)
    return out
