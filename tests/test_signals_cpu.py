"""The stand-alone tools of signals/ (SURVEY.md section 8f row 3): the oracle's hro_tx_signals against the
reference's own programs, built as the reference builds them (oracle/Makefile: interpolateSignal with
g++ -g -O3, the heads with plain g++) and piped exactly like signals/generateBaseband.sh does."""
import os
import subprocess

import numpy as np
import pytest

from cpu_checkers import Oracle
from hackrfdiags_b200 import synth

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref")
needs_tools = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "interpolateSignal")),
                                 reason="oracle/_ref tools not built (no /root/reference)")
HEADS = {"dsb": 1, "am": 2, "pm": 3, "fm": 4}


def _pipe(data: np.ndarray, *programs) -> np.ndarray:
    buf = data.tobytes()
    for prog in programs:
        buf = subprocess.run([os.path.join(REF, prog)], input=buf, capture_output=True, check=True).stdout
    return np.frombuffer(buf, dtype=np.int8)


def _pcm(kind, n=1536, stream=0):
    return synth.tx_stream(n, stream=stream, config=8, kind=kind)


@needs_tools
@pytest.mark.parametrize("kind", ["sine", "noise", "square", "silence"])
def test_interpolate_signal_raw_iq(kind):
    rng = np.random.default_rng(5)
    i, q = _pcm(kind, stream=1), _pcm(kind, stream=2)
    if kind == "noise":
        q = rng.integers(-32768, 32768, size=i.size).astype(np.int16)
    pairs = np.empty(2 * i.size, dtype=np.int16)
    pairs[0::2], pairs[1::2] = i, q
    want = _pipe(pairs, "interpolateSignal")
    got = Oracle().run_tx_signals(0, pairs, chunks=[512, 1, 1023])
    assert want.size == i.size * 512
    assert np.array_equal(got, want)


@needs_tools
@pytest.mark.parametrize("head", list(HEADS))
@pytest.mark.parametrize("kind", ["sine", "noise", "square"])
def test_heads_through_interpolate_signal(head, kind):
    pcm = _pcm(kind, stream=3)
    want = _pipe(pcm, f"sig_{head}", "interpolateSignal")
    got = Oracle().run_tx_signals(HEADS[head], pcm, chunks=[256, 1280])
    err = np.abs(got.astype(np.int32) - want.astype(np.int32)).max()
    # dsb and am are exact float arithmetic; pm goes through libm cos/sin (same libm here: also exact)
    assert err == 0, f"{head}/{kind}: max abs err {err}"


@needs_tools
def test_count_raw_fixture_through_the_pm_chain():
    """signals/count.raw is the reference's own 5 s speech fixture; it is read here only when the reference
    tree is present (this test is skipped on the GPU box)."""
    path = "/root/reference/signals/count.raw"
    if not os.path.exists(path):
        pytest.skip("no reference tree")
    pcm = np.fromfile(path, dtype=np.int16)[:8000]
    want = _pipe(pcm, "sig_pm", "interpolateSignal")
    got = Oracle().run_tx_signals(3, pcm)
    assert np.array_equal(got, want)
