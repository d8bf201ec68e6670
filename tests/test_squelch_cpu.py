"""Squelch gate + signal magnitude: the oracle's restatement of
Squelch::run / SignalDetector::detectSignal / SignalTracker::run / DbfsCalculator against the
UNMODIFIED reference classes driven through IqDataProcessor::acceptIqData, block by block, with the
reference's own notification callbacks reporting the magnitude and the decision of every block."""
import numpy as np
import pytest

from cpu_checkers import Oracle, Ref, have_ref
from hackrfdiags_b200 import synth

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("threshold,gain_db", [(-200, 16), (-40, 16), (-30, 0), (-25, 16), (-50, 40), (0, 16)])
def test_oracle_squelch_matches_reference(mode, threshold, gain_db):
    iq = synth.rx_bursty_stream(mode, 12, stream=mode)
    want = Ref().run_rx_squelch(mode, iq, threshold, gain_db)
    got = Oracle().run_rx_squelch(mode, iq, threshold, gain_db)
    assert np.array_equal(got[1], want[1]), "per-block magnitudes differ"
    assert np.array_equal(got[2], want[2]), "per-block squelch decisions differ"
    assert np.array_equal(got[0], want[0]), "PCM differs"
    if threshold == -200:
        assert got[2].all() and got[0].size == 12 * 512


@needs_ref
def test_squelch_edge_inputs_and_short_blocks():
    ref, ora = Ref(), Oracle()
    for edge in ("noise", "min", "max", "alt", "zero"):
        iq = synth.rx_stream(2, 3 * 131072 + 4096, stream=3, edge=edge)
        for thr in (-60, -20):
            want = ref.run_rx_squelch(2, iq, thr)
            got = ora.run_rx_squelch(2, iq, thr)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (edge, thr)


def test_squelch_gate_actually_gates():
    """The bursty input crosses a -40 dBFS threshold in both directions: some blocks are dropped, and a block
    after a loud one still passes (the one-block tail of SignalTracker.cc:127-138)."""
    iq = synth.rx_bursty_stream(1, 12, stream=0)
    pcm, mags, opens = Oracle().run_rx_squelch(1, iq, -40)
    assert 0 < opens.sum() < 12
    assert pcm.size == 512 * int(opens.sum())
    loud = mags >= mags.max() // 2
    assert any(opens[b] and not loud[b] and loud[b - 1] for b in range(1, 12)), "no squelch tail seen"
