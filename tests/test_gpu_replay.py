"""hrd_replay: the reference's raw file formats -- *.iq int8 I,Q at 2.048 MS/s,
*.pcm S16_LE at 8 kS/s, the 256 kS/s int8 dump -- replayed as one batch of ragged streams through the C ABI,
in blocks of one reference call, against the oracle fed the same files."""
import os
import subprocess

import numpy as np
import pytest

from cpu_checkers import Oracle
from hackrfdiags_b200 import capi, synth

pytestmark = pytest.mark.gpu
TOOL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hackrfdiags_b200", "replay", "hrd_replay")
BLOCK = 262144


def _run(*args):
    subprocess.run([TOOL, *map(str, args)], check=True, capture_output=True)


def test_rx_files_of_different_lengths(tmp_path):
    oracle = Oracle()
    sizes = [3 * BLOCK + 1000 * 512 + 77, 2 * BLOCK, BLOCK + BLOCK // 2]  # the 77-byte tail is ignored
    paths = []
    for i, size in enumerate(sizes):
        iq = synth.rx_stream(capi.MODE_FM, (size + 1) // 2, stream=i, config=12)[:size]
        p = tmp_path / f"capture{i}.iq"
        iq.tofile(p)
        paths.append(p)
    out = tmp_path / "out"
    out.mkdir()
    _run("rx", "fm", out, *paths)
    for i, p in enumerate(paths):
        iq = np.fromfile(p, dtype=np.int8)
        iq = iq[: iq.size // 512 * 512]
        got = np.fromfile(out / f"capture{i}.pcm", dtype=np.int16)
        assert np.array_equal(got, oracle.run_rx(capi.MODE_FM, iq)), f"file {i}"


def test_rx_squelched_and_front_end_dump(tmp_path):
    oracle = Oracle()
    iq = synth.rx_bursty_stream(capi.MODE_AM, 6, stream=4)
    p = tmp_path / "bursty.iq"
    iq.tofile(p)
    out = tmp_path / "out"
    out.mkdir()
    _run("rx", "am", "-s", -40, out, p)
    want = oracle.run_rx_squelch(capi.MODE_AM, iq, -40)
    got = np.fromfile(out / "bursty.pcm", dtype=np.int16)
    assert 0 < want[2].sum() < 6 and np.array_equal(got, want[0])
    _run("fe", out, p)
    h = oracle.rx_new()
    assert np.array_equal(np.fromfile(out / "bursty.iq256k", dtype=np.int8), oracle.rx_front_end(h, iq))
    oracle.rx_free(h)


@pytest.mark.parametrize("mode,name", [(capi.MODE_AM, "am"), (capi.MODE_WBFM, "wbfm"), (capi.MODE_USB, "usb")])
def test_tx_files(tmp_path, mode, name):
    """What `am < x.pcm > x.iq` (AmModulator/am.cc:31-68) does, for several files at once."""
    oracle = Oracle()
    paths = []
    for i, n in enumerate((1536, 700, 512)):
        p = tmp_path / f"voice{i}.pcm"
        synth.tx_stream(n, stream=i, config=13).tofile(p)
        paths.append(p)
    out = tmp_path / "out"
    out.mkdir()
    _run("tx", name, out, *paths)
    for i, p in enumerate(paths):
        pcm = np.fromfile(p, dtype=np.int16)
        assert np.array_equal(np.fromfile(out / f"voice{i}.iq", dtype=np.int8), oracle.run_tx(mode, pcm)), f"file {i}"


def test_rx_loop_replay_the_data_providers_way(tmp_path):
    """-l: every file is a ring handed out 262144 bytes at a time modulo its length (DataProvider.cc:163-212,
    230-286): blocks straddle the end of the file, an odd-length file swaps I and Q on every lap."""
    oracle = Oracle()
    sizes = [BLOCK + 1000, 70001, 3 * BLOCK]
    paths = []
    for i, size in enumerate(sizes):
        p = tmp_path / f"loop{i}.iq"
        synth.rx_stream(capi.MODE_AM, (size + 1) // 2, stream=i, config=14)[:size].tofile(p)
        paths.append(p)
    out = tmp_path / "out"
    out.mkdir()
    blocks = 5
    _run("rx", "am", "-l", blocks, out, *paths)
    for i, p in enumerate(paths):
        looped = np.resize(np.fromfile(p, dtype=np.int8), blocks * BLOCK)  # cyclic repetition
        got = np.fromfile(out / f"loop{i}.pcm", dtype=np.int16)
        assert np.array_equal(got, oracle.run_rx(capi.MODE_AM, looped)), f"file {i}"
