"""Multi-GPU host logic on CPU: stream sharding and the max-over-ranks reduction, world_size 2 over gloo.

The data path has no collective (streams are independent), so what N > 1 adds is (1) the partition of
streams over ranks and (2) the timing reduction bench.py prints.  Both are exercised here with two real
processes; the per-rank "work" is the CPU oracle standing in for the per-GPU library (test
infrastructure: it is the checker, the product path is CUDA-only)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from hackrfdiags_b200 import shard, synth  # noqa: E402


def test_shard_range_partitions():
    for n in (0, 1, 7, 1024, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            got = [shard.shard_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_mixed_mode_plan_is_balanced_per_shard():
    mix = {1: 0.25, 2: 0.25, 3: 0.125, 4: 0.1875, 5: 0.1875}
    modes = shard.mixed_mode_plan(4096, mix)
    assert len(modes) == 4096
    for m, f in mix.items():
        assert abs(modes.count(m) - f * 4096) <= 1
    for world in (2, 4, 8):
        for r in range(world):
            mine = shard.shard_modes(modes, world, r)
            groups = dict(shard.mode_groups(mine))
            for m, f in mix.items():
                assert abs(groups[m] - f * 4096 / world) <= 2
            ids = [s for s, _ in mine]
            lo, hi = shard.shard_range(4096, world, r)
            assert sorted(ids) == list(range(lo, hi))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_streams, n_samples, out_dir):
    import torch
    import torch.distributed as dist
    from cpu_checkers import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        modes = shard.mixed_mode_plan(n_streams, {1: 0.5, 4: 0.25, 5: 0.25})
        mine = shard.shard_modes(modes, world, rank)
        oracle = Oracle()
        pcm = {}
        for sid, mode in mine:  # this rank's disjoint stream set: no data leaves the rank
            pcm[sid] = oracle.run_rx(mode, synth.rx_stream(mode, n_samples, stream=sid, config=40))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{str(k): v for k, v in pcm.items()})
        # the bench's reduction: slowest rank sets the time, units add up
        ms_local = 10.0 + 5.0 * rank
        dist.barrier()
        ms_max = shard.reduce_max(ms_local, dist)
        units = torch.tensor([float(len(mine) * n_samples)], dtype=torch.float64)
        dist.all_reduce(units)
        if rank == 0:
            with open(os.path.join(out_dir, "summary.txt"), "w") as f:
                f.write(f"{ms_max} {units.item()} {shard.job_throughput([units.item()], ms_max)}\n")
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo_match_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp

    n_streams, n_samples, world = 8, 8192, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_streams, n_samples, str(tmp_path)), nprocs=world, join=True)
    modes = shard.mixed_mode_plan(n_streams, {1: 0.5, 4: 0.25, 5: 0.25})
    seen = set()
    for r in range(world):
        data = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = shard.shard_range(n_streams, world, r)
        assert sorted(int(k) for k in data.files) == list(range(lo, hi))
        for k in data.files:
            sid = int(k)
            want = oracle.run_rx(modes[sid], synth.rx_stream(modes[sid], n_samples, stream=sid, config=40))
            assert np.array_equal(data[k], want)
            seen.add(sid)
    assert seen == set(range(n_streams))
    ms_max, units, thr = map(float, open(tmp_path / "summary.txt").read().split())
    assert ms_max == 15.0                      # the slower rank
    assert units == n_streams * n_samples      # whole job
    assert abs(thr - units / 15e-3) < 1e-6
