"""Stream sharding across the GPUs of one box.

Streams are independent (every piece of state belongs to one reference object graph), so a job
of N streams on G GPUs is G disjoint jobs: rank g owns the contiguous range
[g*N/G, (g+1)*N/G) and the state records of those streams, inputs land directly in the owning
GPU's HBM, and NOTHING crosses GPUs on the data path -- no collective, no peer traffic.  The only
communication a multi-GPU run needs is control: a barrier around the timed region and the
max-over-ranks of the per-rank device time (``reduce_max``), which work over any
``torch.distributed`` backend (NCCL on the GPU box, gloo in the CPU tests).

Host-side logic only; no CUDA here.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def shard_range(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the streams rank ``rank`` owns; ranges are contiguous, disjoint, cover 0..n-1
    and differ in size by at most one stream."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of world {world}")
    if n_streams < 0:
        raise ValueError("n_streams must be non-negative")
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


def shard_modes(modes: Sequence[int], world: int, rank: int) -> List[Tuple[int, int]]:
    """(global stream id, mode) of every stream this rank owns, in the order the rank's batch holds them.

    Inside a rank the streams are grouped by mode (the library launches one kernel per mode group, so a
    mode-homogeneous layout keeps each launch dense); the grouping is stable, so the k-th stream of a mode
    on this rank is the k-th stream of that mode in the rank's global range."""
    lo, hi = shard_range(len(modes), world, rank)
    mine = [(s, int(modes[s])) for s in range(lo, hi)]
    return sorted(mine, key=lambda sm: sm[1])


def mode_groups(assignment: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """[(mode, count)] of an assignment produced by ``shard_modes`` (runs of equal mode)."""
    out: List[Tuple[int, int]] = []
    for _, m in assignment:
        if out and out[-1][0] == m:
            out[-1] = (m, out[-1][1] + 1)
        else:
            out.append((m, 1))
    return out


def mixed_mode_plan(n_streams: int, mix: Dict[int, float]) -> List[int]:
    """Deterministic mode of every stream of a mixed job: modes are interleaved so that any contiguous
    shard sees (almost) the same mix (config 5 of BASELINE.json: mixed-mode sweep)."""
    if not mix or abs(sum(mix.values()) - 1.0) > 1e-6:
        raise ValueError("mix must be a {mode: fraction} dict summing to 1")
    acc = {m: 0.0 for m in mix}
    out = []
    for _ in range(n_streams):
        for m in acc:
            acc[m] += mix[m]
        best = max(sorted(acc), key=lambda m: acc[m])
        acc[best] -= 1.0
        out.append(best)
    return out


def reduce_max(value: float, dist=None, device=None) -> float:
    """Max over ranks of a per-rank scalar (the bench's ms_per_step); identity without a process group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: Sequence[float], ms_max: float) -> float:
    """Whole-job units per second: everything all ranks processed divided by the slowest rank's time."""
    return sum(units_per_rank) / (ms_max * 1e-3)
