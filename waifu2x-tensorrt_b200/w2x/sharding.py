"""Frame sharding across the GPUs of one box: frames / images / tiles are independent, so rank r of N
renders frames r, r+N, r+2N, ... with its own engine and weight replica; there is NO data-path collective.  The only
cross-rank traffic is the ordering metadata and the timing reduction (max over ranks) done with torch.distributed."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: frame f -> rank f mod world ('frame-parallel')."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def owner_of(frame: int, world: int) -> int:
    return frame % world


def merge_in_order(per_rank: Sequence[Sequence[Tuple[int, object]]]) -> List[object]:
    """Re-order queue: results arrive as (frame_index, payload) per rank; the writer needs frame order."""
    flat = [x for r in per_rank for x in r]
    flat.sort(key=lambda t: t[0])
    idx = [t[0] for t in flat]
    if idx != list(range(len(idx))):
        raise ValueError("frames missing or duplicated across ranks")
    return [t[1] for t in flat]


def band_rows(n_tile_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Single-large-image mode: contiguous bands of the GLOBAL tile grid's rows (never re-tile: SE pooling is per tile,
    SURVEY 8e note).  Returns [first, last) tile-row indices for this rank; bands differ by at most one row."""
    base, extra = divmod(n_tile_rows, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def max_over_ranks_ms(local_ms: float) -> float:
    """Timing contract of bench.py: the job's time is the slowest rank's device time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_ms
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
