"""Replica bookkeeping: --replicas parsing and the replica -> GPU map.

parse_replicas mirrors reference examples/05_advanced_run.py:1336-1351.  The reference then runs
the replicas one after another in a for loop (:1570-1612); here independent replicas are spread
one per GPU, one process per GPU, with no data-path collective (BASELINE.json configs[2])."""
from __future__ import annotations


def parse_replicas(replicas_str):
    """'1-3,7' -> [1, 2, 3, 7]; empty -> [1] (reference :1336-1351)."""
    if not replicas_str:
        return [1]
    replicas = []
    for part in replicas_str.split(","):
        part = part.strip()
        if "-" in part:
            start, end = part.split("-", 1)
            replicas.extend(range(int(start.strip()), int(end.strip()) + 1))
        else:
            replicas.append(int(part))
    return sorted(set(replicas))


def replicas_for_rank(replicas, rank: int, world_size: int):
    """Round-robin: replica r (position p in the sorted list) runs on rank p mod world_size."""
    return [r for p, r in enumerate(sorted(replicas)) if p % world_size == rank]


def gpu_for_replica(replicas, replica: int, n_gpus: int) -> int:
    return sorted(replicas).index(replica) % n_gpus


def frames_for_rank(n_frames: int, rank: int, world_size: int, block: int = 16, balanced: bool = False):
    """F(k,t) over several GPUs: frames are independent, so rank r takes the blocks of `block`
    consecutive frames b with b mod world_size == r -- no data-path collective; rho[t][k] of all frames is gathered
    once at the end (2*T*K doubles) before the origin/lag table is formed.  -> list of (first, count).
    balanced=True: every rank takes one contiguous range of n_frames/world_size frames (the first n_frames mod
    world_size ranks one more), cut into launches of at most `block` -- 1000 frames over 8 ranks are 125 each instead
    of 4 blocks of 32 on seven ranks and 3.25 on the last (the kernel's grid rule makes any launch size efficient)."""
    out = []
    if balanced:
        base, extra = divmod(n_frames, world_size)
        lo = rank * base + min(rank, extra)
        hi = lo + base + (1 if rank < extra else 0)
        for first in range(lo, hi, block):
            out.append((first, min(block, hi - first)))
        return out
    for b, first in enumerate(range(0, n_frames, block)):
        if b % world_size == rank:
            out.append((first, min(block, n_frames - first)))
    return out
