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


def frames_for_rank(n_frames: int, rank: int, world_size: int, block: int = 16):
    """F(k,t) over several GPUs: frames are independent, so rank r takes the blocks of `block`
    consecutive frames b with b mod world_size == r -- no data-path collective; rho[t][k] of all frames is gathered
    once at the end (2*T*K doubles) before the origin/lag table is formed.  -> list of (first, count)."""
    out = []
    for b, first in enumerate(range(0, n_frames, block)):
        if b % world_size == rank:
            out.append((first, min(block, n_frames - first)))
    return out
