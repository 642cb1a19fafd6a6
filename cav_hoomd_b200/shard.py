"""Particle sharding across GPUs (one process per GPU): host-side partitioning and bootstrap.

The data path is in csrc/shard.cu (one 160-byte record per rank per step crosses NVLink); this
module only decides who owns which particles and exchanges the communicator bootstrap blobs with
whatever host-side transport the launcher has (bench.py: torch.distributed, gloo)."""
from __future__ import annotations

import numpy as np


def shard_bounds(N: int, nranks: int, align: int = 32):
    """Contiguous index blocks [lo, hi) per rank, block starts aligned to `align` particles so that
    every shard's Scalar4 / int3 sub-arrays keep their 32-byte / 4-byte alignment."""
    per = -(-N // nranks)
    per = -(-per // align) * align
    out = []
    for r in range(nranks):
        lo = min(r * per, N)
        hi = min(lo + per, N)
        out.append((lo, hi))
    return out


def shard_system(system, rank: int, nranks: int):
    """-> (sub-system arrays, index offset, (group_first, n_group) of the local thermostatted range).
    The thermostatted group is 'everything that is not type L'; it must be a contiguous local range
    (true when the photon is the globally last or first particle, as the reference script builds it)."""
    from .synth import System, w_to_typeid
    lo, hi = shard_bounds(system.N, nranks)[rank]
    sub = System(system.pos[lo:hi].copy(), system.vel[lo:hi].copy(), system.charge[lo:hi].copy(),
                 system.image[lo:hi].copy(), system.box, system.L_typeid, system.types)
    tid = w_to_typeid(sub.pos[:, 3]) if sub.N else np.zeros(0, np.int32)
    mol = np.nonzero(tid != system.L_typeid)[0]
    if len(mol) and not np.array_equal(mol, np.arange(mol[0], mol[0] + len(mol))):
        raise ValueError("local thermostatted group is not contiguous")
    group = (int(mol[0]) if len(mol) else 0, int(len(mol)))
    return sub, lo, group


def exchange_blobs(blob: bytes, dist) -> list:
    """all-gather one bytes object per rank through torch.distributed (any backend)."""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, blob)
    return out


def bootstrap(handle, dist, mode: str = "nvlink"):
    """Wire a capi.Handle into the sharded communicator.  mode: 'nccl' or 'nvlink'."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if mode == "nccl":
        uid = [handle.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        handle.shard_init_nccl(uid[0], rank, world)
    elif mode == "nvlink":
        handles = exchange_blobs(handle.shard_mailbox_export(), dist)
        handle.shard_mailbox_open(handles, rank, world)
    else:
        raise ValueError(mode)
