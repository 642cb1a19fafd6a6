"""N > 1 host path on CPU: two gloo ranks partition a system with shard_system, exchange bootstrap
blobs and per-rank partial records through torch.distributed, combine them in RANK ORDER, and must
reproduce the unsharded oracle.  (The device data path of the same exchange is csrc/shard.cu;
tools/shard_check.py runs it under torchrun on >= 2 GPUs.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_mol, photon, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cav_hoomd_b200 import shard, synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = synth.make_system(n_mol, photon=photon)
        sub, off, (first, n) = shard.shard_system(s, rank, world)
        # bootstrap exchange: every rank must see every rank's blob in rank order
        blobs = shard.exchange_blobs(bytes([rank]) * 64, dist)
        assert [b[0] for b in blobs] == list(range(world))
        # per-rank partial record on the CPU: local dipole (exact), first local 'L', local m|v|^2 sum
        co = O.COracle()
        tid = synth.w_to_typeid(sub.pos[:, 3])
        isL = np.nonzero(tid == s.L_typeid)[0]
        first_L = int(off + isL[0]) if len(isL) else -1
        u = sub.pos[:, :3] + sub.image * np.asarray(s.box)
        skip = int(isL[0]) if len(isL) else -1
        d_local = co.dipole_exact(sub.pos, sub.charge, sub.image, s.box, skip) if sub.N else np.zeros(3)
        rec = dict(d=d_local, first_L=first_L, q=u[isL[0]] if len(isL) else np.zeros(3),
                   t=sub.charge[isL[0]] * u[isL[0]] if len(isL) else np.zeros(3),
                   ke=float(np.sum(sub.vel[first:first + n, 3] * np.sum(sub.vel[first:first + n, :3] ** 2, axis=1))))
        recs = [None] * world
        dist.all_gather_object(recs, rec)
        # combine in rank order: the global first 'L' wins, the others' terms go back into d
        cands = [r["first_L"] for r in recs if r["first_L"] >= 0]
        gmin = min(cands) if cands else -1
        d = np.zeros(3)
        ke = 0.0
        qph = np.zeros(3)
        for r in recs:
            d = d + r["d"]
            ke += r["ke"]
            if r["first_L"] >= 0:
                if r["first_L"] == gmin:
                    qph = r["q"]
                else:
                    d = d + r["t"]
        q.put((rank, d, ke, gmin, qph))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("photon", ["last", "first"])
def test_two_rank_sharded_combine_matches_unsharded_oracle(coracle, photon):
    import multiprocessing as mp
    from cav_hoomd_b200 import synth
    n_mol, world = 3000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_mol, photon, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = synth.make_system(n_mol, photon=photon)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    exact = coracle.dipole_exact(s.pos, s.charge, s.image, s.box, ref["photon_idx"])
    mol = synth.molecular_group(s)
    ke_ref = coracle.kinetic_energy(s.vel, mol) * 2.0
    results.sort()
    for rank, d, ke, gmin, qph in results:
        assert gmin == ref["photon_idx"]
        assert np.allclose(d, exact, rtol=1e-13)
        assert abs(ke - ke_ref) <= 1e-12 * ke_ref
        u = s.pos[gmin, :3] + s.image[gmin] * np.asarray(s.box)
        assert np.array_equal(qph, u)
    # every rank computed the same thing
    assert np.array_equal(results[0][1], results[1][1])


def test_shard_system_refuses_a_split_thermostat_group():
    """A photon in the middle of a shard would need an index list; shard_system says so loudly."""
    from cav_hoomd_b200 import shard, synth
    s = synth.make_system(1000, photon="middle")
    with pytest.raises(ValueError):
        shard.shard_system(s, 0, 2)
