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


def _fkt_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from cav_hoomd_b200 import replicas, synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, K, n = 23, 6, 400
        s = synth.make_system(n)
        rng = np.random.default_rng(3)
        frames = s.pos[None, :, :3] + 0.1 * np.cumsum(rng.standard_normal((T, s.N, 3)), axis=0)
        kvec = O.numpy_fibonacci_sphere(K) * 0.8
        mine = replicas.frames_for_rank(T, rank, world, block=4)
        rho = torch.zeros((T, K, 2), dtype=torch.float64)
        for first, count in mine:                       # on a GPU box: one cavb200_rhok launch per block
            for t in range(first, first + count):
                r = O.numpy_density_field(frames[t], kvec)
                rho[t, :, 0] = torch.from_numpy(r.real.copy())
                rho[t, :, 1] = torch.from_numpy(r.imag.copy())
        dist.all_reduce(rho)                            # the one exchange: 2*T*K doubles at the end
        full = np.stack([O.numpy_density_field(frames[t], kvec) for t in range(T)])
        got = rho[..., 0].numpy() + 1j * rho[..., 1].numpy()
        covered = sorted(t for f, c in mine for t in range(f, f + c))
        q.put((rank, bool(np.array_equal(got, full)), covered))
    finally:
        dist.destroy_process_group()


def test_fkt_frames_partition_over_ranks_gloo():
    """SURVEY.md 8e, F(k,t): frames are independent -> blocks of frames round-robin over ranks, no data-path collective,
    one all-reduce of rho at the end.  Two gloo ranks on CPU (the NumPy restatement stands in for cavb200_rhok)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fkt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sorted(res[0][2] + res[1][2]) == list(range(23)) and not set(res[0][2]) & set(res[1][2])


def test_frames_for_rank_covers_every_frame_once():
    from cav_hoomd_b200 import replicas
    for T in (0, 1, 15, 16, 17, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                for first, count in replicas.frames_for_rank(T, r, world):
                    assert 0 < count <= 16
                    seen.extend(range(first, first + count))
            assert sorted(seen) == list(range(T))
            # balanced: contiguous ranges whose sizes differ by at most one frame
            seen, sizes = [], []
            for r in range(world):
                mine = replicas.frames_for_rank(T, r, world, block=16, balanced=True)
                assert all(0 < c <= 16 for _, c in mine)
                frames = [f for first, c in mine for f in range(first, first + c)]
                assert frames == list(range(frames[0], frames[0] + len(frames))) if frames else True
                sizes.append(len(frames))
                seen.extend(frames)
            assert sorted(seen) == list(range(T)) and max(sizes) - min(sizes) <= 1
