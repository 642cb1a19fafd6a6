"""Host-side logic (no GPU): synthetic inputs, replica parsing, shard partitioning, RNG keying,
the API surface of the host mirrors."""
import os

import numpy as np
import pytest

from cav_hoomd_b200 import replicas, rng, shard, synth
from oracle import oracle as O


def test_type_id_packing_round_trip():
    tid = np.array([0, 1, 2, 7, 2**31 - 1], dtype=np.int32)
    w = synth.typeid_to_w(tid)
    assert np.array_equal(synth.w_to_typeid(w), tid)
    assert np.all(w.view(np.uint64) >> 32 == 0)  # only the low 32 bits carry the id


def test_synthetic_system_recipe():
    s = synth.make_system(1000)
    assert s.N == 1001 and s.typeid[-1] == synth.L_TYPEID and s.charge[-1] == 0.0
    assert abs(s.charge[:-1].sum()) == 0.0  # neutral
    assert np.array_equal(s.pos[:1000, :3], s.pos[:1000, :3].astype(np.float32))  # float32-exact (GSD)
    assert abs(s.box[0] - (1000 / 5.5e-3) ** (1 / 3)) < 1e-9
    assert set(np.unique(s.image)) <= {-1, 0, 1}
    again = synth.make_system(1000)
    assert np.array_equal(s.pos, again.pos) and np.array_equal(s.vel, again.vel)
    for photon, n in (("first", 1001), ("middle", 1001), ("absent", 1000), ("duplicated", 1002)):
        assert synth.make_system(1000, photon=photon).N == n
    g = synth.molecular_group(synth.make_system(10, photon="middle"))
    assert len(g) == 10 and 5 not in g


def test_fibonacci_sphere_matches_reference_loop():
    for K in (2, 50, 64, 100):
        assert np.allclose(synth.fibonacci_sphere(K), O.numpy_fibonacci_sphere(K), rtol=0, atol=2e-16)
    v = synth.fibonacci_sphere(64)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-15)


def test_parse_replicas_like_the_reference():
    """reference examples/05_advanced_run.py:1336-1351"""
    assert replicas.parse_replicas("") == [1]
    assert replicas.parse_replicas(None) == [1]
    assert replicas.parse_replicas("1-8") == list(range(1, 9))
    assert replicas.parse_replicas("3, 1-2,2 , 7") == [1, 2, 3, 7]
    with pytest.raises(ValueError):
        replicas.parse_replicas("a-b")
    r = replicas.parse_replicas("1-8")
    assert [replicas.replicas_for_rank(r, k, 4) for k in range(4)] == [[1, 5], [2, 6], [3, 7], [4, 8]]
    assert replicas.gpu_for_replica(r, 6, 8) == 5


def test_shard_bounds_cover_and_align():
    for N, R in ((1000001, 8), (16000001, 8), (100, 8), (5, 2), (0, 4)):
        b = shard.shard_bounds(N, R)
        assert len(b) == R and b[0][0] == 0 and b[-1][1] == N
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2 and (lo % 32 == 0 or lo == hi)  # empty trailing shards start at N
    s = synth.make_system(1000)
    sub, off, (first, n) = shard.shard_system(s, 1, 2)
    assert off == 512 and sub.N == 1001 - 512 and first == 0 and n == sub.N - 1


def test_bussi_draws_are_counter_based():
    a = rng.bussi_draws(10, 42, 0, 2997.0)
    assert a == rng.bussi_draws(10, 42, 0, 2997.0)
    assert a != rng.bussi_draws(11, 42, 0, 2997.0) and a != rng.bussi_draws(10, 43, 0, 2997.0)
    assert rng.bussi_draws(10, 42, 0, 1.0)[1] == 0.0  # dof <= 1: no gamma draw
    g = np.array([rng.bussi_draws(t, 1, 0, 101.0)[1] for t in range(2000)])
    assert abs(g.mean() - 50.0) < 1.0  # Gamma((dof-1)/2, 1) has mean (dof-1)/2


def test_host_mirror_api_surface():
    """Same names and constructor signatures as the reference wrappers."""
    import inspect
    from cav_hoomd_b200 import BussiReservoir, CavityForce
    assert list(inspect.signature(CavityForce.__init__).parameters)[1:] == ["kvector", "couplstr", "omegac", "phmass",
                                                                           "force_python"]
    assert list(inspect.signature(BussiReservoir.__init__).parameters)[1:] == ["kT", "tau"]
    for name in ("harmonic_energy", "coupling_energy", "dipole_self_energy", "total_cavity_energy", "energy",
                 "implementation", "forces"):
        assert hasattr(CavityForce, name)
    for name in ("reservoir_energy_translational", "reservoir_energy_rotational", "total_reservoir_energy",
                 "instantaneous_reservoir_translational", "instantaneous_reservoir_rotational",
                 "instantaneous_reservoir_total", "reset_reservoir_energy"):
        assert hasattr(BussiReservoir, name)
    with pytest.raises(NotImplementedError):
        CavityForce([0, 0, 1], 1e-3, 0.01, force_python=True)  # no fallback in this build
    t = BussiReservoir(kT=1.5, tau=0.1)
    assert t.kT == 1.5 and t.tau == 0.1 and t.total_reservoir_energy == 0.0  # reference test :59-61
    t.reset_reservoir_energy()  # no-op when not attached (reference thermostats.py:137-158)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_bench_sharded_blocks_tile_the_16m_box(world):
    """bench.py's `sharded` leg (BASELINE configs[3]) gives rank r the block shard_bounds(16 000 001, world)[r]; the
    blocks must tile the index space in rank order with 32-particle-aligned starts, and the photon (globally last) must
    fall into the last non-empty block, whose thermostatted window then excludes exactly that particle."""
    from cav_hoomd_b200 import shard
    N = 16_000_001
    b = shard.shard_bounds(N, world)
    assert b[0][0] == 0 and b[-1][1] == N and all(lo % 32 == 0 for lo, _ in b)
    assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
    sizes = [hi - lo for lo, hi in b]
    assert max(sizes) - min(sizes) <= 32 * world + 1 and min(sizes) > 0
    last = world - 1
    assert b[last][0] <= N - 1 < b[last][1]
    # what the leg computes per rank: molecular particles of the block, the window [0, n_loc_mol)
    n_loc_mol = [(hi - lo) - (1 if r == last else 0) for r, (lo, hi) in enumerate(b)]
    assert sum(n_loc_mol) == N - 1


def test_bench_config_is_identical_on_both_arms():
    """The driver compares the `config` dictionaries of the two arms: both come from bench.bench_config."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c1, c8 = bench.bench_config(1_000_000, 1), bench.bench_config(1_000_000, 8)
    assert set(c1) == {"workload", "l2"} and "BASELINE configs[1]" in c1["workload"]
    assert "8 independent replicas" in c8["workload"] and c1["l2"] == c8["l2"]
    assert bench.bench_config(1_000_000, 1) == c1
