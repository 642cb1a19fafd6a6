"""Thermostatted harness step: cavb200_nvt_step_one ; cavb200_force ; cavb200_nvt_step_two
against oracle/cavity_oracle.c orc_nvt_step on the same input and the same injected random draws.

The harness is the repo's own (HOOMD's TwoStepConstantVolume is upstream code that is not in the
reference tree); what is the reference's on this path is compute_rescale_factor and the reservoir
bookkeeping (reference src/BussiReservoirThermostat.h:86-95,177-225) and the cavity force
(src/CavityForceCompute.cc:134-208).  Tolerances: alpha <= 1e-12 relative; KE and reservoir <= 1e-12 on the
first step and <= 1e-10 along the trajectory; positions, velocities, forces <= 1e-10 relative to the
array's largest magnitude (BASELINE tolerance).  Only the ORDER of the dipole sum differs between the
arms; in the synthetic box the cavity force dominates the motion, so that last-bit
difference in d feeds back through the trajectory (4e-12 in KE after 20 steps at N = 20000)."""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, rng, synth

pytestmark = pytest.mark.gpu

OMEGAC, G, PHMASS = 0.01, 1e-3, 1.0


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _run_both(handle, coracle, s, steps, dt, first, n, dof, kT, tau, thermostat=True):
    st = capi.Stream()
    p = capi.Params.make(OMEGAC, G, PHMASS)
    # --- oracle arm ---
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    force[:] = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G, PHMASS)["force"]
    idx = np.arange(first, first + n, dtype=np.uint32)
    ke = np.array([coracle.kinetic_energy(vel, idx)])
    res = np.zeros(2)
    # --- GPU arm ---
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    handle.bussi_reset(st.ptr)
    handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
    handle.bussi_ke(d["vel"], None, first, n, st.ptr)
    assert abs(handle.bussi_read(st.ptr)["ke"] - ke[0]) <= 1e-12 * abs(ke[0]) + 1e-300
    out = []
    for t in range(steps):
        r, gm = rng.bussi_draws(t, 7, first, dof) if thermostat else (0.0, 0.0)
        a_ref, en_ref = coracle.nvt_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, OMEGAC, G, PHMASS, dt,
                                         first, n, dof if thermostat else 0.0, kT, tau, r, gm, res, ke)
        a = capi.BussiArgs(kT, tau, dt, dof, r, gm) if thermostat else None
        handle.nvt_step_one(d["pos"], d["vel"], d_f, s.N, dt, first, n, a, st.ptr)
        handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
        handle.nvt_step_two(d["vel"], d_f, s.N, dt, first, n, st.ptr)
        b = handle.bussi_read(st.ptr)
        en = handle.force_read(st.ptr)[0]
        out.append((a_ref, b["alpha"], ke[0], b["ke"], res.copy(), (b["cumulative"], b["instantaneous"]), en_ref, en))
    return out, (pos, vel, force), (d["pos"].numpy(st.ptr), d["vel"].numpy(st.ptr), d_f.numpy(st.ptr))


@pytest.mark.parametrize("n_mol", [1, 33, 257, 1000, 20000])
def test_nvt_step_matches_oracle(handle, coracle, n_mol):
    s = synth.make_system(n_mol, replica=3)
    dof = max(3.0 * n_mol - 3.0, 1.0)
    out, cpu, gpu = _run_both(handle, coracle, s, 20, synth.DT_1FS, 0, n_mol, dof, synth.KT_100K, synth.TAU_5PS)
    for t, (a_ref, a, ke_ref, ke, res_ref, res, en_ref, en) in enumerate(out):
        tol = 1e-12 if t == 0 else 1e-10
        assert abs(a - a_ref) <= 1e-12 * abs(a_ref)
        assert abs(ke - ke_ref) <= tol * abs(ke_ref)
        assert abs(res[0] - res_ref[0]) <= tol * abs(ke_ref) and abs(res[1] - res_ref[1]) <= tol * abs(ke_ref)
        assert np.allclose(en, en_ref, rtol=1e-10, atol=1e-300)
    for c, g_ in zip(cpu, gpu):
        assert _rel(g_[:, :3], c[:, :3]) <= 1e-10
    # mass / type words untouched
    assert np.array_equal(gpu[0][:, 3].view(np.uint64), s.pos[:, 3].view(np.uint64))
    assert np.array_equal(gpu[1][:, 3], s.vel[:, 3])
    assert handle.bussi_read()["err"] == 0.0


def test_nvt_windowed_group_and_photon_outside(handle, coracle):
    """The thermostatted group is a window [first, first+n); the photon (last) and the particles outside
    the window are integrated but never rescaled, and do not count in KE."""
    n_mol = 3000
    s = synth.make_system(n_mol, replica=5)
    first, n = 500, 2001
    out, cpu, gpu = _run_both(handle, coracle, s, 10, synth.DT_1FS, first, n, 3.0 * n - 3.0, synth.KT_100K,
                              synth.TAU_5PS)
    for a_ref, a, ke_ref, ke, *_ in out:
        assert abs(a - a_ref) <= 1e-12 * abs(a_ref) and abs(ke - ke_ref) <= 1e-10 * abs(ke_ref)
    for c, g_ in zip(cpu, gpu):
        assert _rel(g_[:, :3], c[:, :3]) <= 1e-10


def test_nvt_without_thermostat_is_the_nve_harness(handle, coracle):
    """bussi == NULL: alpha = 1, identical (bit for bit) to nve_kick_drift ; force ; nve_half_kick."""
    n_mol = 5000
    s = synth.make_system(n_mol, replica=2)
    out, cpu, gpu = _run_both(handle, coracle, s, 5, 5.0, 0, n_mol, 0.0, synth.KT_100K, synth.TAU_5PS,
                              thermostat=False)
    st = capi.Stream()
    p = capi.Params.make(OMEGAC, G, PHMASS)
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
    for _ in range(5):
        handle.nve_kick_drift(d["pos"], d["vel"], d_f, s.N, 5.0, st.ptr)
        handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
        handle.nve_half_kick(d["vel"], d_f, s.N, 5.0, st.ptr)
    assert np.array_equal(d["pos"].numpy(st.ptr), gpu[0])
    assert np.array_equal(d["vel"].numpy(st.ptr), gpu[1])
    for c, g_ in zip(cpu, gpu):
        assert _rel(g_[:, :3], c[:, :3]) <= 1e-10


def test_nvt_zero_kinetic_energy_sets_error_flag(handle):
    """dof != 0 with KE == 0: the reference throws (src/BussiReservoirThermostat.h:57-61); here the
    device error flag is raised, no rescale is applied and the step still integrates."""
    s = synth.make_system(100)
    s.vel[:, :3] = 0.0
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "vel")}
    d_f = capi.DeviceArray.from_numpy(np.zeros((s.N, 4)))
    handle.bussi_reset()
    handle.bussi_ke(d["vel"], None, 0, 100)
    handle.nvt_step_one(d["pos"], d["vel"], d_f, s.N, 1.0, 0, 100, capi.BussiArgs(1e-3, 10.0, 1.0, 297.0, 0.1, 140.0))
    assert handle.bussi_read()["err"] == 1.0
    assert np.array_equal(d["vel"].numpy(), s.vel) and np.array_equal(d["pos"].numpy(), s.pos)
    handle.bussi_reset()
    assert handle.bussi_read()["err"] == 0.0


def test_nvt_argument_errors(handle):
    a = capi.DeviceArray((8, 4), np.float64)
    with pytest.raises(capi.CavbError):
        handle.nvt_step_one(a, a, a, 8, 1.0, 4, 5)  # window past the end
    with pytest.raises(capi.CavbError):
        handle.nvt_step_two(a, None, 8, 1.0, 0, 8)
    handle.nvt_step_one(None, None, None, 0, 1.0, 0, 0)  # N == 0: success, no-op
    handle.nvt_step_two(None, None, 0, 1.0, 0, 0)


# ---- rank-1 cavity force --------------------------------------------------------
@pytest.mark.parametrize("photon", ["last", "first", "middle", "absent", "duplicated"])
@pytest.mark.parametrize("n_mol", [1, 33, 1000, 20000])
def test_force_rank1_scalars_and_net_force_add(handle, coracle, n_mol, photon):
    """cavb200_force_rank1 leaves the same energies / dipole / photon index as cavb200_force without writing a
    force array; net_force += F_i formed from the charge equals net + the oracle's stored force exactly
    (one rounding: the add), for every photon placement the reference handles."""
    s = synth.make_system(n_mol, replica=n_mol + 1, photon=photon)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G, PHMASS)
    p = capi.Params.make(OMEGAC, G, PHMASS)
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p)
    en0, dip0, ph0 = handle.force_read()
    f_stored = d_f.numpy()
    handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p)
    en1, dip1, ph1 = handle.force_read()
    assert np.array_equal(en0, en1) and np.array_equal(dip0, dip1) and ph0 == ph1 == ref["photon_idx"]
    dq, fl, ph, n_L = handle.rank1_read()
    assert ph == ref["photon_idx"]
    assert n_L == int((s.pos[:, 3].view(np.int64) & 0xFFFFFFFF == s.L_typeid).sum())
    if ph >= 0:
        assert np.array_equal(fl, f_stored[ph, :3])
    rng_ = np.random.default_rng(n_mol)
    net = rng_.normal(size=(s.N, 4)) * 1e-4
    d_net = capi.DeviceArray.from_numpy(net)
    handle.net_force_add_rank1(d_net, d["charge"], d["pos"], s.N, s.L_typeid, G)
    got = d_net.numpy()
    want = net.copy()
    want[:, :3] += f_stored[:, :3]
    assert np.array_equal(got, want)                      # bit-identical to adding the stored GPU force
    assert _rel(got[:, :3] - net[:, :3], ref["force"][:, :3]) <= 1e-10 or np.abs(ref["force"]).max() == 0.0
    assert np.array_equal(got[:, 3], net[:, 3])           # .w (potential energy slot) untouched


@pytest.mark.parametrize("photon", ["last", "duplicated"])
def test_nvt_rank1_steps_bit_identical_to_stored_force_steps(handle, coracle, photon):
    n_mol, steps, dt = 4000, 10, synth.DT_1FS
    s = synth.make_system(n_mol, replica=9, photon=photon)
    dof = 3.0 * n_mol - 3.0
    p = capi.Params.make(OMEGAC, G, PHMASS)
    st = capi.Stream()

    def run(rank1):
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray((s.N, 4), np.float64)
        handle.bussi_reset(st.ptr)
        if rank1:
            handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
        else:
            handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
        handle.bussi_ke(d["vel"], None, 0, n_mol, st.ptr)
        hist = []
        for t in range(steps):
            r, gm = rng.bussi_draws(t, 3, 0, dof)
            a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, dt, dof, r, gm)
            if rank1:
                handle.nvt_step_one_rank1(d["pos"], d["vel"], None, d["charge"], s.N, dt, s.L_typeid, G, 0, n_mol, a, st.ptr)
                handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
                handle.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, 0, n_mol, st.ptr)
            else:
                handle.nvt_step_one(d["pos"], d["vel"], d_f, s.N, dt, 0, n_mol, a, st.ptr)
                handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
                handle.nvt_step_two(d["vel"], d_f, s.N, dt, 0, n_mol, st.ptr)
            b = handle.bussi_read(st.ptr)
            hist.append((b["alpha"], b["ke"], b["cumulative"]))
        return d["pos"].numpy(st.ptr), d["vel"].numpy(st.ptr), hist

    pa, va, ha = run(False)
    pb, vb, hb = run(True)
    assert np.array_equal(pa, pb) and np.array_equal(va, vb) and ha == hb


def test_nvt_rank1_with_other_forces(handle, coracle):
    """force_other != NULL: the kick uses force_other + F_i (one add), as a net-force sum would."""
    n_mol, dt = 3000, 5.0
    s = synth.make_system(n_mol, replica=4)
    p = capi.Params.make(OMEGAC, G, PHMASS)
    other = np.random.default_rng(1).normal(size=(s.N, 4)) * 1e-5
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_o = capi.DeviceArray.from_numpy(other)
    handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p)
    handle.nvt_step_one_rank1(d["pos"], d["vel"], d_o, d["charge"], s.N, dt, s.L_typeid, G, 0, n_mol)
    fc = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G, PHMASS)["force"]
    ftot = other[:, :3] + fc[:, :3]
    hm = (0.5 * dt / s.vel[:, 3])[:, None]
    v = s.vel[:, :3] + hm * ftot
    r = s.pos[:, :3] + dt * v
    assert _rel(d["vel"].numpy()[:, :3], v) <= 1e-12 and _rel(d["pos"].numpy()[:, :3], r) <= 1e-14
    handle.nvt_step_two_rank1(d["vel"], d_o, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, 0, n_mol)
    v2 = v + hm * ftot  # same Final record (no new force_rank1 in between)
    assert _rel(d["vel"].numpy()[:, :3], v2) <= 1e-12
    ke = 0.5 * np.sum(s.vel[:n_mol, 3] * np.sum(v2[:n_mol] ** 2, axis=1))
    assert abs(handle.bussi_read()["ke"] - ke) <= 1e-12 * ke


# ---- step one with the next force's reduce inside (cavb200_md_step_one) ----------------------------
@pytest.mark.parametrize("photon", ["last", "middle", "duplicated"])
@pytest.mark.parametrize("n_mol", [1, 300, 20000, 200000])
def test_md_step_one_matches_oracle_and_rank1_path(handle, coracle, n_mol, photon):
    """cavb200_md_step_one ; cavb200_nvt_step_two_rank1 against orc_nvt_step (same tolerances as above) and
    against the three-launch rank-1 path: identical trajectories to <= 1e-14 (the dipole of the new positions
    is the same compensated sum over another partition of the particles)."""
    steps, dt = 8, synth.DT_1FS
    s = synth.make_system(n_mol, replica=13, photon=photon)
    first, n = (0, n_mol) if photon == "last" else (0, s.N)
    dof = max(3.0 * n - 3.0, 1.0)
    p = capi.Params.make(OMEGAC, G, PHMASS)
    st = capi.Stream()

    def run(fused):
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        handle.bussi_reset(st.ptr)
        handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
        handle.bussi_ke(d["vel"], None, first, n, st.ptr)
        hist = []
        for t in range(steps):
            r, gm = rng.bussi_draws(t, 5, 0, dof)
            a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, dt, dof, r, gm)
            if fused:
                handle.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n,
                                   a, st.ptr)
            else:
                handle.nvt_step_one_rank1(d["pos"], d["vel"], None, d["charge"], s.N, dt, s.L_typeid, G, first, n, a, st.ptr)
                handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
            handle.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, first, n, st.ptr)
            b = handle.bussi_read(st.ptr)
            en, dip, ph = handle.force_read(st.ptr)
            hist.append((b["alpha"], b["ke"], b["cumulative"], en, dip, ph))
        return d["pos"].numpy(st.ptr), d["vel"].numpy(st.ptr), hist

    pa, va, ha = run(False)
    pb, vb, hb = run(True)
    assert _rel(pb[:, :3], pa[:, :3]) <= 1e-14 and _rel(vb[:, :3], va[:, :3]) <= 1e-14
    for x, y in zip(ha, hb):
        assert abs(x[0] - y[0]) <= 1e-14 * abs(x[0]) and abs(x[1] - y[1]) <= 1e-13 * abs(x[1]) and x[5] == y[5]
        assert np.allclose(x[3], y[3], rtol=1e-13, atol=0) and np.allclose(x[4], y[4], rtol=1e-12, atol=1e-9)
    # oracle arm
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    force[:] = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G, PHMASS)["force"]
    idx = np.arange(first, first + n, dtype=np.uint32)
    ke = np.array([coracle.kinetic_energy(vel, idx)])
    res = np.zeros(2)
    for t in range(steps):
        r, gm = rng.bussi_draws(t, 5, 0, dof)
        a_ref, en_ref = coracle.nvt_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, OMEGAC, G, PHMASS, dt, first, n,
                                         dof, synth.KT_100K, synth.TAU_5PS, r, gm, res, ke)
        assert abs(hb[t][0] - a_ref) <= 1e-12 * abs(a_ref)
        assert abs(hb[t][1] - ke[0]) <= 1e-10 * abs(ke[0])
        assert np.allclose(hb[t][3], en_ref, rtol=1e-10, atol=1e-300)
    assert _rel(pb[:, :3], pos[:, :3]) <= 1e-10 and _rel(vb[:, :3], vel[:, :3]) <= 1e-10
    assert np.array_equal(pb[:, 3].view(np.uint64), s.pos[:, 3].view(np.uint64)) and np.array_equal(vb[:, 3], s.vel[:, 3])


def test_md_step_one_argument_errors(handle):
    a = capi.DeviceArray((8, 4), np.float64)
    p = capi.Params.make(OMEGAC, G, PHMASS)
    with pytest.raises(capi.CavbError):
        handle.md_step_one(a, a, None, None, a, 8, 1.0, (1.0, 1.0, 1.0), 2, p, 0, 8)
    handle.md_step_one(None, None, None, None, None, 0, 1.0, (1.0, 1.0, 1.0), 2, p, 0, 0)  # N == 0: no-op


# ---- one launch per MD step (cavb200_md_step_fused) --------------------------------------------------
@pytest.mark.parametrize("photon", ["last", "middle", "duplicated", "absent"])
@pytest.mark.parametrize("n_mol", [1, 300, 20000, 200000])
def test_md_step_fused_matches_two_launch_path_and_oracle(handle, coracle, n_mol, photon):
    """md_step_one ; md_step_fused x (T-1) ; nvt_step_two_rank1  ==  (md_step_one ; nvt_step_two_rank1) x T  (<= 1e-13: the
    kinetic energy is summed over another partition of the particles)  ==  orc_nvt_step x T (tolerances as above)."""
    steps, dt = 8, synth.DT_1FS
    s = synth.make_system(n_mol, replica=17, photon=photon)
    first, n = (0, n_mol) if photon in ("last", "absent") else (0, s.N)
    dof = max(3.0 * n - 3.0, 1.0)
    p = capi.Params.make(OMEGAC, G, PHMASS)
    st = capi.Stream()
    draws = [rng.bussi_draws(t, 9, 0, dof) for t in range(steps)]

    def run(fused):
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        handle.bussi_reset(st.ptr)
        handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
        handle.bussi_ke(d["vel"], None, first, n, st.ptr)
        alphas = []
        for t in range(steps):
            a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, dt, dof, *draws[t])
            if fused and t > 0:
                handle.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n,
                                     a, st.ptr)
            else:
                handle.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n,
                                   a, st.ptr)
            alphas.append(handle.bussi_read(st.ptr)["alpha"])
            if not fused or t == steps - 1:
                handle.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, first, n, st.ptr)
        b = handle.bussi_read(st.ptr)
        en, dip, ph = handle.force_read(st.ptr)
        return d["pos"].numpy(st.ptr), d["vel"].numpy(st.ptr), alphas, b, en, ph

    pa, va, aa, ba, ena, pha = run(False)
    pb, vb, ab, bb, enb, phb = run(True)
    assert pha == phb and bb["err"] == 0.0
    assert _rel(pb[:, :3], pa[:, :3]) <= 1e-13 and _rel(vb[:, :3], va[:, :3]) <= 1e-13
    assert np.allclose(aa, ab, rtol=1e-13, atol=0) and abs(ba["ke"] - bb["ke"]) <= 1e-12 * ba["ke"]
    assert abs(ba["cumulative"] - bb["cumulative"]) <= 1e-12 * max(abs(ba["ke"]), 1e-300)
    assert np.allclose(ena, enb, rtol=1e-12, atol=1e-300)
    # oracle
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    force[:] = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G, PHMASS)["force"]
    idx = np.arange(first, first + n, dtype=np.uint32)
    ke = np.array([coracle.kinetic_energy(vel, idx)])
    res = np.zeros(2)
    for t in range(steps):
        a_ref, en_ref = coracle.nvt_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, OMEGAC, G, PHMASS, dt, first, n,
                                         dof, synth.KT_100K, synth.TAU_5PS, draws[t][0], draws[t][1], res, ke)
        assert abs(ab[t] - a_ref) <= 1e-12 * abs(a_ref)
    assert _rel(pb[:, :3], pos[:, :3]) <= 1e-10 and _rel(vb[:, :3], vel[:, :3]) <= 1e-10
    assert abs(bb["ke"] - ke[0]) <= 1e-10 * ke[0] and np.allclose(enb, en_ref, rtol=1e-10, atol=1e-300)
    assert np.array_equal(pb[:, 3].view(np.uint64), s.pos[:, 3].view(np.uint64)) and np.array_equal(vb[:, 3], s.vel[:, 3])


def test_md_step_fused_windowed_group_and_other_forces(handle):
    """Thermostatted window + an `other forces` array: fused launch == the two-launch path."""
    n_mol, dt = 30000, 5.0
    s = synth.make_system(n_mol, replica=23)
    first, n = 1000, 20000
    dof = 3.0 * n - 3.0
    p = capi.Params.make(OMEGAC, G, PHMASS)
    other = np.random.default_rng(5).normal(size=(s.N, 4)) * 1e-5
    outs = []
    for fused in (False, True):
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_o = capi.DeviceArray.from_numpy(other)
        handle.bussi_reset()
        handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p)
        handle.bussi_ke(d["vel"], None, first, n)
        for t in range(4):
            a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, dt, dof, *rng.bussi_draws(t, 2, 0, dof))
            if fused and t > 0:
                handle.md_step_fused(d["pos"], d["vel"], d_o, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n, a)
            else:
                handle.md_step_one(d["pos"], d["vel"], d_o, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n, a)
            if not fused or t == 3:
                handle.nvt_step_two_rank1(d["vel"], d_o, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, first, n)
        outs.append((d["pos"].numpy(), d["vel"].numpy(), handle.bussi_read()))
    assert _rel(outs[1][0][:, :3], outs[0][0][:, :3]) <= 1e-13 and _rel(outs[1][1][:, :3], outs[0][1][:, :3]) <= 1e-13
    assert abs(outs[1][2]["alpha"] - outs[0][2]["alpha"]) <= 1e-13


def test_bussi_conserved_quantity_one_launch_per_step(handle, coracle):
    """Bussi's conserved quantity: system energy + energy handed to the reservoir (the cumulative KE (1 - alpha^2) of
    reference src/BussiReservoirThermostat.h:86-95) stays constant up to the integrator's own error.  2000 thermostatted
    steps through cavb200_md_step_fused; the same run through the oracle gives the yardstick."""
    n_mol, steps, dt = 2000, 2000, 5.0
    s = synth.make_system(n_mol, replica=31, images=False)
    dof = 3.0 * n_mol - 3.0
    kT, tau = synth.KT_100K, 200.0  # strong coupling so that the reservoir term is large
    p = capi.Params.make(OMEGAC, G, PHMASS)
    draws = [rng.bussi_draws(t, 4, 0, dof) for t in range(steps)]

    def total(vel, en):
        return 0.5 * np.sum(vel[:, 3] * np.sum(vel[:, :3] ** 2, axis=1)) + en.sum()

    # oracle
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    e0 = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G, PHMASS)
    force[:] = e0["force"]
    ke = np.array([coracle.kinetic_energy(vel, np.arange(n_mol, dtype=np.uint32))])
    res = np.zeros(2)
    H_cpu = [total(vel, e0["energies"])]
    for t in range(steps):
        _, en = coracle.nvt_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, OMEGAC, G, PHMASS, dt, 0, n_mol, dof, kT,
                                 tau, draws[t][0], draws[t][1], res, ke)
        if (t + 1) % 200 == 0:
            H_cpu.append(total(vel, en) + res[0])
    # GPU, one launch per step
    st = capi.Stream()
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    handle.bussi_reset(st.ptr)
    handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
    handle.bussi_ke(d["vel"], None, 0, n_mol, st.ptr)
    H_gpu = [total(s.vel, handle.force_read(st.ptr)[0])]
    for t in range(steps):
        a = capi.BussiArgs(kT, tau, dt, dof, *draws[t])
        if t == 0:
            handle.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, 0, n_mol, a, st.ptr)
        else:
            handle.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, 0, n_mol, a,
                                 st.ptr)
        if (t + 1) % 200 == 0:
            v_copy = capi.DeviceArray.from_numpy(d["vel"].numpy(st.ptr))
            handle.nvt_step_two_rank1(v_copy, None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, 0, n_mol, st.ptr)
            b = handle.bussi_read(st.ptr)
            H_gpu.append(total(v_copy.numpy(st.ptr), handle.force_read(st.ptr)[0]) + b["cumulative"])
    H_cpu, H_gpu = np.array(H_cpu), np.array(H_gpu)
    dev_cpu = np.abs(H_cpu - H_cpu[0]).max() / abs(H_cpu[0])
    dev_gpu = np.abs(H_gpu - H_gpu[0]).max() / abs(H_gpu[0])
    print(f"Bussi conserved quantity over {steps} steps: relative deviation cpu {dev_cpu:.3e} gpu {dev_gpu:.3e}; reservoir {res[0]:.4e}")
    assert abs(res[0]) > 1e-3 * abs(H_cpu[0])          # the thermostat really exchanged energy
    assert dev_gpu <= 1.05 * dev_cpu + 1e-12
    assert np.abs(H_gpu - H_cpu).max() <= 1e-8 * abs(H_cpu[0])
