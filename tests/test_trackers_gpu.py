"""Device-side trackers: the per-step records of cavb200_track_record against NumPy
restatements of the reference's snapshot-based trackers (compute_total_dipole_moment, reference
src/cavitymd/analysis.py:18-31; AutocorrelationTracker :152-258; CavityModeTracker :1285-1417) evaluated on
snapshots downloaded in the test.  Tolerance: 1e-10 relative to sum_i |q_i r_i| for the dipole (np.dot's
summation order is BLAS's, ours is a compensated tree), 1e-10 relative for derived quantities."""
import numpy as np
import pytest

from cav_hoomd_b200 import analysis, capi, rng, state, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

OMEGAC, G = 0.01, 1e-3


def _typeid(pos):
    return (pos[:, 3].view(np.int64) & 0xFFFFFFFF).astype(np.int64)


@pytest.mark.parametrize("photon", ["last", "middle", "absent"])
def test_track_records_match_snapshot_trackers(handle, photon):
    n_mol, steps, dt = 5000, 12, synth.DT_1FS
    s = synth.make_system(n_mol, replica=21, photon=photon)
    st8 = state.DeviceState(s)
    p = capi.Params.make(OMEGAC, G)
    dof = 3.0 * n_mol - 3.0
    handle.bussi_reset()
    handle.track_open(8)  # smaller than the number of steps: the ring wraps
    handle.force(st8.pos, st8.charge, st8.image, st8.force, s.N, s.box, s.L_typeid, p)
    handle.bussi_ke(st8.vel, None, 0, n_mol)
    handle.track_set_reference()
    snap0 = st8.snapshot()
    d0 = O.numpy_total_dipole(snap0.pos[:, :3], snap0.image, snap0.charge, s.box)
    want = []
    for t in range(1, steps + 1):
        r, gm = rng.bussi_draws(t, 1, 0, dof)
        handle.nvt_step_one(st8.pos, st8.vel, st8.force, s.N, dt, 0, n_mol, capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, dt, dof, r, gm))
        handle.force(st8.pos, st8.charge, st8.image, st8.force, s.N, s.box, s.L_typeid, p)
        handle.nvt_step_two(st8.vel, st8.force, s.N, dt, 0, n_mol)
        handle.track_record(t, st8.vel, s.N)
        # what the reference's trackers would compute from a snapshot at this step
        snap = st8.snapshot()
        d = O.numpy_total_dipole(snap.pos[:, :3], snap.image, snap.charge, s.box)
        en = handle.force_read()[0]
        cav = O.numpy_cavity_mode(snap.pos[:, :3], snap.image, snap.vel[:, :3], snap.vel[:, 3], _typeid(snap.pos), s.box, en[0])
        scale = np.abs(snap.charge[:, None] * O.numpy_unwrap(snap.pos[:, :3], snap.image, s.box)).sum(axis=0)
        want.append((t, d, np.dot(d0, d), cav, en, scale, handle.bussi_read()))
    rec, total = handle.track_read(100)
    assert total == steps and len(rec) == 8                      # ring kept the newest 8
    assert np.array_equal(rec[:, 0], np.arange(steps - 7, steps + 1))
    for row, (t, d, c, cav, en, scale, bo) in zip(rec, want[-8:]):
        if photon == "absent":
            # no 'L' particle: the cavity force computes nothing (reference CavityForceCompute.cc:149-156),
            # so there is no dipole on the device to record -- documented in include/cavb200.h
            assert np.all(row[1:11] == 0.0) and row[11] == 0.0 and row[15] == -1
            continue
        assert np.all(np.abs(row[1:4] - d) <= 1e-10 * scale)
        assert abs(row[10] - c) <= 1e-10 * np.dot(np.abs(d0), scale)
        assert np.array_equal(row[7:10], en)
        assert abs(row[11] - cav[0]) <= 1e-12 * max(cav[0], 1e-300)
        assert row[12] == bo["ke"] and row[13] == bo["alpha"] and row[14] == bo["cumulative"]
        assert row[15] == (s.N - 1 if photon == "last" else (-1 if photon == "absent" else row[15]))
    rec2, _ = handle.track_read(3)
    assert np.array_equal(rec2, rec[-3:])


def test_tracker_mirrors(handle, tmp_path):
    """DipoleAutocorrelation / CavityModeTracker host mirrors: reference file format, new reference on a
    multiple of 10000 steps (analysis.py:247-249), logged quantities."""
    n_mol, dt = 2000, 5.0
    s = synth.make_system(n_mol, replica=8)
    st8 = state.DeviceState(s)
    p = capi.Params.make(OMEGAC, G)
    handle.force(st8.pos, st8.charge, st8.image, st8.force, s.N, s.box, s.L_typeid, p)
    recd = analysis.StepRecorder(st8, handle, capacity=64)
    dac = analysis.DipoleAutocorrelation(recd, dt, output_prefix=str(tmp_path / "dipole_autocorr"), output_period_steps=5)
    cav = analysis.CavityModeTracker(recd)
    snap = st8.snapshot()
    d0 = O.numpy_total_dipole(snap.pos[:, :3], snap.image, snap.charge, s.box)
    assert abs(dac.current_autocorr - np.dot(d0, d0)) <= 1e-10 * np.dot(d0, d0)
    for t in range(1, 21):
        handle.nve_kick_drift(st8.pos, st8.vel, st8.force, s.N, dt)
        handle.force(st8.pos, st8.charge, st8.image, st8.force, s.N, s.box, s.L_typeid, p)
        handle.nve_half_kick(st8.vel, st8.force, s.N, dt)
        dac.act(t)
    snap = st8.snapshot()
    d = O.numpy_total_dipole(snap.pos[:, :3], snap.image, snap.charge, s.box)
    assert abs(dac.current_autocorr - np.dot(d0, d)) <= 1e-9 * abs(np.dot(d0, d0))
    lines = open(str(tmp_path / "dipole_autocorr_0.txt")).read().splitlines()
    assert lines[0] == "# Dipole autocorrelation data" and lines[3] == "# timestep t(ps) C(t)"
    assert [int(l.split()[0]) for l in lines[4:]] == [0, 5, 10, 15, 20]
    cav.refresh()
    ke = 0.5 * snap.vel[-1, 3] * np.sum(snap.vel[-1, :3] ** 2)
    assert abs(cav.cavity_kinetic_energy - ke) <= 1e-12 * ke
    assert cav.cavity_potential_energy_harmonic == handle.force_read()[0][0]
    assert abs(cav.cavity_temperature - (2.0 / 3.0) * ke / 3.167e-6) <= 1e-9 * cav.cavity_temperature
    assert cav.cavity_total_energy == cav.cavity_kinetic_energy + cav.cavity_potential_energy_harmonic


def test_track_argument_errors(handle):
    h2 = capi.Handle(0)
    with pytest.raises(capi.CavbError):
        h2.track_record(0, None, 0)  # ring not opened
    with pytest.raises(capi.CavbError):
        h2.track_open(0)
    h2.track_open(4)
    rec, total = h2.track_read(4)
    assert len(rec) == 0 and total == 0
    h2.close()
