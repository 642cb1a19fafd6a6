"""The host-side mirrors of the reference's Python wrappers (CavityForce, BussiReservoir, F(k,t)
tracker) on a GPU: written like the reference's own tests (src/pytest/test_bussi_reservoir.py) plus
the value checks those tests never made."""
import numpy as np
import pytest

from cav_hoomd_b200 import BussiReservoir, CavityForce, DensityField, DeviceState, FieldAutocorrelationTracker, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_cavity_force_wrapper(coracle):
    s = synth.make_system(5000)
    state = DeviceState(s)
    cf = CavityForce(kvector=[0, 0, 1], couplstr=1e-3, omegac=0.01, phmass=1.0)
    assert cf.implementation == "cuda"
    with pytest.raises(RuntimeError):
        cf.compute()  # not attached
    cf._attach(state)
    with pytest.raises(RuntimeError):
        _ = cf.harmonic_energy  # requires_run
    cf.compute(0)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert np.isclose(cf.harmonic_energy, ref["energies"][0], rtol=1e-10)
    assert np.isclose(cf.coupling_energy, ref["energies"][1], rtol=1e-10)
    assert np.isclose(cf.dipole_self_energy, ref["energies"][2], rtol=1e-10)
    assert np.isclose(cf.total_cavity_energy, ref["energies"].sum(), rtol=1e-10)
    assert cf.energy == cf.total_cavity_energy
    assert np.abs(cf.forces - ref["force"][:, :3]).max() <= 1e-10 * np.abs(ref["force"]).max()
    cf._detach()


def test_cavity_force_without_L_type():
    """A system with no type named 'L': zero energies, zero forces (reference GPU.cc:114-123)."""
    s = synth.make_system(100, photon="absent")
    s.types = ("O", "N")
    state = DeviceState(s)
    cf = CavityForce([0, 0, 1], 1e-3, 0.01)._attach(state)
    cf.compute(0)
    assert cf.total_cavity_energy == 0.0 and np.all(cf.forces == 0.0)


def test_bussi_reservoir_wrapper_like_the_reference_tests(coracle):
    """reference src/pytest/test_bussi_reservoir.py:11-76 -- kT / tau round trip, zeros before the
    first step and after reset -- plus the values the reference only prints."""
    s = synth.make_system(100)
    state = DeviceState(s)
    state.seed = 42
    bussi = BussiReservoir(kT=synth.KT_100K, tau=synth.TAU_5PS)
    assert bussi.kT == synth.KT_100K and bussi.tau == synth.TAU_5PS
    assert bussi.total_reservoir_energy == 0.0  # not attached
    bussi._attach(state)
    assert bussi.reservoir_energy_translational == 0.0 and bussi.reservoir_energy_rotational == 0.0
    assert bussi.instantaneous_reservoir_total == 0.0
    vref = s.vel.copy()
    res = np.zeros(2)
    idx = synth.molecular_group(s)
    from cav_hoomd_b200 import rng
    for t in range(10):
        bussi.rescale(t, synth.DT_1FS)
        r, g = rng.bussi_draws(t, 42, int(idx[0]), bussi.dof)
        alpha, ke = coracle.bussi_step(vref, idx, bussi.dof, synth.DT_1FS, synth.KT_100K, synth.TAU_5PS, r, g, res)
        assert abs(bussi.last_alpha - alpha) <= 1e-12 * abs(alpha)
    assert np.isclose(bussi.total_reservoir_energy, res[0], rtol=1e-9, atol=1e-18)
    assert np.isclose(bussi.instantaneous_reservoir_translational, res[1], rtol=1e-9, atol=1e-18)
    assert np.allclose(state.vel.numpy(), vref, rtol=1e-11, atol=0)
    bussi.reset_reservoir_energy()
    assert bussi.total_reservoir_energy == 0.0 and bussi.instantaneous_reservoir_total == 0.0
    # zero momenta: the reference throws "requires non-zero initial momenta"
    z = s.copy()
    z.vel[:, :3] = 0.0
    b2 = BussiReservoir(kT=1.0, tau=1.0)._attach(DeviceState(z))
    b2.rescale(0, 1.0)
    with pytest.raises(RuntimeError):
        _ = b2.total_reservoir_energy


def test_bussi_wrapper_with_photon_in_the_middle(coracle):
    """filter.Type(['O','N']) is an index list when the photon is not last."""
    s = synth.make_system(2000, photon="middle")
    state = DeviceState(s)
    b = BussiReservoir(kT=synth.KT_100K, tau=0.0)._attach(state)
    b.rescale(3, synth.DT_1FS, draws=(0.25, 2990.0))
    vref = s.vel.copy()
    idx = synth.molecular_group(s)
    alpha, _ = coracle.bussi_step(vref, idx, b.dof, synth.DT_1FS, synth.KT_100K, 0.0, 0.25, 2990.0, np.zeros(2))
    assert abs(b.last_alpha - alpha) <= 1e-12 * abs(alpha)
    assert np.allclose(state.vel.numpy(), vref, rtol=1e-12, atol=0)
    assert np.array_equal(state.vel.numpy()[1000], s.vel[1000])  # the photon is not thermostatted


def test_density_field_and_tracker():
    s = synth.make_system(3000)
    state = DeviceState(s)
    kvec = synth.fibonacci_sphere(50) * 1.0
    field = DensityField(kvec)
    rho = field.compute_density_field(s.pos[:, :3])
    ref = O.numpy_density_field(s.pos[:, :3], kvec)
    assert np.abs(rho - ref).max() <= 1e-13 * s.N + 1e-12
    tr = FieldAutocorrelationTracker(state, kmag=1.0, num_wavevectors=50, reference_interval_steps=2, max_references=3)
    assert tr.act(0) == []
    out = tr.act(1)
    assert len(out) == 1 and out[0][0] == 0
    assert np.isclose(out[0][2], O.numpy_field_autocorr(ref, ref), rtol=1e-10)  # positions unchanged: F(k,0)
    tr.act(2)
    tr.act(4)
    assert len(tr.references) == 3  # references at 0, 2, 4; capped at max_references
    tr.act(6)
    assert len(tr.references) == 3
