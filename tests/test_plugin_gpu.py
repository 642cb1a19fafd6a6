"""The plugin glue (plugin/src: CavityForceComputeGPU, BussiReservoirThermostat -- the pybind11 /
HOOMD host classes that call the C ABI) built against hoomd_shim and driven the way HOOMD's
integrator drives them: ForceCompute.compute(timestep) and Thermostat.getRescalingFactorsOne(
timestep, dt).  Same class names, constructors and methods as the reference's _cavitymd /
_bussi_reservoir modules."""
import os
import sys

import numpy as np
import pytest

from cav_hoomd_b200 import synth

pytestmark = pytest.mark.gpu
BUILD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plugin", "build")


@pytest.fixture(scope="module")
def mods():
    sys.path.insert(0, BUILD)
    import _hoomd_shim  # noqa: F401  (registers ForceCompute / Thermostat bases first)
    import _bussi_reservoir
    import _cavitymd
    return _hoomd_shim, _cavitymd, _bussi_reservoir


def make_sysdef(shim, s, seed=7):
    exec_conf = shim.ExecutionConfiguration(True, 0)
    pd = shim.ParticleData(s.N, s.box[0], s.box[1], s.box[2], list(s.types), exec_conf)
    pd.setPositions(s.pos)
    pd.setVelocities(s.vel)
    pd.setCharges(s.charge)
    pd.setImages(s.image)
    return shim.SystemDefinition(pd, seed), pd


def test_cavity_force_compute_gpu_class(mods, coracle):
    shim, cav, _ = mods
    s = synth.make_system(30000)
    sysdef, pd = make_sysdef(shim, s)
    fc = cav.CavityForceComputeGPU(sysdef, 0.01, 1e-3)  # (sysdef, omegac, couplstr, phmass=1.0)
    p = fc.getParams()
    assert p["omegac"] == 0.01 and p["couplstr"] == 1e-3 and p["phmass"] == 1.0 and p["K"] == 1.0 * 0.01 * 0.01
    assert fc.getHarmonicEnergy() == 0.0  # before the first compute
    fc.compute(0)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    f = fc.getForces()
    assert np.abs(f - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
    assert np.isclose(fc.getHarmonicEnergy(), ref["energies"][0], rtol=1e-10)
    assert np.isclose(fc.getCouplingEnergy(), ref["energies"][1], rtol=1e-10)
    assert np.isclose(fc.getDipoleSelfEnergy(), ref["energies"][2], rtol=1e-10)
    assert fc.calcEnergySum() == 0.0  # per-particle PE stays 0 (reference CavityForceCompute.cc:178-180)
    # setParams takes effect on the next compute
    fc.setParams(0.01, 2e-3, 1.0)
    fc.compute(1)
    ref2 = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 2e-3)
    assert np.isclose(fc.getCouplingEnergy(), ref2["energies"][1], rtol=1e-10)


def test_cavity_force_requires_gpu_exec_conf(mods):
    shim, cav, _ = mods
    s = synth.make_system(10)
    pd = shim.ParticleData(s.N, 1, 1, 1, list(s.types), shim.ExecutionConfiguration(False, 0))
    with pytest.raises(RuntimeError):
        cav.CavityForceComputeGPU(shim.SystemDefinition(pd, 0), 0.01, 1e-3)  # no CPU fallback


@pytest.mark.parametrize("fused", [False, True])
def test_bussi_reservoir_thermostat_class(mods, coracle, fused):
    shim, _, bus = mods
    n = 4000
    s = synth.make_system(n)
    sysdef, pd = make_sysdef(shim, s)
    group = shim.ParticleGroup(sysdef, list(range(n)))
    dof = 3.0 * n - 3.0
    group.setTranslationalDOF(dof)
    thermo = shim.ComputeThermo(sysdef, group)
    kT = shim.VariantConstant(synth.KT_100K)
    th = bus.BussiReservoirThermostat(kT, group, thermo, sysdef, synth.TAU_5PS)
    assert th.tau == synth.TAU_5PS and th.kT(0) == synth.KT_100K
    assert th.getTotalReservoirEnergy() == 0.0 and th.getReservoirEnergyRotational() == 0.0
    th.fused_rescale = fused
    assert th.getRescalingFactorsOne(0, 0.0) == [1.0, 1.0]  # deltaT == 0 (reference :45-48)

    vref = s.vel.copy()
    res = np.zeros(2)
    idx = np.arange(n, dtype=np.uint32)
    rng = np.random.default_rng(3)
    for t in range(4):
        r, g = rng.standard_normal(), rng.gamma((dof - 1) / 2)
        bus._inject_draws([r, g])
        factors = th.getRescalingFactorsOne(t, synth.DT_1FS)
        alpha, ke = coracle.bussi_step(vref, idx, dof, synth.DT_1FS, synth.KT_100K, synth.TAU_5PS, r, g, res)
        if fused:
            assert factors == [1.0, 1.0]  # the kernel already rescaled
        else:
            assert abs(factors[0] - alpha) <= 1e-12 * abs(alpha) and factors[1] == 1.0
            v = pd.getVelocities()  # what HOOMD's step one does with the factor
            v[:n, :3] *= factors[0]
            pd.setVelocities(v)
        assert np.isclose(th.getInstantaneousReservoirTranslational(), res[1], rtol=1e-9, atol=1e-18)
    assert np.isclose(th.getTotalReservoirEnergy(), res[0], rtol=1e-9, atol=1e-18)
    assert np.allclose(pd.getVelocities(), vref, rtol=1e-11, atol=0)
    th.resetReservoirEnergy()
    assert th.getTotalReservoirEnergy() == 0.0 and th.getInstantaneousReservoirTotal() == 0.0


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("case", ["rot_n200", "rot_n64_rdof1", "rot_n128_tau0"])
def test_bussi_thermostat_rotational_dof(mods, case, fused):
    """Rotational degrees of freedom: alpha_r and the rotational reservoir from the same generator, after the
    translational draws (reference src/BussiReservoirThermostat.h:53-55,77-81,87-95) -- against vectors minted from
    the reference's own translation unit (tests/golden/make_golden.py)."""
    shim, _, bus = mods
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bussi.npz"))
    dof, kT, tau, dt, rdof, rke = g[f"{case}/args"]
    vel0 = g[f"{case}/vel0"]
    n = vel0.shape[0] - 1
    s = synth.make_system(n, replica=6)
    assert np.array_equal(s.vel, vel0)
    sysdef, pd = make_sysdef(shim, s)
    group = shim.ParticleGroup(sysdef, list(range(n)))
    group.setTranslationalDOF(dof)
    group.setRotationalDOF(rdof)
    thermo = shim.ComputeThermo(sysdef, group)
    th = bus.BussiReservoirThermostat(shim.VariantConstant(kT), group, thermo, sysdef, tau)
    th.fused_rescale = fused
    for k, d in enumerate(g[f"{case}/draws"]):
        row = g[f"{case}/rows"][k]
        thermo.setRotationalKineticEnergy(rke * (1 + 0.1 * k))
        bus._inject_draws([float(x) for x in d])
        factors = th.getRescalingFactorsOne(k, dt)
        assert factors[1] == row[1]  # host formula, the reference's operation order: bitwise
        assert th.getInstantaneousReservoirRotational() == row[6] and th.getReservoirEnergyRotational() == row[5]
        if fused:
            assert factors[0] == 1.0
        else:
            assert abs(factors[0] - row[0]) <= 1e-12 * abs(row[0])
        # (the golden rows never rescale the velocities, so the translational KE is the same every step)
        assert np.isclose(th.getInstantaneousReservoirTranslational(), row[4], rtol=1e-9, atol=1e-18)
        if fused:  # undo the in-kernel rescale to stay on the golden trajectory
            pd.setVelocities(s.vel)
    # zero rotational kinetic energy with rotational dof: the reference throws (:57-61)
    thermo.setRotationalKineticEnergy(0.0)
    bus._inject_draws([0.1, 1.0, 0.1, 1.0])
    with pytest.raises(RuntimeError, match="non-zero initial momenta"):
        th.getRescalingFactorsOne(9, dt)


def test_bussi_thermostat_zero_momenta_throws(mods):
    shim, _, bus = mods
    s = synth.make_system(100)
    s.vel[:, :3] = 0.0
    sysdef, pd = make_sysdef(shim, s)
    group = shim.ParticleGroup(sysdef, list(range(100)))
    group.setTranslationalDOF(297.0)
    th = bus.BussiReservoirThermostat(shim.VariantConstant(1.0), group, shim.ComputeThermo(sysdef, group), sysdef, 1.0)
    bus._inject_draws([0.1, 140.0])
    with pytest.raises(RuntimeError, match="non-zero initial momenta"):
        th.getRescalingFactorsOne(0, 0.005)


def test_cavity_force_class_device_side_tracker(mods):
    """trackOpen / trackSetReference / trackRecord / trackRead on the plugin class: the dipole autocorrelation and the
    cavity-mode kinetic energy without a snapshot (reference src/cavitymd/analysis.py:18-31,222-224,1352-1354)."""
    from oracle import oracle as O
    shim, cav, _ = mods
    s = synth.make_system(5000, replica=6)
    sysdef, pd = make_sysdef(shim, s)
    fc = cav.CavityForceComputeGPU(sysdef, 0.01, 1e-3)
    fc.trackOpen(16)
    fc.compute(0)
    fc.trackSetReference()
    d0 = O.numpy_total_dipole(s.pos[:, :3], s.image, s.charge, s.box)
    pos = s.pos.copy()
    for t in range(1, 4):
        pos[:, :3] += 0.01 * t
        pd.setPositions(pos)
        fc.compute(t)
        fc.trackRecord(t)
    rows = fc.trackRead(8)
    assert [int(r[0]) for r in rows] == [1, 2, 3] and all(len(r) == 16 for r in rows)
    d3 = O.numpy_total_dipole(pos[:, :3], s.image, s.charge, s.box)
    scale = np.abs(s.charge[:, None] * O.numpy_unwrap(pos[:, :3], s.image, s.box)).sum(axis=0)
    assert np.all(np.abs(np.array(rows[-1][1:4]) - d3) <= 1e-10 * scale)
    assert abs(rows[-1][10] - np.dot(d0, d3)) <= 1e-10 * np.dot(np.abs(d0), scale)
    ke_ph = 0.5 * s.vel[-1, 3] * np.sum(s.vel[-1, :3] ** 2)
    assert abs(rows[-1][11] - ke_ph) <= 1e-12 * ke_ph and int(rows[-1][15]) == s.N - 1
    assert rows[-1][7] == fc.getHarmonicEnergy()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("cooperative", [False, True])
def test_two_force_objects_on_unordered_streams(mods, coracle, cooperative):
    """Two CavityForceComputeGPU objects driven on two unordered streams (HOOMD itself uses one stream; this is the
    shared-GPU situation).  With cooperative_launch every compute is correct.  With the default launches a grid that
    was not co-resident is REPORTED -- the next compute of that object throws once, the object then launches
    cooperatively -- and every compute after the report is correct; nothing hangs and no wrong force passes as good."""
    from cav_hoomd_b200 import capi
    shim, cav, _ = mods
    s = synth.make_system(400000)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    scale = np.abs(ref["force"]).max()
    objs, streams = [], [capi.Stream(), capi.Stream()]
    for k in range(2):
        sysdef, pd = make_sysdef(shim, s)
        fc = cav.CavityForceComputeGPU(sysdef, 0.01, 1e-3)
        fc.stream = streams[k].ptr
        fc.cooperative_launch = cooperative
        objs.append((fc, sysdef, pd))
    thrown = 0
    for t in range(10):
        for fc, _, _ in objs:
            try:
                fc.compute(t)
            except RuntimeError as e:
                assert "cooperative" in str(e) and not cooperative
                thrown += 1
                assert fc.cooperative_launch and fc.getFaultCount() >= 1
                fc.compute(t)  # redone, now co-scheduled by the driver
        for fc, _, _ in objs:
            try:
                e_h = fc.getHarmonicEnergy()  # synchronises the object's stream
            except RuntimeError:
                assert not cooperative and not fc.cooperative_launch  # this very call timed out: reported, not wrong
                continue
            assert np.isclose(e_h, ref["energies"][0], rtol=1e-10)
            f = fc.getForces()
            assert np.abs(f - ref["force"]).max() <= 1e-10 * scale
    if cooperative:
        assert thrown == 0 and all(fc.getFaultCount() == 0 for fc, _, _ in objs)
    capi.sync()


@pytest.mark.parametrize("mode", ["stored", "rank1", "one_launch"])
@pytest.mark.parametrize("n_mol", [300, 40000])
def test_two_step_constant_volume_cavity(mods, coracle, mode, n_mol):
    """The fused integration method (plugin/src/TwoStepConstantVolumeCavity) driven the way HOOMD's
    IntegratorTwoStep::update drives a method -- integrateStepOne(t); every force computes; net force; integrateStepTwo(t)
    -- with the thermostat's draws made through the BussiReservoirThermostat object and the cavity force through the
    CavityForceComputeGPU object.  Against orc_nvt_step_w (alpha-rescale, kick, drift, box wrap + image update, cavity
    force, kick) on a box where a large share of the particles crosses a face."""
    from cav_hoomd_b200 import rng as crng
    shim, cav, bus = mods
    s = synth.make_system(n_mol, replica=51)
    Lbox = s.box[0]
    mol = s.typeid != s.L_typeid
    s.vel[mol, :3] *= (Lbox / 6.0) / (np.abs(s.vel[mol, :3]).max() * synth.DT_1FS)
    steps, dt = 6, synth.DT_1FS
    dof = 3.0 * n_mol - 3.0
    sysdef, pd = make_sysdef(shim, s)
    group = shim.ParticleGroup(sysdef, list(range(n_mol)))
    group.setTranslationalDOF(dof)
    th = bus.BussiReservoirThermostat(shim.VariantConstant(synth.KT_100K), group, shim.ComputeThermo(sysdef, group), sysdef,
                                      synth.TAU_5PS)
    g = 1e-8  # weak coupling: the fast charges of this box would otherwise be thrown many box lengths per step
    fc = cav.CavityForceComputeGPU(sysdef, 0.01, g)
    method = cav.TwoStepConstantVolumeCavity(sysdef, group, th, fc, mode)
    assert method.mode == mode
    method.setDeltaT(dt)
    draws = [crng.bussi_draws(t, 11, 0, dof) for t in range(steps)]

    # Integrator::prepRun: the forces of the initial configuration
    fc.compute(0)
    net = fc.getForces()
    if mode != "stored":
        assert not net.any()  # rank-1: nothing is stored, the contribution to the net force is zero
    pd.setNetForce(net)
    for t in range(steps):
        bus._inject_draws(list(draws[t]))
        method.integrateStepOne(t)
        fc.compute(t + 1)
        pd.setNetForce(fc.getForces())
        method.integrateStepTwo(t)
    method.flush()

    # oracle
    pos, vel, image, force = s.pos.copy(), s.vel.copy(), s.image.copy(), np.zeros((s.N, 4))
    force[:] = coracle.cavity_force(pos, s.charge, image, s.box, s.L_typeid, 0.01, g)["force"]
    ke = np.array([coracle.kinetic_energy(vel, np.arange(n_mol, dtype=np.uint32))])
    res = np.zeros(2)
    for t in range(steps):
        a_ref, en_ref = coracle.nvt_step(pos, vel, s.charge, image, force, s.box, s.L_typeid, 0.01, g, 1.0, dt, 0, n_mol, dof,
                                         synth.KT_100K, synth.TAU_5PS, draws[t][0], draws[t][1], res, ke, wrap=True)
    gp, gv, gi = pd.getPositions(), pd.getVelocities(), pd.getImages()
    assert np.any(image != s.image, axis=1).sum() > 0.2 * s.N and np.array_equal(gi, image)
    assert np.abs(gp[:, :3] - pos[:, :3]).max() <= 1e-10 * np.abs(pos[:, :3]).max()
    assert np.abs(gv[:, :3] - vel[:, :3]).max() <= 1e-10 * np.abs(vel[:, :3]).max()
    # the reference's getters keep working: energies from the force object, reservoir from the thermostat object
    assert np.isclose(fc.getHarmonicEnergy(), en_ref[0], rtol=1e-10) and np.isclose(fc.getCouplingEnergy(), en_ref[1], rtol=1e-10)
    assert np.isclose(th.getReservoirEnergyTranslational(), res[0], rtol=1e-9, atol=1e-18)
    assert np.isclose(th.getInstantaneousReservoirTranslational(), res[1], rtol=1e-9, atol=1e-18)
    launches = method.getLaunchCount()
    # stored / rank1: KE once, then two launches per step (+ the force object's own); one_launch: KE, then ONE per step + flush
    assert launches == {"stored": 1 + 2 * steps, "rank1": 1 + 2 * steps + steps + 1, "one_launch": 1 + steps + 1 + 1}[mode]


def test_two_step_method_argument_checks(mods):
    shim, cav, bus = mods
    s = synth.make_system(100)
    sysdef, pd = make_sysdef(shim, s)
    fc = cav.CavityForceComputeGPU(sysdef, 0.01, 1e-3)
    gap = shim.ParticleGroup(sysdef, [0, 1, 2, 4, 5])  # not a contiguous range
    gap.setTranslationalDOF(12.0)
    with pytest.raises((ValueError, RuntimeError)):
        cav.TwoStepConstantVolumeCavity(sysdef, gap, None, fc, "bogus")
    with pytest.raises((ValueError, RuntimeError)):
        cav.TwoStepConstantVolumeCavity(sysdef, gap, None, None, "rank1")
    m = cav.TwoStepConstantVolumeCavity(sysdef, gap, None, fc, "stored")
    m.setDeltaT(1.0)
    pd.setNetForce(np.zeros((s.N, 4)))
    with pytest.raises(RuntimeError, match="contiguous"):
        m.integrateStepOne(0)
    # no thermostat: plain velocity Verlet with the wrap
    full = shim.ParticleGroup(sysdef, list(range(100)))
    m2 = cav.TwoStepConstantVolumeCavity(sysdef, full, None, fc, "stored")
    m2.setDeltaT(1.0)
    v0, p0 = pd.getVelocities(), pd.getPositions()
    m2.integrateStepOne(0)
    m2.integrateStepTwo(0)
    assert np.array_equal(pd.getVelocities(), v0)  # zero net force
    L = np.array(s.box)
    u = pd.getPositions()[:, :3] + pd.getImages() * L
    assert np.allclose(u, p0[:, :3] + s.image * L + v0[:, :3], rtol=1e-13, atol=1e-9)
