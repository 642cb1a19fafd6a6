"""10k-step NVE energy drift: velocity Verlet around the cavity force, GPU arm vs CPU-oracle arm on
the same input (BASELINE.json: "energy-conservation drift over a 10k-step NVE run no worse than
the reference").  Conserved: KE(molecules + photon) + E_harmonic + E_coupling + E_dipole_self."""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, synth

pytestmark = pytest.mark.gpu


def total_energy(vel, energies):
    return 0.5 * np.sum(vel[:, 3] * np.sum(vel[:, :3] ** 2, axis=1)) + energies.sum()


@pytest.mark.parametrize("dt", [synth.DT_1FS, 5.0])
def test_nve_drift_no_worse_than_oracle(handle, coracle, dt):
    """dt = 1 fs is the BASELINE step (omega dt ~ 0.5: a bounded Verlet energy oscillation of a few
    per cent, identical on both arms); dt = 5 a.u. shows the conserved quantity tightly."""
    n_mol, steps, every = 2000, 10000, 250
    omegac, g, phmass = 0.01, 1e-3, 1.0
    s = synth.make_system(n_mol, images=False)
    # --- CPU oracle arm ---
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    e = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, omegac, g, phmass)
    force[:] = e["force"]
    E_cpu = [total_energy(vel, e["energies"])]
    for k in range(steps):
        en = coracle.nve_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, omegac, g, phmass, dt)
        if (k + 1) % every == 0:
            E_cpu.append(total_energy(vel, en))
    # --- GPU arm ---
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    p = capi.Params.make(omegac, g, phmass)
    st = capi.Stream()
    handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
    E_gpu = [total_energy(d["vel"].numpy(st.ptr), handle.force_read(st.ptr)[0])]
    for k in range(steps):
        handle.nve_kick_drift(d["pos"], d["vel"], d_f, s.N, dt, st.ptr)
        handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
        handle.nve_half_kick(d["vel"], d_f, s.N, dt, st.ptr)
        if (k + 1) % every == 0:
            E_gpu.append(total_energy(d["vel"].numpy(st.ptr), handle.force_read(st.ptr)[0]))
    E_cpu, E_gpu = np.array(E_cpu), np.array(E_gpu)
    drift_cpu = np.abs(E_cpu - E_cpu[0]).max() / abs(E_cpu[0])
    drift_gpu = np.abs(E_gpu - E_gpu[0]).max() / abs(E_gpu[0])
    print(f"NVE 10k steps dt={dt}: relative drift cpu {drift_cpu:.3e} gpu {drift_gpu:.3e}; |E_gpu-E_cpu|max "
          f"{np.abs(E_gpu - E_cpu).max() / abs(E_cpu[0]):.3e}")
    assert drift_gpu <= 1.02 * drift_cpu + 1e-12
    # same integrator arithmetic on both arms: the trajectories stay together
    assert np.abs(E_gpu - E_cpu).max() <= 1e-9 * abs(E_cpu[0])
    assert np.allclose(d["pos"].numpy(st.ptr)[:, :3], pos[:, :3], rtol=1e-8, atol=1e-8)


def test_nve_drift_large_system(handle, coracle):
    """The same comparison on a system that fills the persistent grids (262 145 particles: past every failure threshold
    of the reference's own GPU kernel, SURVEY.md Appendix C, and large enough for the folder / full-grid paths): 1000
    steps, the step size scaled to the collective frequency of the bigger box (omega^2 ~ g^2 sum c^2 / (K m) grows with
    N; dt = 1 a.u. keeps omega dt ~ 0.15 as dt = 5 a.u. does at 2000 particles).  Three integrators: the three-call
    harness, the same with the box wrap, and one persistent launch per step."""
    n_mol, steps, every, dt = 262144, 1000, 100, 1.0
    omegac, g, phmass = 0.01, 1e-3, 1.0
    s = synth.make_system(n_mol, replica=2, images=False)
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    e = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, omegac, g, phmass)
    force[:] = e["force"]
    E_cpu = [total_energy(vel, e["energies"])]
    for k in range(steps):
        en = coracle.nve_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, omegac, g, phmass, dt)
        if (k + 1) % every == 0:
            E_cpu.append(total_energy(vel, en))
    E_cpu = np.array(E_cpu)
    drift_cpu = np.abs(E_cpu - E_cpu[0]).max() / abs(E_cpu[0])
    p = capi.Params.make(omegac, g, phmass)
    st = capi.Stream()
    for path in ("three_calls", "three_calls_wrap", "one_launch"):
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray((s.N, 4), np.float64)
        if path == "one_launch":
            handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
            E_gpu = [total_energy(d["vel"].numpy(st.ptr), handle.force_read(st.ptr)[0])]
            handle.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, 0, 0, None, st.ptr)
            for k in range(steps):
                if (k + 1) % every == 0:
                    v_copy = capi.DeviceArray.from_numpy(d["vel"].numpy(st.ptr))
                    handle.nvt_step_two_rank1(v_copy, None, d["charge"], d["pos"], s.N, dt, s.L_typeid, g, 0, 0, st.ptr)
                    E_gpu.append(total_energy(v_copy.numpy(st.ptr), handle.force_read(st.ptr)[0]))
                if k < steps - 1:
                    handle.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, 0, 0,
                                         None, st.ptr)
        else:
            wrap = path.endswith("wrap")
            handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
            E_gpu = [total_energy(d["vel"].numpy(st.ptr), handle.force_read(st.ptr)[0])]
            for k in range(steps):
                handle.nve_kick_drift(d["pos"], d["vel"], d_f, s.N, dt, st.ptr, image=d["image"] if wrap else None,
                                      box=s.box if wrap else None)
                handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
                handle.nve_half_kick(d["vel"], d_f, s.N, dt, st.ptr)
                if (k + 1) % every == 0:
                    E_gpu.append(total_energy(d["vel"].numpy(st.ptr), handle.force_read(st.ptr)[0]))
        E_gpu = np.array(E_gpu)
        drift_gpu = np.abs(E_gpu - E_gpu[0]).max() / abs(E_gpu[0])
        print(f"NVE {steps} steps at N={s.N}, {path}: relative drift cpu {drift_cpu:.3e} gpu {drift_gpu:.3e}; "
              f"|E_gpu-E_cpu|max {np.abs(E_gpu - E_cpu).max() / abs(E_cpu[0]):.3e}")
        assert len(E_gpu) == len(E_cpu)
        assert drift_gpu <= 1.02 * drift_cpu + 1e-12
        assert np.abs(E_gpu - E_cpu).max() <= 1e-9 * abs(E_cpu[0])


def test_nve_drift_one_launch_per_step(handle, coracle):
    """The same 10k-step NVE comparison through the most fused path: cavb200_md_step_fused with no thermostat
    (alpha = 1), i.e. ONE persistent launch per MD step with the rank-1 cavity force never stored."""
    n_mol, steps, every, dt = 2000, 10000, 500, 5.0
    omegac, g, phmass = 0.01, 1e-3, 1.0
    s = synth.make_system(n_mol, images=False)
    pos, vel, force = s.pos.copy(), s.vel.copy(), np.zeros((s.N, 4))
    e = coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, omegac, g, phmass)
    force[:] = e["force"]
    E_cpu = [total_energy(vel, e["energies"])]
    for k in range(steps):
        en = coracle.nve_step(pos, vel, s.charge, s.image, force, s.box, s.L_typeid, omegac, g, phmass, dt)
        if (k + 1) % every == 0:
            E_cpu.append(total_energy(vel, en))
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    p = capi.Params.make(omegac, g, phmass)
    st = capi.Stream()
    handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
    E_gpu = [total_energy(d["vel"].numpy(st.ptr), handle.force_read(st.ptr)[0])]
    handle.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, 0, 0, None, st.ptr)
    for k in range(steps):
        if (k + 1) % every == 0 or k == steps - 1:
            # close the step (second half kick) on a copy of the velocities to read the energy at a whole step
            v_copy = capi.DeviceArray.from_numpy(d["vel"].numpy(st.ptr))
            handle.nvt_step_two_rank1(v_copy, None, d["charge"], d["pos"], s.N, dt, s.L_typeid, g, 0, 0, st.ptr)
            E_gpu.append(total_energy(v_copy.numpy(st.ptr), handle.force_read(st.ptr)[0]))
            v_last = v_copy
        if k < steps - 1:
            handle.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, 0, 0, None,
                                 st.ptr)
    E_cpu, E_gpu = np.array(E_cpu), np.array(E_gpu)
    drift_cpu = np.abs(E_cpu - E_cpu[0]).max() / abs(E_cpu[0])
    drift_gpu = np.abs(E_gpu - E_gpu[0]).max() / abs(E_gpu[0])
    print(f"NVE 10k steps, one launch per step: relative drift cpu {drift_cpu:.3e} gpu {drift_gpu:.3e}")
    assert len(E_gpu) == len(E_cpu)
    assert drift_gpu <= 1.02 * drift_cpu + 1e-12
    assert np.abs(E_gpu - E_cpu).max() <= 1e-9 * abs(E_cpu[0])
    assert np.allclose(d["pos"].numpy(st.ptr)[:, :3], pos[:, :3], rtol=1e-8, atol=1e-8)
    assert np.allclose(v_last.numpy(st.ptr)[:, :3], vel[:, :3], rtol=1e-7, atol=1e-12)
