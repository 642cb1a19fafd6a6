"""GPU parity of the fused Bussi kinetic-energy reduce + rescale, and of the fused force+Bussi step.

Oracle: orc_bussi_step (restates reference src/BussiReservoirThermostat.h:43-98,177-225 with the
random draws injected; KE and the v*=alpha loop are stand-ins for HOOMD upstream code).
Tolerances: KE is a sum of N positive terms evaluated in a different order (tree vs serial):
|dKE|/KE <= 1e-12.  alpha given KE is evaluated with the reference's operation order: compared at
1e-12 relative (it inherits KE's difference).  Velocities: 1e-12 relative.
"""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, synth

pytestmark = pytest.mark.gpu

KT, TAU, DT = synth.KT_100K, synth.TAU_5PS, synth.DT_1FS


def bargs(dof, r_normal=0.3, gamma_draw=None, kT=KT, tau=TAU, dt=DT):
    if gamma_draw is None:
        gamma_draw = (dof - 1.0) / 2.0 * 1.001  # a plausible Gamma((dof-1)/2, 1) variate
    return capi.BussiArgs(kT, tau, dt, dof, r_normal, gamma_draw)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n_mol", [1, 2, 33, 1000, 65537, 300001])
def test_bussi_contiguous_group(handle, coracle, n_mol, variant):
    handle.set_tuning(variant=variant, threads=512, ctas_per_sm=2, unroll=2)
    s = synth.make_system(n_mol)
    n = n_mol  # molecular group = [0, n_mol); the photon (last) is not thermostatted
    dof = 3.0 * n - 3.0 if n > 1 else 3.0
    a = bargs(dof)
    vel_ref = s.vel.copy()
    res = np.zeros(2)
    alpha_ref, ke_ref = coracle.bussi_step(vel_ref, None if n == s.N else np.arange(n, dtype=np.uint32), dof, DT, KT,
                                           TAU, a.r_normal, a.gamma_draw, res)
    d_vel = capi.DeviceArray.from_numpy(s.vel)
    handle.bussi_reset()
    handle.bussi(d_vel, None, 0, n, a)
    out = handle.bussi_read()
    assert out["err"] == 0.0
    assert abs(out["ke"] - ke_ref) <= 1e-12 * ke_ref
    assert abs(out["alpha"] - alpha_ref) <= 1e-12 * abs(alpha_ref)
    assert abs(out["instantaneous"] - res[1]) <= 1e-9 * max(abs(res[1]), 1e-12 * ke_ref)
    v = d_vel.numpy()
    assert np.array_equal(v[:, 3], s.vel[:, 3])  # masses untouched
    assert np.array_equal(v[n:], s.vel[n:])  # photon untouched
    assert np.allclose(v[:n, :3], vel_ref[:n, :3], rtol=1e-12, atol=0)


def test_bussi_alpha_bit_exact_given_ke(handle, coracle):
    """With a power-of-two number of equal terms the KE sum is exact in any order, so alpha and the
    reservoir delta must match the oracle bit for bit (same operation order, no FMA contraction)."""
    n = 4096
    vel = np.zeros((n, 4))
    vel[:, 0], vel[:, 1], vel[:, 2], vel[:, 3] = 0.5, -0.25, 0.125, 8.0
    for r_normal, tau in [(0.3, TAU), (-2.5, TAU), (1.1, 0.0), (-0.7, 50.0)]:
        a = bargs(3.0 * n, r_normal=r_normal, tau=tau)
        vref = vel.copy()
        res = np.zeros(2)
        alpha_ref, ke_ref = coracle.bussi_step(vref, None, a.dof, DT, KT, tau, a.r_normal, a.gamma_draw, res)
        d_vel = capi.DeviceArray.from_numpy(vel)
        handle.bussi_reset()
        handle.bussi(d_vel, None, 0, n, a)
        out = handle.bussi_read()
        assert out["ke"] == ke_ref
        assert out["alpha"] == alpha_ref
        assert out["instantaneous"] == res[1] and out["cumulative"] == res[0]
        assert np.array_equal(d_vel.numpy(), vref)


def test_bussi_index_list_group_and_cumulative(handle, coracle):
    s = synth.make_system(20000, photon="middle")
    idx = synth.molecular_group(s)
    assert len(idx) == 20000 and not np.array_equal(idx, np.arange(20000))
    d_idx = capi.DeviceArray.from_numpy(idx)
    d_vel = capi.DeviceArray.from_numpy(s.vel)
    vref = s.vel.copy()
    res = np.zeros(2)
    handle.bussi_reset()
    rng = np.random.default_rng(5)
    dof = 3.0 * len(idx) - 3
    for step in range(5):
        a = bargs(dof, r_normal=rng.standard_normal(), gamma_draw=rng.gamma((dof - 1) / 2))
        alpha_ref, ke_ref = coracle.bussi_step(vref, idx, dof, DT, KT, TAU, a.r_normal, a.gamma_draw, res)
        handle.bussi(d_vel, d_idx, 0, len(idx), a)
        out = handle.bussi_read()
        assert abs(out["alpha"] - alpha_ref) <= 1e-12 * abs(alpha_ref)
        assert abs(out["cumulative"] - res[0]) <= 1e-9 * max(abs(res[0]), 1e-10 * ke_ref)
    assert np.allclose(d_vel.numpy(), vref, rtol=1e-11, atol=0)
    handle.bussi_reset()
    out = handle.bussi_read()
    assert out["cumulative"] == 0.0 and out["instantaneous"] == 0.0  # reference test :59-61,74-76


def test_bussi_edge_cases(handle):
    s = synth.make_system(1000)
    d_vel = capi.DeviceArray.from_numpy(s.vel)
    # dt == 0: no-op, alpha = 1 (BussiReservoirThermostat.h:45-48)
    handle.bussi_reset()
    handle.bussi(d_vel, None, 0, 1000, bargs(2997.0, dt=0.0))
    assert np.array_equal(d_vel.numpy(), s.vel)
    # dof == 0: alpha = 1 without touching velocities (:183-184)
    handle.bussi(d_vel, None, 0, 1000, bargs(0.0, gamma_draw=0.0))
    out = handle.bussi_read()
    assert out["alpha"] == 1.0 and out["instantaneous"] == 0.0
    assert np.array_equal(d_vel.numpy(), s.vel)
    # zero kinetic energy with dof != 0: the reference throws (:57-61) -> error flag, no rescale
    z = s.vel.copy()
    z[:, :3] = 0.0
    d_z = capi.DeviceArray.from_numpy(z)
    handle.bussi(d_z, None, 0, 1000, bargs(2997.0))
    assert handle.bussi_read()["err"] == 1.0
    handle.bussi_reset()
    # KE-only entry point
    handle.bussi_ke(d_vel, None, 0, 1000)
    ke = 0.5 * np.sum(s.vel[:1000, 3] * np.sum(s.vel[:1000, :3] ** 2, axis=1))
    assert abs(handle.bussi_read()["ke"] - ke) <= 1e-12 * ke
    # negative alpha (sign rule, eq. A8): strongly negative normal draw with tau -> 0 limit
    a = bargs(2997.0, r_normal=-80.0, gamma_draw=1490.0, tau=1e-3)
    handle.bussi(d_vel, None, 0, 1000, a)
    assert handle.bussi_read()["alpha"] < 0


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n_mol", [1000, 262145])
def test_fused_step_equals_separate_calls(handle, coracle, n_mol, variant):
    """cavb200_step == cavb200_force followed by cavb200_bussi, and both == oracle."""
    handle.set_tuning(variant=variant, threads=512, ctas_per_sm=2, unroll=2)
    s = synth.make_system(n_mol)
    p = capi.Params.make(0.01, 1e-3)
    a = bargs(3.0 * n_mol - 3.0)
    dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
    handle.bussi_reset()
    handle.step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, s.box, s.L_typeid, p, 0, n_mol, a)
    en, dip, ph = handle.force_read()
    bo = handle.bussi_read()
    f_fused, v_fused = d_f.numpy(), dev["vel"].numpy()

    d_vel2 = capi.DeviceArray.from_numpy(s.vel)
    d_f2 = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
    handle.bussi_reset()
    handle.force(dev["pos"], dev["charge"], dev["image"], d_f2, s.N, s.box, s.L_typeid, p)
    handle.bussi(d_vel2, None, 0, n_mol, a)
    en2, dip2, ph2 = handle.force_read()
    bo2 = handle.bussi_read()
    assert np.array_equal(f_fused.view(np.uint64), d_f2.numpy().view(np.uint64))
    assert np.array_equal(v_fused.view(np.uint64), d_vel2.numpy().view(np.uint64))
    assert np.array_equal(en, en2) and bo["alpha"] == bo2["alpha"]

    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert np.abs(f_fused - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
    vref = s.vel.copy()
    alpha_ref, _ = coracle.bussi_step(vref, np.arange(n_mol, dtype=np.uint32), a.dof, DT, KT, TAU, a.r_normal,
                                      a.gamma_draw, np.zeros(2))
    assert np.allclose(v_fused, vref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("n_mol", [500, 3000])
def test_small_system_kernels_under_cuda_graph(coracle, n_mol):
    """The single-CTA kernel (500 particles) and the single-cluster kernel (3000 particles: cluster launch attribute, 16
    CTAs) captured in a CUDA graph as the FIRST thing a fresh handle does -- nothing the launcher needs may be set up
    lazily inside the capture -- and replayed: force call + Bussi call, three steps."""
    h = capi.Handle(0)
    try:
        s = synth.make_system(n_mol)
        p = capi.Params.make(0.01, 1e-3)
        a = bargs(3.0 * n_mol - 3.0)
        dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
        st = capi.Stream()

        def two_calls():
            h.force(dev["pos"], dev["charge"], dev["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
            h.bussi(dev["vel"], None, 0, n_mol, a, st.ptr)

        g = h.graph_capture(st.ptr, two_calls)
        vref = s.vel.copy()
        res = np.zeros(2)
        for _ in range(3):
            h.graph_launch(g, st.ptr)
            coracle.bussi_step(vref, np.arange(n_mol, dtype=np.uint32), a.dof, DT, KT, TAU, a.r_normal, a.gamma_draw, res)
        st.sync()
        h.graph_destroy(g)
        assert np.allclose(dev["vel"].numpy(), vref, rtol=1e-11, atol=0)
        ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
        assert np.abs(d_f.numpy() - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
        assert h.fault_count == 0
    finally:
        h.close()


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_step_under_cuda_graph(handle, coracle, variant):
    """The step captured in a CUDA graph and replayed gives the same result as direct launches
    (the hand-off keeps no host-side state)."""
    s = synth.make_system(50000)
    p = capi.Params.make(0.01, 1e-3)
    a = bargs(3.0 * 50000 - 3.0)
    dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    st = capi.Stream()
    handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
    handle.bussi_reset(st.ptr)
    g = handle.graph_capture(st.ptr, lambda: handle.step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N,
                                                         s.box, s.L_typeid, p, 0, 50000, a, st.ptr))
    vref = s.vel.copy()
    res = np.zeros(2)
    for _ in range(3):
        handle.graph_launch(g, st.ptr)
        coracle.bussi_step(vref, np.arange(50000, dtype=np.uint32), a.dof, DT, KT, TAU, a.r_normal, a.gamma_draw, res)
    st.sync()
    handle.graph_destroy(g)
    assert np.allclose(dev["vel"].numpy(), vref, rtol=1e-11, atol=0)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert np.abs(d_f.numpy() - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
    assert abs(handle.bussi_read(st.ptr)["cumulative"] - res[0]) <= 1e-9 * abs(res[0])


def test_step_host_buffers(handle, coracle):
    """Host-buffer entry point (the e2e path): same results as the device path, pinned or pageable."""
    n_mol = 100000
    s = synth.make_system(n_mol)
    p = capi.Params.make(0.01, 1e-3)
    a = bargs(3.0 * n_mol - 3.0)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    vref = s.vel.copy()
    coracle.bussi_step(vref, np.arange(n_mol, dtype=np.uint32), a.dof, DT, KT, TAU, a.r_normal, a.gamma_draw, np.zeros(2))
    for pinned in (True, False):
        if pinned:
            bufs = {k: capi.PinnedArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
            f = capi.PinnedArray((s.N, 4), np.float64)
            args = [bufs["pos"], bufs["charge"], bufs["image"], f, bufs["vel"]]
            fa, va = f.array, bufs["vel"].array
        else:
            va = s.vel.copy()
            fa = np.full((s.N, 4), np.nan)
            args = [s.pos, s.charge, s.image, fa, va]
        handle.bussi_reset()
        en, bo = handle.step_host(*args, s.N, s.box, s.L_typeid, p, 0, n_mol, a)
        assert np.abs(fa - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
        assert np.allclose(va, vref, rtol=1e-12, atol=0)
        assert np.allclose(en, ref["energies"], rtol=1e-10)
        assert bo["err"] == 0.0


def test_step_host_pipelined_slots(handle, coracle):
    """cavb200_step_host_submit / _wait: several independent host-resident systems of DIFFERENT sizes in flight
    at once (next upload under this download); every slot returns its own system's forces, velocities,
    energies and alpha (the per-handle Scalars block is snapshotted per slot), and slots are reusable."""
    p = capi.Params.make(0.01, 1e-3)
    sizes = [60000, 1000, 33, 150000, 60001, 7, 20000]
    systems, refs = [], []
    for r, n_mol in enumerate(sizes):
        s = synth.make_system(n_mol, replica=r)
        a = bargs(max(3.0 * n_mol - 3.0, 1.0))
        ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
        vref = s.vel.copy()
        alpha, ke = coracle.bussi_step(vref, np.arange(n_mol, dtype=np.uint32), a.dof, DT, KT, TAU, a.r_normal,
                                       a.gamma_draw, np.zeros(2))
        bufs = {k: capi.PinnedArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        bufs["force"] = capi.PinnedArray((s.N, 4), np.float64)
        bufs["force"].array[:] = np.nan
        systems.append((s, a, bufs))
        refs.append((ref, vref, alpha, ke))
    handle.bussi_reset()
    SLOTS = 3

    def submit(k):
        s, a, b = systems[k]
        handle.step_host_submit(k % SLOTS, b["pos"], b["charge"], b["image"], b["force"], b["vel"], s.N, s.box, s.L_typeid,
                                p, 0, s.N - 1, a)

    for k in range(SLOTS - 1):
        submit(k)
    for k in range(len(sizes)):
        if k + SLOTS - 1 < len(sizes):
            submit(k + SLOTS - 1)
        en, bo = handle.step_host_wait(k % SLOTS)
        s, a, b = systems[k]
        ref, vref, alpha, ke = refs[k]
        assert np.abs(b["force"].array - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
        assert np.allclose(b["vel"].array, vref, rtol=1e-12, atol=0)
        assert np.allclose(en, ref["energies"], rtol=1e-10)
        assert abs(bo["alpha"] - alpha) <= 1e-12 * abs(alpha) and abs(bo["ke"] - ke) <= 1e-12 * ke
        assert bo["err"] == 0.0
    with pytest.raises(capi.CavbError):
        handle.step_host_wait(99)


def test_step_host_reduced_copies(handle, coracle):
    """cavb200_step_host_submit_ex: charges / images flagged "unchanged since the slot's last submit" are not uploaded
    (the host arrays may then hold anything -- here NaN / garbage), and the rank-1 result {Dq, F_L, photon index} stands
    in for the force array: F_i = (-g c_i) Dq bit for bit (reference src/CavityForceCompute.cc:183,188-200)."""
    n_mol = 80000
    p = capi.Params.make(0.01, 1e-3)
    s = synth.make_system(n_mol, replica=2)
    a = bargs(3.0 * n_mol - 3.0)
    bufs = {k: capi.PinnedArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    force = capi.PinnedArray((s.N, 4), np.float64)
    args = (s.N, s.box, s.L_typeid, p, 0, n_mol, a)
    # KEEP_* without an earlier submit of this size on the slot is refused
    with pytest.raises(capi.CavbError):
        handle.step_host_submit_ex(3, bufs["pos"], None, bufs["image"], force, bufs["vel"], *args, handle.HOST_KEEP_CHARGE)
    handle.bussi_reset()
    handle.step_host_submit_ex(3, bufs["pos"], bufs["charge"], bufs["image"], force, bufs["vel"], *args, 0)
    handle.step_host_wait_ex(3)
    # second step: new positions / velocities, charges and images not sent (and poisoned on the host)
    rng = np.random.default_rng(5)
    s2 = s.copy()
    s2.pos[:, :3] += 0.01 * rng.standard_normal((s.N, 3))
    bufs["pos"].array[...] = s2.pos
    s2.vel[...] = bufs["vel"].array  # rescaled by the first step
    bufs["charge"].array[...] = np.nan
    bufs["image"].array[...] = 12345
    ref = coracle.cavity_force(s2.pos, s2.charge, s2.image, s2.box, s2.L_typeid, 0.01, 1e-3)
    vref = s2.vel.copy()
    alpha, ke = coracle.bussi_step(vref, np.arange(n_mol, dtype=np.uint32), a.dof, DT, KT, TAU, a.r_normal, a.gamma_draw,
                                   np.zeros(2))
    force.array[...] = np.nan
    handle.step_host_submit_ex(3, bufs["pos"], None, None, force, bufs["vel"], *args,
                               handle.HOST_KEEP_CHARGE | handle.HOST_KEEP_IMAGE)
    en, bo, r1 = handle.step_host_wait_ex(3)
    scale = np.abs(ref["force"]).max()
    assert np.abs(force.array - ref["force"]).max() <= 1e-10 * scale
    assert np.allclose(bufs["vel"].array, vref, rtol=1e-12, atol=0) and np.allclose(en, ref["energies"], rtol=1e-10)
    assert abs(bo["alpha"] - alpha) <= 1e-12 * abs(alpha) and r1["photon_idx"] == ref["photon_idx"]
    full = force.array.copy()
    # third call: the same inputs (velocities restored), rank-1 result instead of the force array
    bufs["vel"].array[...] = s2.vel
    handle.step_host_submit_ex(3, bufs["pos"], None, None, None, bufs["vel"], *args,
                               handle.HOST_KEEP_CHARGE | handle.HOST_KEEP_IMAGE | handle.HOST_RANK1_RESULT)
    en3, bo3, r3 = handle.step_host_wait_ex(3)
    assert np.array_equal(en3, en) and np.array_equal(r3["Dq"], r1["Dq"]) and np.array_equal(r3["F_L"], r1["F_L"])
    f1 = np.zeros((s.N, 4))
    f1[:, 0] = (-1e-3 * s2.charge) * r3["Dq"][0]
    f1[:, 1] = (-1e-3 * s2.charge) * r3["Dq"][1]
    f1[r3["photon_idx"], :3] = r3["F_L"]
    assert np.array_equal(f1, full)  # the stored force IS the rank-1 product
    assert np.allclose(bufs["vel"].array, vref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("n,first,use_list", [(1, 0, False), (33, 0, False), (1000, 7, False), (262145, 0, False),
                                               (50001, 0, True), (0, 0, False)])
def test_bussi_ke_only(handle, n, first, use_list, variant):
    """cavb200_bussi_ke = the kinetic energy getRescalingFactorsOne reads from ComputeThermo
    (reference src/BussiReservoirThermostat.h:50-55): 1/2 sum m|v|^2 over the group, velocities untouched.
    Tolerance 1e-12 relative (tree sum vs NumPy pairwise sum of positive terms)."""
    handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
    s = synth.make_system(max(n + first, 300000 if use_list else 1))
    d_vel = capi.DeviceArray.from_numpy(s.vel)
    if use_list:
        idx = np.arange(1, 2 * n, 2, dtype=np.uint32)
        d_idx = capi.DeviceArray.from_numpy(idx)
        handle.bussi_ke(d_vel, d_idx, 0, n)
        sel = s.vel[idx]
    else:
        handle.bussi_ke(d_vel, None, first, n)
        sel = s.vel[first:first + n]
    ke = 0.5 * np.sum(sel[:, 3] * np.sum(sel[:, :3] ** 2, axis=1))
    got = handle.bussi_read()["ke"]
    assert abs(got - ke) <= 1e-12 * max(ke, 1e-300)
    assert np.array_equal(d_vel.numpy().view(np.uint64), s.vel.view(np.uint64))
    handle.set_tuning(variant=3)
