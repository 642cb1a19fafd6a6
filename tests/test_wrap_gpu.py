"""Box wrap + image update of the drift (r <- r + v dt, then wrap): every drift entry point
with the wrap against the oracle's orc_nve_step_w / orc_nvt_step_w (HOOMD's BoxDim::wrap restated, oracle/cavity_oracle.c
orc_wrap), on boxes where MANY particles cross a face -- both directions, all three axes -- within a few steps.

What is compared after the run: image flags exactly, wrapped positions and velocities to 1e-10 of the array's largest
magnitude (the BASELINE tolerance), and that the unwrapped positions pos + image * L -- what the dipole uses, reference
src/CavityForceCompute.cc:107-109 -- agree with a run WITHOUT the wrap."""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, rng, synth

pytestmark = pytest.mark.gpu
# a weak coupling: with the velocities scaled up as below, g = 1e-3 lets the collective cavity force throw the particles
# many box lengths per step (positions ~1e8 at 200k particles), which a once-per-step wrap cannot and need not follow
OMEGAC, G, PHMASS = 0.01, 1e-8, 1.0
KT, TAU = synth.KT_100K, synth.TAU_5PS


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def fast_system(n_mol, replica, photon="last"):
    """The Appendix-D box with velocities scaled up so that a particle travels ~L/6 per step: a large fraction of the
    particles crosses a face within 6 steps (the photon, heavy in q, stays put)."""
    s = synth.make_system(n_mol, replica=replica, photon=photon)
    L = s.box[0]
    mol = s.typeid != s.L_typeid
    vmax = np.abs(s.vel[mol, :3]).max()
    s.vel[mol, :3] *= (L / 6.0) / (vmax * synth.DT_1FS)
    return s


def oracle_run(coracle, s, steps, dt, first, n, dof, draws, wrap, thermostat=True):
    pos, vel, image, force = s.pos.copy(), s.vel.copy(), s.image.copy(), np.zeros((s.N, 4))
    force[:] = coracle.cavity_force(pos, s.charge, image, s.box, s.L_typeid, OMEGAC, G, PHMASS)["force"]
    ke = np.array([coracle.kinetic_energy(vel, np.arange(first, first + n, dtype=np.uint32))])
    res = np.zeros(2)
    alphas = []
    for t in range(steps):
        if thermostat:
            a, en = coracle.nvt_step(pos, vel, s.charge, image, force, s.box, s.L_typeid, OMEGAC, G, PHMASS, dt, first, n, dof,
                                     KT, TAU, draws[t][0], draws[t][1], res, ke, wrap=wrap)
            alphas.append(a)
        else:
            en = coracle.nve_step(pos, vel, s.charge, image, force, s.box, s.L_typeid, OMEGAC, G, PHMASS, dt, wrap=wrap)
    return pos, vel, image, alphas, en


def check(s, got, ref, steps):
    pos, vel, image = got
    rpos, rvel, rimage = ref[:3]
    L = np.array(s.box)
    crossed = np.any(rimage != s.image, axis=1).sum()
    assert crossed > 0.2 * s.N or s.N < 10, f"only {crossed} of {s.N} particles crossed a face: the test does not test"
    assert (rimage > s.image).any() and (rimage < s.image).any() or s.N < 10
    assert np.array_equal(image, rimage)
    assert _rel(pos[:, :3], rpos[:, :3]) <= 1e-10 and _rel(vel[:, :3], rvel[:, :3]) <= 1e-10
    # every molecular particle is back inside the box (the photon coordinate of this synthetic system is driven further
    # than a box length per step by the fast charges: one shift per step, as HOOMD's wrap does, leaves it outside)
    mol = s.typeid != s.L_typeid
    assert np.all(pos[mol, :3] >= -L / 2) and np.all(pos[mol, :3] < L / 2)
    assert np.array_equal(pos[:, 3].view(np.uint64), s.pos[:, 3].view(np.uint64))  # type words untouched


@pytest.mark.parametrize("n_mol", [1, 33, 1000, 50000])
def test_nve_kick_drift_wrap(handle, coracle, n_mol):
    s = fast_system(n_mol, 41)
    steps, dt = 6, synth.DT_1FS
    p = capi.Params.make(OMEGAC, G, PHMASS)
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p)
    for _ in range(steps):
        handle.nve_kick_drift(d["pos"], d["vel"], d_f, s.N, dt, None, image=d["image"], box=s.box)
        handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p)
        handle.nve_half_kick(d["vel"], d_f, s.N, dt)
    ref = oracle_run(coracle, s, steps, dt, 0, 0, 0.0, None, True, thermostat=False)
    check(s, (d["pos"].numpy(), d["vel"].numpy(), d["image"].numpy()), ref, steps)
    # the wrap changes nothing physical: unwrapped positions equal those of the run without it
    nowrap = oracle_run(coracle, s, steps, dt, 0, 0, 0.0, None, False, thermostat=False)
    L = np.array(s.box)
    u_w = ref[0][:, :3] + ref[2] * L
    u_n = nowrap[0][:, :3] + nowrap[2] * L
    assert np.abs(u_w - u_n).max() <= 1e-12 * np.abs(u_n).max()


@pytest.mark.parametrize("path", ["stored", "rank1", "md_one", "md_fused"])
@pytest.mark.parametrize("n_mol,photon", [(1, "last"), (300, "middle"), (20000, "last"), (200000, "last")])
def test_thermostatted_steps_with_wrap(handle, coracle, n_mol, photon, path):
    """cavb200_nvt_step_one_wrap (stored force), _rank1_wrap, cavb200_md_step_one_wrap and cavb200_md_step_fused_wrap,
    each against orc_nvt_step_w with the same draws."""
    s = fast_system(n_mol, 43, photon)
    steps, dt = 6, synth.DT_1FS
    first, n = (0, n_mol) if photon == "last" else (0, s.N)
    dof = max(3.0 * n - 3.0, 1.0)
    draws = [rng.bussi_draws(t, 3, 0, dof) for t in range(steps)]
    p = capi.Params.make(OMEGAC, G, PHMASS)
    st = capi.Stream()
    d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    handle.bussi_reset(st.ptr)
    if path == "stored":
        handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
    else:
        handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
    handle.bussi_ke(d["vel"], None, first, n, st.ptr)
    alphas = []
    for t in range(steps):
        a = capi.BussiArgs(KT, TAU, dt, dof, *draws[t])
        if path == "stored":
            handle.nvt_step_one(d["pos"], d["vel"], d_f, s.N, dt, first, n, a, st.ptr, image=d["image"], box=s.box)
            handle.force(d["pos"], d["charge"], d["image"], d_f, s.N, s.box, s.L_typeid, p, st.ptr)
            handle.nvt_step_two(d["vel"], d_f, s.N, dt, first, n, st.ptr)
        elif path == "rank1":
            handle.nvt_step_one_rank1(d["pos"], d["vel"], None, d["charge"], s.N, dt, s.L_typeid, G, first, n, a, st.ptr,
                                      image=d["image"], box=s.box)
            handle.force_rank1(d["pos"], d["charge"], d["image"], s.N, s.box, s.L_typeid, p, st.ptr)
            handle.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, first, n, st.ptr)
        elif path == "md_one" or t == 0:
            handle.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n, a,
                               st.ptr, wrap=True)
            if path == "md_one" or steps == 1:
                handle.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, first, n, st.ptr)
        else:
            handle.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], s.N, dt, s.box, s.L_typeid, p, first, n, a,
                                 st.ptr, wrap=True)
            if t == steps - 1:
                handle.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], s.N, dt, s.L_typeid, G, first, n, st.ptr)
        alphas.append(handle.bussi_read(st.ptr)["alpha"])
    ref = oracle_run(coracle, s, steps, dt, first, n, dof, draws, True)
    check(s, (d["pos"].numpy(st.ptr), d["vel"].numpy(st.ptr), d["image"].numpy(st.ptr)), ref, steps)
    assert np.allclose(alphas, ref[3], rtol=1e-11, atol=0)
    assert np.allclose(handle.force_read(st.ptr)[0], ref[4], rtol=1e-10, atol=1e-300)
    assert handle.bussi_read(st.ptr)["err"] == 0.0


def test_wrap_face_conventions(handle, coracle):
    """Exactly on the upper face -> wrapped down (>= hi); exactly on the lower face -> stays (< lo is strict); one
    shift per step even when the particle is more than a box length out."""
    L = 10.0
    pos = np.zeros((4, 4))
    vel = np.zeros((4, 4))
    vel[:, 3] = 1.0
    pos[0, 0] = 4.0      # moves to exactly +5.0 = hi  -> -5.0, image +1
    vel[0, 0] = 1.0
    pos[1, 1] = -4.0     # moves to exactly -5.0 = lo  -> stays, image unchanged
    vel[1, 1] = -1.0
    pos[2, 2] = 4.0      # moves to 4 + 27 = 31 -> ONE shift: 21, image +1 (HOOMD wraps once per call)
    vel[2, 2] = 27.0
    pos[3, :3] = (-4.5, 4.5, -4.5)
    vel[3, :3] = (-1.0, 1.0, -0.25)
    image = np.array([[0, 0, 0], [2, -3, 1], [0, 0, 7], [1, 1, 1]], dtype=np.int32)
    force = np.zeros((4, 4))
    d_p, d_v, d_i, d_f = (capi.DeviceArray.from_numpy(x) for x in (pos, vel, image, force))
    handle.nve_kick_drift(d_p, d_v, d_f, 4, 1.0, None, image=d_i, box=(L, L, L))
    p, im = d_p.numpy(), d_i.numpy()
    rp, ri = pos.copy(), image.copy()
    rp[:, :3] += vel[:, :3]
    coracle.wrap(rp, ri, (L, L, L))
    assert np.array_equal(p, rp) and np.array_equal(im, ri)
    assert p[0, 0] == -5.0 and im[0, 0] == 1 and p[1, 1] == -5.0 and im[1, 1] == -3 and p[2, 2] == 21.0 and im[2, 2] == 8
    assert np.array_equal(im[3], [0, 2, 1]) and np.array_equal(p[3, :3], [4.5, -4.5, -4.75])


def test_wrap_argument_errors(handle):
    a = capi.DeviceArray((8, 4), np.float64)
    img = capi.DeviceArray((8, 3), np.int32)
    with pytest.raises(capi.CavbError):
        handle.nve_kick_drift(a, a, a, 8, 1.0, None, image=img, box=(1.0, 0.0, 1.0))  # a box length must be positive
    # no image array: cudaErrorInvalidValue, nothing launched
    assert handle.lib.cavb200_nve_kick_drift_wrap(handle.h, a.ptr, a.ptr, a.ptr, None, 8, 1.0, 1.0, 1.0, 1.0, None) == 1
