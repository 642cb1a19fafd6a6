"""GPU parity of the cavity force through the C ABI against the CPU oracle.

Oracle: oracle/liboracle.so (bit-exact restatement of reference src/CavityForceCompute.cc:73-208,
itself pinned against the verbatim-compiled reference in tests/test_oracle.py).
Tolerance (BASELINE.json north_star): forces and cavity energies within 1e-10 relative, fp64.
Reduction-order difference, stated: the reference sums c_i*u_i index-ascending in plain doubles,
the kernel sums the same (bit-identical) terms with compensated pairs in a fixed tree; both are
compared with the exact sum of the terms (orc_dipole_exact) and the kernel must be the closer.
"""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-10
OMEGAC, G = 0.01, 1e-3

GRID_N = [1, 2, 31, 32, 33, 255, 256, 257, 1000, 65537, 100001, 262145]
PHOTON = ["last", "first", "middle", "absent", "duplicated"]


def run_gpu(handle, s, omegac=OMEGAC, g=G, phmass=1.0, L_typeid=None, prefill=np.nan):
    L_typeid = s.L_typeid if L_typeid is None else L_typeid
    d_pos = capi.DeviceArray.from_numpy(s.pos)
    d_q = capi.DeviceArray.from_numpy(s.charge)
    d_img = capi.DeviceArray.from_numpy(s.image)
    d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), prefill))  # every entry must be overwritten
    handle.force(d_pos, d_q, d_img, d_f, s.N, s.box, L_typeid, capi.Params.make(omegac, g, phmass))
    en, dip, ph = handle.force_read()
    return dict(force=d_f.numpy(), energies=en, dipole=dip, photon_idx=ph)


def assert_parity(gpu, ref, exact_dipole=None):
    assert gpu["photon_idx"] == ref["photon_idx"]
    assert not np.isnan(gpu["force"]).any(), "some force entries were not written"
    fr, fg = ref["force"], gpu["force"]
    scale = np.abs(fr).max()
    if scale == 0:
        assert np.all(fg == 0.0)
    else:
        assert np.abs(fg - fr).max() <= RTOL * scale
        nz = np.abs(fr) > 1e-6 * scale
        assert np.all(np.abs(fg[nz] - fr[nz]) <= RTOL * np.abs(fr[nz]))
        # z and w components are exact zeros for molecular particles
        assert np.all(fg[:, 3] == 0.0)
    for a, b in zip(gpu["energies"], ref["energies"]):
        assert abs(a - b) <= RTOL * max(abs(b), 1e-300)
    if exact_dipole is not None and ref["photon_idx"] >= 0:
        err_gpu = np.abs(gpu["dipole"] - exact_dipole)
        err_ref = np.abs(ref["dipole"] - exact_dipole)
        ulp = np.spacing(np.abs(exact_dipole))
        assert np.all(err_gpu <= 2 * ulp), (err_gpu, ulp)
        assert np.all(err_gpu <= err_ref + 2 * ulp)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n_mol", GRID_N)
def test_force_grid(handle, coracle, n_mol, variant):
    handle.set_tuning(variant=variant, threads=512, ctas_per_sm=2, unroll=2)
    s = synth.make_system(n_mol, replica=n_mol % 7)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G)
    exact = coracle.dipole_exact(s.pos, s.charge, s.image, s.box, ref["photon_idx"])
    assert_parity(run_gpu(handle, s), ref, exact)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("photon", PHOTON)
@pytest.mark.parametrize("n_mol", [1, 33, 1000, 100001])
def test_force_photon_placement(handle, coracle, n_mol, photon, variant):
    handle.set_tuning(variant=variant, threads=256, ctas_per_sm=1, unroll=8)
    s = synth.make_system(n_mol, photon=photon)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G)
    exact = coracle.dipole_exact(s.pos, s.charge, s.image, s.box, ref["photon_idx"])
    assert_parity(run_gpu(handle, s), ref, exact)


@pytest.mark.parametrize("charges,images", [("zero", True), ("nonneutral", True), ("neutral", False)])
def test_force_charge_and_image_cases(handle, coracle, charges, images):
    handle.set_tuning(variant=1, threads=512, ctas_per_sm=2, unroll=4)
    s = synth.make_system(4097, charges=charges, images=images)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G)
    assert_parity(run_gpu(handle, s), ref)


@pytest.mark.parametrize("omegac,g,phmass", [(1e-4, 1e-3, 1.0), (2000 / 219474.63, 1e-3, 1.0), (0.01, -2e-3, 3.0)])
def test_force_parameter_cases(handle, coracle, omegac, g, phmass):
    s = synth.make_system(10000, omegac=omegac, phmass=phmass)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, omegac, g, phmass)
    assert_parity(run_gpu(handle, s, omegac, g, phmass), ref)


def test_force_many_L_particles(handle, coracle):
    """Pathological input: a third of the particles are of type 'L' (only the first is the photon)."""
    s = synth.make_system(5000)
    tid = s.typeid.copy()
    tid[::3] = s.L_typeid
    s.pos[:, 3] = synth.typeid_to_w(tid)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G)
    for variant in (0, 1):
        handle.set_tuning(variant=variant)
        assert_parity(run_gpu(handle, s), ref)


def test_force_no_L_type(handle):
    """No type named 'L' (L_typeid = UINT32_MAX): zero forces and energies
    (reference src/CavityForceComputeGPU.cc:114-123)."""
    s = synth.make_system(1000)
    out = run_gpu(handle, s, L_typeid=0xFFFFFFFF)
    assert np.all(out["force"] == 0.0) and np.all(out["energies"] == 0.0) and out["photon_idx"] == -1


def test_force_error_convention(handle):
    """NULL array -> cudaErrorInvalidValue (reference .cu:522-528); N == 0 -> success (:530-532)."""
    p = capi.Params.make(OMEGAC, G)
    handle.force(None, None, None, None, 0, (1, 1, 1), 2, p)
    with pytest.raises(capi.CavbError) as e:
        handle.force(None, None, None, None, 10, (1, 1, 1), 2, p)
    assert e.value.code == 1
    d = capi.DeviceArray(64, np.float64)
    with pytest.raises(capi.CavbError) as e:  # Scalar4 arrays must be 32-byte aligned
        handle.force(d.ptr + 8, d.ptr, d.ptr, d.ptr, 1, (1, 1, 1), 2, p)
    assert e.value.code != 0


def test_force_deterministic_and_idempotent(handle):
    """Same inputs, same launch shape -> bitwise identical output; inputs are not modified."""
    s = synth.make_system(300000)
    handle.set_tuning(variant=1, threads=512, ctas_per_sm=2, unroll=2)
    a = run_gpu(handle, s)
    b = run_gpu(handle, s)
    assert np.array_equal(a["force"].view(np.uint64), b["force"].view(np.uint64))
    assert np.array_equal(a["dipole"], b["dipole"])


@pytest.mark.parametrize("threads,ctas,unroll,variant", [(256, 4, 2, 1), (1024, 1, 2, 1), (512, 1, 4, 1), (512, 2, 4, 0),
                                                         (128, 8, 2, 0), (256, 1, 8, 1), (128, 2, 8, 0)])
def test_force_launch_shapes(handle, coracle, threads, ctas, unroll, variant):
    handle.set_tuning(variant=variant, threads=threads, ctas_per_sm=ctas, unroll=unroll)
    s = synth.make_system(200003)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G)
    exact = coracle.dipole_exact(s.pos, s.charge, s.image, s.box, ref["photon_idx"])
    assert_parity(run_gpu(handle, s), ref, exact)
    handle.set_tuning(variant=1, threads=512, ctas_per_sm=2, unroll=2)


def test_force_full_size_properties(handle, coracle):
    """BASELINE size (1M + photon): oracle parity plus size-independent identities
    (SURVEY.md 8c): sum_mol F = -g Q_tot Dq, F_L,xy = -K Dq, E_h+E_c+E_d = K/2 (|Dq|^2 + q_z^2)."""
    s = synth.make_system(1_000_000)
    out = run_gpu(handle, s)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, OMEGAC, G)
    assert_parity(out, ref, coracle.dipole_exact(s.pos, s.charge, s.image, s.box, ref["photon_idx"]))
    K = OMEGAC * OMEGAC
    ph = out["photon_idx"]
    u_ph = s.pos[ph, :3] + s.image[ph] * np.asarray(s.box)
    Dq = u_ph[:2] + (G / K) * out["dipole"][:2]
    FL = out["force"][ph]
    assert np.allclose(FL[:2], -K * Dq, rtol=1e-12)
    mol = np.ones(s.N, bool)
    mol[ph] = False
    # rank-1 structure: every molecular force is (-g c_i) Dq
    assert np.allclose(out["force"][mol, :2], (-G * s.charge[mol])[:, None] * Dq[None, :], rtol=1e-12, atol=0)
    etot = out["energies"].sum()
    assert abs(etot - 0.5 * K * (Dq @ Dq + u_ph[2] ** 2)) <= 1e-9 * abs(etot)
