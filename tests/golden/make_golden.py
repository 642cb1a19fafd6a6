#!/usr/bin/env python
"""tests/golden/make_golden.py -- mint the golden vectors from the REFERENCE ITSELF.

Run in the build container (needs /root/reference): the outputs come from oracle/_ref/libcavref.so,
i.e. the reference's own src/CavityForceCompute.cc and src/BussiReservoirThermostat.h compiled
verbatim by path (oracle/Makefile, oracle/ref_driver.cc), and for F(k,t) from the reference's own
Python, src/cavitymd/analysis.py:34-66,359-364, imported by path under a stub `hoomd` (oracle/refpy.py).
The reference's own tests hold no golden vectors for this path, so these are
the pin.  Inputs are stored next to the outputs (small cases) so the fixtures do not depend on
NumPy's random streams.

    python tests/golden/make_golden.py   ->  tests/golden/cavity_force.npz, bussi.npz, fkt.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cav_hoomd_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import refpy  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

FORCE_CASES = [
    # (name, n_mol, photon, charges, images, omegac, g, phmass)
    ("n1_last", 1, "last", "neutral", True, 0.01, 1e-3, 1.0),
    ("n2_first", 2, "first", "neutral", True, 0.01, 1e-3, 1.0),
    ("n33_middle", 33, "middle", "neutral", True, 0.01, 1e-3, 1.0),
    ("n257_last", 257, "last", "neutral", True, 0.01, 1e-3, 1.0),
    ("n1000_last_cfg1", 1000, "last", "neutral", True, 2000 / 219474.63, 1e-3, 1.0),
    ("n1000_absent", 1000, "absent", "neutral", True, 0.01, 1e-3, 1.0),
    ("n1000_duplicated", 1000, "duplicated", "neutral", True, 0.01, 1e-3, 1.0),
    ("n500_nonneutral_noimg", 500, "last", "nonneutral", False, 0.01, -2e-3, 3.0),
    ("n300_zerocharge", 300, "middle", "zero", True, 1e-4, 1e-3, 1.0),
]


def main():
    ref = O.RefOracle()
    out = {}
    for name, n_mol, photon, charges, images, omegac, g, phmass in FORCE_CASES:
        s = synth.make_system(n_mol, replica=11, photon=photon, charges=charges, images=images, omegac=omegac,
                              phmass=phmass)
        r = ref.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, omegac, g, phmass)
        out[f"{name}/pos"] = s.pos
        out[f"{name}/charge"] = s.charge
        out[f"{name}/image"] = s.image
        out[f"{name}/box"] = np.array(s.box)
        out[f"{name}/params"] = np.array([omegac, g, phmass, s.L_typeid])
        out[f"{name}/force"] = r["force"]
        out[f"{name}/energies"] = r["energies"]
    np.savez_compressed(os.path.join(HERE, "cavity_force.npz"), **out)

    # Bussi: the reference's getRescalingFactorsOne with injected draws, 4 consecutive steps per case
    rng = np.random.default_rng(2026)
    bout = {}
    cases = [("n50", 50, 5.0), ("n100_tau0", 100, 0.0), ("n1000", 1000, synth.TAU_5PS), ("n64_dof1", 64, 7.5)]
    for name, n, tau in cases:
        s = synth.make_system(n, replica=5)
        idx = np.arange(n, dtype=np.uint32)
        dof = 1.0 if name.endswith("dof1") else 3.0 * n - 3.0
        kT, dt = synth.KT_100K, synth.DT_1FS
        h = ref.bussi_open(s.vel, idx, dof, kT, tau)
        draws = np.column_stack([rng.standard_normal(4), rng.gamma(max((dof - 1) / 2, 0.5), size=4)])
        draws[1, 0] = -3.2  # a strongly negative normal draw (sign rule)
        rows = []
        for k in range(4):
            r = ref.bussi_step(h, k, dt, draws[k, 0], draws[k, 1])
            rows.append([r["alpha"], r["ke"], r["cumulative"], r["instantaneous"]])
        bout[f"{name}/vel0"] = s.vel
        bout[f"{name}/args"] = np.array([dof, kT, tau, dt])
        bout[f"{name}/draws"] = draws
        bout[f"{name}/rows"] = np.array(rows)
        bout[f"{name}/vel_final"] = ref.bussi_read(h, s.N)
        ref.bussi_close(h)
    # rotational degrees of freedom (reference src/BussiReservoirThermostat.h:53-55,77-81,87-95): injected rotational
    # dof / kinetic energy, four draws in the reference's order; three consecutive steps per case
    rot_cases = [("rot_n200", 200, synth.TAU_5PS, 3.0 * 200, 0.04), ("rot_n64_rdof1", 64, 12.0, 1.0, 3e-4),
                 ("rot_n128_tau0", 128, 0.0, 2.0 * 128, 0.02)]
    for name, n, tau, rdof, rke in rot_cases:
        s = synth.make_system(n, replica=6)
        dof = 3.0 * n - 3.0
        kT, dt = synth.KT_100K, synth.DT_1FS
        h = ref.bussi_open(s.vel, np.arange(n, dtype=np.uint32), dof, kT, tau)
        draws = np.column_stack([rng.standard_normal(3), rng.gamma((dof - 1) / 2, size=3), rng.standard_normal(3),
                                 rng.gamma(max((rdof - 1) / 2, 0.5), size=3)])
        draws[2, 2] = -2.9  # strongly negative rotational normal draw
        rows = []
        for k in range(3):
            r = ref.bussi_step_rot(h, k, dt, draws[k], rdof, rke * (1 + 0.1 * k))
            rows.append([r["alpha"], r["alpha_rot"], r["ke"], r["cumulative"], r["instantaneous"], r["cumulative_rot"],
                         r["instantaneous_rot"], r["draws_left"]])
        bout[f"{name}/vel0"] = s.vel
        bout[f"{name}/args"] = np.array([dof, kT, tau, dt, rdof, rke])
        bout[f"{name}/draws"] = draws
        bout[f"{name}/rows"] = np.array(rows)
        ref.bussi_close(h)
    np.savez_compressed(os.path.join(HERE, "bussi.npz"), **bout)

    # F(k,t): the reference's own functions (compute_density_field, generate_fibonacci_sphere,
    # FieldAutocorrelationTracker.compute_field_autocorr), executed from /root/reference
    A = refpy.load().analysis
    fa = A.FieldAutocorrelationTracker.compute_field_autocorr
    fout = {}
    s = synth.make_system(400, replica=3)
    frames = synth.random_walk_frames(s, 6, sigma=0.3)
    kvec = A.generate_fibonacci_sphere(50) * 1.0
    rho = np.array([A.compute_density_field(refpy.Snapshot(frames[t]), kvec) for t in range(6)])
    F = np.array([[fa(None, rho[o], rho[o + l]) if o + l < 6 else np.nan for l in range(4)] for o in range(6)])
    fout["frames"] = frames
    fout["kvec"] = kvec
    fout["rho"] = rho
    fout["F"] = F
    fout["fib64"] = A.generate_fibonacci_sphere(64)
    np.savez_compressed(os.path.join(HERE, "fkt.npz"), **fout)
    print("wrote cavity_force.npz, bussi.npz, fkt.npz")


if __name__ == "__main__":
    main()
