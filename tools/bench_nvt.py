#!/usr/bin/env python
"""tools/bench_nvt.py -- the thermostatted harness step on a B200, CUDA events.

Compares, per MD step at N particles, rotating over systems larger than L2:
  A  nve_kick_drift ; force ; nve_half_kick ; bussi          (thermostat as its own pass: 404 B/particle)
  B  nvt_step_one ; force ; nvt_step_two                      (thermostat folded in:        340 B/particle)
  C  nvt_step_one_rank1 ; force_rank1 ; nvt_step_two_rank1    (8f.2, force never written:   260 B/particle)
  D  md_step_one ; nvt_step_two_rank1                         (next dipole reduce inside step one: 220 B/particle, 2 launches)
  E  md_step_fused                                            (step two of t-1 + step one of t in ONE launch: 148 B/particle)
Algorithmic bytes: kick+drift 160 (vel, force, pos in; vel, pos out), force 84, half kick 96, Bussi 64;
rank-1: kick+drift reading charge instead of force 136, dipole reduce 52, half kick 72."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mol", type=int, nargs="+", default=[1_000_000, 16_000_000])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    h = capi.Handle(0)
    st = capi.Stream()
    p = capi.Params.make(0.01, 1e-3)
    rows = []
    for n_mol in args.n_mol:
        base = synth.make_system(n_mol)
        N = base.N
        n_sys = max(2, min(8, int(2e9 // (148 * N))))
        systems = []
        for k in range(n_sys):
            d = {f: capi.DeviceArray.from_numpy(getattr(base, f)) for f in ("pos", "charge", "image", "vel")}
            d["force"] = capi.DeviceArray.from_numpy(np.zeros((N, 4)))
            systems.append(d)
        dof = 3.0 * n_mol - 3
        a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
        dt = synth.DT_1FS
        for d in systems:
            h.bussi_ke(d["vel"], None, 0, n_mol, st.ptr)
        h.force_rank1(systems[0]["pos"], systems[0]["charge"], systems[0]["image"], N, base.box, base.L_typeid, p, st.ptr)

        def step(kind, d):
            if kind == "A":
                h.nve_kick_drift(d["pos"], d["vel"], d["force"], N, dt, st.ptr)
                h.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, st.ptr)
                h.nve_half_kick(d["vel"], d["force"], N, dt, st.ptr)
                h.bussi(d["vel"], None, 0, n_mol, a, st.ptr)
            elif kind == "B":
                h.nvt_step_one(d["pos"], d["vel"], d["force"], N, dt, 0, n_mol, a, st.ptr)
                h.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, st.ptr)
                h.nvt_step_two(d["vel"], d["force"], N, dt, 0, n_mol, st.ptr)
            elif kind == "C":
                h.nvt_step_one_rank1(d["pos"], d["vel"], None, d["charge"], N, dt, base.L_typeid, 1e-3, 0, n_mol, a, st.ptr)
                h.force_rank1(d["pos"], d["charge"], d["image"], N, base.box, base.L_typeid, p, st.ptr)
                h.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], N, dt, base.L_typeid, 1e-3, 0, n_mol, st.ptr)
            elif kind == "D":
                h.md_step_one(d["pos"], d["vel"], None, d["charge"], d["image"], N, dt, base.box, base.L_typeid, p, 0, n_mol,
                              a, st.ptr)
                h.nvt_step_two_rank1(d["vel"], None, d["charge"], d["pos"], N, dt, base.L_typeid, 1e-3, 0, n_mol, st.ptr)
            else:
                h.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], N, dt, base.box, base.L_typeid, p, 0, n_mol,
                                a, st.ptr)

        def run(kind, steps):
            e0, e1 = capi.Event(), capi.Event()
            capi.sync()
            e0.record(st.ptr)
            for k in range(steps):
                step(kind, systems[k % n_sys])
            e1.record(st.ptr)
            return e1.elapsed_ms_since(e0) / steps

        kinds = [("A", 404), ("B", 340)]
        kinds += [("C", 260), ("D", 220), ("E", 148)]
        print(f"N={N} systems={n_sys} peak={peak} GB/s")
        for kind, nbytes in kinds:
            run(kind, 5)
            ms = run(kind, args.steps)
            gbs = nbytes * N / (ms * 1e-3) / 1e9
            rows.append(dict(N=N, kind=kind, us_per_step=ms * 1e3, algorithmic_bytes_per_particle=nbytes, GBs=gbs,
                             frac=gbs / peak, M_particle_steps_per_s=N / (ms * 1e-3) / 1e6))
            print(f"  {kind}: {ms * 1e3:9.2f} us/step  {nbytes} B/particle -> {gbs:7.1f} GB/s ({gbs / peak:5.3f} of peak)  "
                  f"{N / (ms * 1e-3) / 1e6:9.1f} M particle-steps/s", flush=True)
        err = h.bussi_read(st.ptr)["err"]
        assert err == 0.0, err
        for d in systems:
            for x in d.values():
                x.free()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
