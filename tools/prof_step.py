#!/usr/bin/env python
"""tools/prof_step.py -- a short, fixed sequence of hot-path calls for ncu (and an N sweep).

    python tools/prof_step.py --kind force|bussi|step|both --steps 12 [--variant 1 ...]
    python tools/prof_step.py --nsweep      # CUDA-event time per call for N = 64k .. 16M
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402


def make(h, n_mol, nsys):
    base = synth.make_system(n_mol)
    systems = []
    for k in range(nsys):
        d = {f: capi.DeviceArray.from_numpy(getattr(base, f)) for f in ("pos", "charge", "image", "vel")}
        d["force"] = capi.DeviceArray((base.N, 4), np.float64)
        systems.append(d)
    return base, systems


def call(h, kind, d, base, n_mol, p, a, st):
    N = base.N
    if kind in ("force", "both"):
        h.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, st)
    if kind in ("bussi", "both"):
        h.bussi(d["vel"], None, 0, n_mol, a, st)
    if kind == "step":
        h.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], N, base.box, base.L_typeid, p, 0, n_mol, a, st)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="step")
    ap.add_argument("--n-mol", type=int, default=1_000_000)
    ap.add_argument("--systems", type=int, default=8)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--unroll", type=int, default=0)
    ap.add_argument("--nsweep", action="store_true")
    args = ap.parse_args()
    h = capi.Handle(0)
    for k in ("variant", "threads", "ctas_per_sm", "unroll"):
        v = getattr(args, k)
        if v:
            h.set_tuning(**{k: v})
    if args.variant == 0:
        h.set_tuning(variant=0)
    p = capi.Params.make(0.01, 1e-3)
    st = capi.Stream()
    if args.nsweep:
        print("n_mol  kind  us_per_call  GB/s(algorithmic)  frac_of_6454.9")
        for n_mol in (65536, 262144, 1_000_000, 4_000_000, 16_000_000):
            nsys = max(2, min(8, int(1.2e9 // (116 * n_mol)) + 1))
            base, systems = make(h, n_mol, nsys)
            dof = 3.0 * n_mol - 3
            a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
            for kind, nb in (("force", 84), ("bussi", 64), ("step", 148)):
                steps = max(10, min(200, int(4e8 // n_mol)))
                for k in range(3):
                    call(h, kind, systems[k % nsys], base, n_mol, p, a, st.ptr)
                e0, e1 = capi.Event(), capi.Event()
                capi.sync()
                e0.record(st.ptr)
                for k in range(steps):
                    call(h, kind, systems[k % nsys], base, n_mol, p, a, st.ptr)
                e1.record(st.ptr)
                us = e1.elapsed_ms_since(e0) / steps * 1e3
                gbs = nb * n_mol / (us * 1e-6) / 1e9
                print(f"{n_mol:9d} {kind:6s} {us:9.2f} {gbs:9.1f} {gbs / 6454.9:6.3f}", flush=True)
            for d in systems:
                for x in d.values():
                    x.free()
        return
    n_mol = args.n_mol
    base, systems = make(h, n_mol, args.systems)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    for k in range(args.steps):
        call(h, args.kind, systems[k % len(systems)], base, n_mol, p, a, st.ptr)
    capi.sync()
    print("done", args.kind, h.launch_count)


if __name__ == "__main__":
    main()
