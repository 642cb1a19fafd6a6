#!/usr/bin/env python
"""tools/ab.py -- A/B timing of force / Bussi / step / force+Bussi under tuning overrides (run on a B200).

    python tools/ab.py [--n-mol 1000000] [--steps 200] [--reps 3] [--index-list] --set variant=3,pdl=1 --set variant=3,pdl=0 ...

Each --set is one row: comma-separated tuning keys of cavb200_set_tuning applied to a fresh handle.  Times are CUDA
events around `steps` back-to-back calls, rotating over 8 systems (inputs larger than L2), best of `reps`.
With CAVB200_LIB=path the same script times another build of the library."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mol", type=int, default=1_000_000)
    ap.add_argument("--systems", type=int, default=8)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--kinds", default="force,bussi,step,both")
    ap.add_argument("--index-list", action="store_true", help="pass the Bussi group as an index list (0..n_mol-1)")
    ap.add_argument("--set", action="append", default=[], dest="sets")
    args = ap.parse_args()
    peak = 6454.9
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n_mol = args.n_mol
    base = synth.make_system(n_mol)
    N = base.N
    systems = []
    for k in range(args.systems):
        d = {f: capi.DeviceArray.from_numpy(getattr(base, f)) for f in ("pos", "charge", "image", "vel")}
        d["force"] = capi.DeviceArray((N, 4), np.float64)
        systems.append(d)
    gidx = capi.DeviceArray.from_numpy(np.arange(n_mol, dtype=np.uint32)) if args.index_list else None
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    st = capi.Stream()
    print(f"lib={capi.LIB_PATH} N={N} systems={len(systems)} steps={args.steps} peak={peak} GB/s index_list={args.index_list}")
    nbytes = {"force": 84 * N, "bussi": 64 * n_mol, "step": 84 * N + 64 * n_mol, "both": 84 * N + 64 * n_mol}
    for spec in (args.sets or [""]):
        h = capi.Handle(0)
        kw = {}
        for item in filter(None, spec.split(",")):
            k, v = item.split("=")
            kw[k] = int(v)
        if kw:
            h.set_tuning(**kw)

        def run(kind, steps):
            e0, e1 = capi.Event(), capi.Event()
            capi.sync()
            e0.record(st.ptr)
            for k in range(steps):
                d = systems[k % len(systems)]
                if kind in ("force", "both"):
                    h.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, st.ptr)
                if kind in ("bussi", "both"):
                    h.bussi(d["vel"], gidx, 0, n_mol, a, st.ptr)
                if kind == "step":
                    h.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], N, base.box, base.L_typeid, p, 0, n_mol,
                           a, st.ptr)
            e1.record(st.ptr)
            return e1.elapsed_ms_since(e0) / steps

        cells = []
        for kind in args.kinds.split(","):
            try:
                run(kind, 5)
                us = 1e3 * min(run(kind, args.steps) for _ in range(args.reps))
                cells.append(f"{kind} {us:7.2f} us {nbytes[kind] / (us * 1e-6) / 1e9 / peak:5.3f}")
            except capi.CavbError as e:
                cells.append(f"{kind} ERR {e.code}")
        bo = h.bussi_read(st.ptr)
        print(f"[{spec or 'default':28s}] " + " | ".join(cells) + f" | err={bo['err']}", flush=True)
        h.close()


if __name__ == "__main__":
    main()
