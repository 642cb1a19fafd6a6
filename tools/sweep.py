#!/usr/bin/env python
"""tools/sweep.py -- time the launch shapes of the streaming kernels (run on a B200 via gpurun).

For every (variant, threads, ctas_per_sm, unroll) it reports the CUDA-event time per call of
cavb200_force, cavb200_bussi and the fused cavb200_step, rotating over systems larger than L2,
and the fraction of the measured HBM peak the algorithmic bytes correspond to."""
import argparse
import itertools
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--pdl", type=int, default=1)
    args = ap.parse_args()
    peak = 6454.9
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    h = capi.Handle(0)
    n_mol = args.n_mol
    base = synth.make_system(n_mol)
    N = base.N
    systems = []
    for k in range(args.systems):
        d = {f: capi.DeviceArray.from_numpy(getattr(base, f)) for f in ("pos", "charge", "image", "vel")}
        d["force"] = capi.DeviceArray((N, 4), np.float64)
        systems.append(d)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    st = capi.Stream()

    def run(kind, steps):
        e0, e1 = capi.Event(), capi.Event()
        capi.sync()
        e0.record(st.ptr)
        for k in range(steps):
            d = systems[k % len(systems)]
            if kind == "force":
                h.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, st.ptr)
            elif kind == "bussi":
                h.bussi(d["vel"], None, 0, n_mol, a, st.ptr)
            elif kind == "step":
                h.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], N, base.box, base.L_typeid, p, 0, n_mol,
                       a, st.ptr)
            else:
                h.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, st.ptr)
                h.bussi(d["vel"], None, 0, n_mol, a, st.ptr)
        e1.record(st.ptr)
        return e1.elapsed_ms_since(e0) / steps

    shapes = [(v, t, c, u) for v in (1, 0) for (t, c, u) in ((384, 2, 2), (320, 2, 2), (256, 3, 2), (512, 2, 2), (1024, 1, 2), (256, 4, 2), (512, 1, 2), (512, 1, 4),
                                                             (256, 2, 4), (256, 1, 4), (384, 1, 4), (256, 1, 8), (192, 1, 8),
                                                             (128, 1, 8), (128, 2, 8))]
    if args.quick:
        shapes = [(1, 384, 2, 2), (0, 384, 2, 2), (1, 1024, 1, 2), (1, 512, 1, 4)]
    rows = []
    print(f"N={N} systems={len(systems)} peak={peak} GB/s")
    print("variant threads ctas unroll | force_us frac | bussi_us frac | step_us frac | force+bussi_us frac")
    for v, t, c, u in shapes:
        try:
            h.set_tuning(variant=v, threads=t, ctas_per_sm=c, unroll=u, pdl=args.pdl)
            res = {}
            for kind, nbytes in (("force", 84 * N), ("bussi", 64 * n_mol), ("step", 84 * N + 64 * n_mol),
                                 ("both", 84 * N + 64 * n_mol)):
                run(kind, 5)
                ms = run(kind, args.steps)
                res[kind] = (ms * 1e3, nbytes / (ms * 1e-3) / 1e9 / peak)
            rows.append(dict(variant=v, threads=t, ctas=c, unroll=u, **{k: list(x) for k, x in res.items()}))
            print(f"{v:7d} {t:7d} {c:4d} {u:6d} | " + " | ".join(f"{res[k][0]:8.2f} {res[k][1]:5.3f}" for k in
                                                              ("force", "bussi", "step", "both")), flush=True)
        except capi.CavbError as e:
            print(f"{v} {t} {c} {u}: {e}", flush=True)
    if args.out:
        json.dump(dict(N=N, peak=peak, rows=rows), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
