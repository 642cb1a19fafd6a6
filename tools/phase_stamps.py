#!/usr/bin/env python
"""tools/phase_stamps.py -- where the time of the cooperative kernel goes (per-CTA globaltimer stamps).

Prints, for a few launches, the spread over CTAs of: start skew, reduce duration, barrier wait,
combine, apply, and the total from first start to last finish, next to the CUDA-event time."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402
from tools.prof_step import call, make  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="step")
    ap.add_argument("--n-mol", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=384)
    ap.add_argument("--ctas-per-sm", type=int, default=2)
    ap.add_argument("--unroll", type=int, default=2)
    args = ap.parse_args()
    h = capi.Handle(0)
    h.set_tuning(variant=1, threads=args.threads, ctas_per_sm=args.ctas_per_sm, unroll=args.unroll, stamps=1)
    n_mol = args.n_mol
    base, systems = make(h, n_mol, 8)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    st = capi.Stream()
    G = min(148 * args.ctas_per_sm, (base.N + args.threads - 1) // args.threads)
    for k in range(5):
        call(h, args.kind, systems[k % 8], base, n_mol, p, a, st.ptr)
    capi.sync()
    print(f"kind={args.kind} N={base.N} grid={G}x{args.threads} unroll={args.unroll}  (all times us)")
    print("launch | event | first..last_start | reduce med/max | barrier-wait med/max | combine med | apply med/max | span")
    for k in range(6):
        e0, e1 = capi.Event(), capi.Event()
        e0.record(st.ptr)
        call(h, args.kind, systems[(5 + k) % 8], base, n_mol, p, a, st.ptr)
        e1.record(st.ptr)
        ev = e1.elapsed_ms_since(e0) * 1e3
        raw = h.debug_stamps(G).astype(np.float64) * 1e-3
        s = raw[:, :5]
        sub = (np.median(raw[:, 5] - raw[:, 2]), np.median(raw[:, 6] - raw[:, 5]), np.median(raw[:, 7] - raw[:, 6]))
        t0 = s[:, 0].min()
        red = s[:, 1] - s[:, 0]
        bar = s[:, 2] - s[:, 1]
        comb = s[:, 3] - s[:, 2]
        app = s[:, 4] - s[:, 3]
        print(f"{k:6d} | {ev:6.2f} | {s[:, 0].max() - t0:6.2f} | {np.median(red):6.2f} {red.max():6.2f} | "
              f"{np.median(bar):6.2f} {bar.max():6.2f} | {np.median(comb):6.2f} | {np.median(app):6.2f} {app.max():6.2f} | "
              f"{s[:, 4].max() - t0:6.2f}   last-reduce-end={s[:, 1].max() - t0:6.2f} barrier-exit={s[:, 2].min() - t0:6.2f}..{s[:, 2].max() - t0:6.2f} combine[load+vote {sub[0]:.2f} fold+tree {sub[1]:.2f} finalize {sub[2]:.2f}]")


if __name__ == "__main__":
    main()
