#!/usr/bin/env python
"""tools/stamps_split.py -- variant 2 (split-phase kernel): bitwise check against variant 1, per-CTA
phase stamps and back-to-back step time."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402
from tools.prof_step import call, make  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mol", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=384)
    ap.add_argument("--ctas-per-sm", type=int, default=2)
    ap.add_argument("--unroll", type=int, default=2)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--ke-first", type=int, default=0)
    args = ap.parse_args()
    h = capi.Handle(0)
    h.set_tuning(ke_first=args.ke_first)
    n_mol = args.n_mol
    base, systems = make(h, n_mol, 8)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    st = capi.Stream()
    G = min(148 * args.ctas_per_sm, (base.N + args.threads - 1) // args.threads)
    # bitwise comparison on two fresh copies of the same system
    out = {}
    for variant in (1, 2, 3):
        h.set_tuning(variant=variant, threads=args.threads, ctas_per_sm=args.ctas_per_sm, unroll=args.unroll, stamps=0)
        d = systems[variant]
        call(h, "step", d, base, n_mol, p, a, st.ptr)
        out[variant] = (d["force"].numpy(st.ptr).copy(), d["vel"].numpy(st.ptr).copy(), h.force_read(st.ptr), h.bussi_read(st.ptr))
    for v in (2, 3):
        same_f = np.array_equal(out[1][0], out[v][0])
        same_v = np.array_equal(out[1][1], out[v][1])
        relf = np.abs(out[1][0] - out[v][0]).max() / np.abs(out[1][0]).max()
        relv = np.abs(out[1][1] - out[v][1]).max() / np.abs(out[1][1]).max()
        print(f"variant {v} vs 1: force bitwise {same_f} (max rel {relf:.2e})  vel bitwise {same_v} (max rel {relv:.2e})  energies",
              out[1][2][0], out[v][2][0], " alpha", out[1][3]["alpha"], out[v][3]["alpha"])
    for variant in (1, 2, 3, 2, 3):
        h.set_tuning(variant=variant, stamps=0)
        for k in range(10):
            call(h, "step", systems[k % 8], base, n_mol, p, a, st.ptr)
        capi.sync()
        e0, e1 = capi.Event(), capi.Event()
        e0.record(st.ptr)
        for k in range(args.steps):
            call(h, "step", systems[k % 8], base, n_mol, p, a, st.ptr)
        e1.record(st.ptr)
        us = e1.elapsed_ms_since(e0) / args.steps * 1e3
        print(f"variant {variant}: {us:.2f} us/step back to back  frac {148 * base.N / (us * 1e-6) / 6454.9e9:.3f}")
    for variant in (2, 3):
        h.set_tuning(variant=variant, stamps=1)
        print("variant", variant, "(2: dip-R | KE-R | Dq combine | forces | alpha combine | rescale;  "
              "3: dip-R | KE-R / helper merge | sync | forces / helper merge | sync | rescale;  4: dip-R | KE-R | [s2->s3 n/a] | forces | [n/a] | [n/a])  med max per phase, span")
        for k in range(5):
            call(h, "step", systems[k % 8], base, n_mol, p, a, st.ptr)
        capi.sync()
        for k in range(6):
            e0, e1 = capi.Event(), capi.Event()
            e0.record(st.ptr)
            call(h, "step", systems[(5 + k) % 8], base, n_mol, p, a, st.ptr)
            e1.record(st.ptr)
            ev = e1.elapsed_ms_since(e0) * 1e3
            s = h.debug_stamps(G).astype(np.float64)[:, :7] * 1e-3
            t0 = s[:, 0].min()
            rows = s[1:G] if variant == 3 else s
            ph = [rows[:, i + 1] - rows[:, i] for i in range(6)]
            print(f"{k:6d} | {ev:6.2f} | " + " | ".join(f"{np.median(x):6.2f} {x.max():6.2f}" for x in ph) + f" | {s[:, 6].max() - t0:6.2f}")
            if variant == 3:
                # folder CTA = CTA 0: [3] Final(F) published, [5] Final(K) published, since first start
                fs = s[0]
                st_ = s[1:G]
                rawf = h.debug_stamps(1024 + G).astype(np.float64)[1024] * 1e-3
                print("         folder: start %.2f, after pdl_wait %.2f, records loaded+voted %.2f, trees+barrier %.2f, warp-0 fold %.2f, finalize %.2f"
                      % (fs[0] - t0, fs[1] - t0, rawf[2] - t0, rawf[5] - t0, rawf[6] - t0, rawf[7] - t0))
                print("         streaming CTAs: start med %.2f max %.2f" % (np.median(st_[:, 0]) - t0, st_[:, 0].max() - t0))
                print("         folder: Final(F) out at %.2f, Final(K) out at %.2f | streaming CTAs: dipole records out med %.2f max %.2f, "
                      "KE records out med %.2f max %.2f, took Dq at med %.2f, took alpha at med %.2f"
                      % (fs[3] - t0, fs[5] - t0, np.median(st_[:, 1]) - t0, st_[:, 1].max() - t0, np.median(st_[:, 2]) - t0,
                         st_[:, 2].max() - t0, np.median(st_[:, 3]) - t0, np.median(st_[:, 5]) - t0))
                continue
            raw = h.debug_stamps(1024 + G).astype(np.float64) * 1e-3
            c0 = raw[:G, 2]  # start of the dipole combine = end of the KE pass
            d2 = raw[1024:1024 + G]
            print("         dipole combine, medians: records loaded+voted %.2f | trees+barrier %.2f | fold+vote in warp 0 %.2f | finalize %.2f | to end %.2f"
                  % (np.median(d2[:, 2] - c0), np.median(d2[:, 5] - d2[:, 2]), np.median(d2[:, 6] - d2[:, 5]),
                     np.median(d2[:, 7] - d2[:, 6]), np.median(raw[:G, 3] - d2[:, 7])))
            if variant == 4:
                rel = s - t0
                names = ["start", "dipR+pub", "KE pub", "F: Dq ready", "F: forces done", "K: alpha ready", "end"]
                print("        since first start, med/max: " + "  ".join(f"{n} {np.median(rel[:, i]):.2f}/{rel[:, i].max():.2f}" for i, n in enumerate(names)))

if __name__ == "__main__":
    main()
