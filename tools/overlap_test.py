#!/usr/bin/env python
"""tools/overlap_test.py -- feasibility: cavity-force kernel and Bussi kernel as two concurrent
half-machine pipelines (one CTA per SM each, two streams, two handles) instead of one fused launch.
The write half of a step is L2-store-ingest bound while HBM idles, the read half is HBM bound
(profiles/microwb_r1b.txt); two skewed pipelines overlap one's write half with the other's read half."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402
from tools.prof_step import make  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mol", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=400)
    args = ap.parse_args()
    n_mol = args.n_mol
    hf, hb = capi.Handle(0), capi.Handle(0)
    base, systems = make(hf, n_mol, 8)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    sf, sb = capi.Stream(), capi.Stream()
    N = base.N

    def run(cfg_f, cfg_b, fused=False):
        hf.set_tuning(**cfg_f)
        hb.set_tuning(**cfg_b)
        def go(steps):
            for k in range(steps):
                d = systems[k % 8]
                if fused:
                    hf.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], N, base.box, base.L_typeid, p, 0, n_mol, a, sf.ptr)
                else:
                    hf.force(d["pos"], d["charge"], d["image"], d["force"], N, base.box, base.L_typeid, p, sf.ptr)
                    hb.bussi(d["vel"], None, 0, n_mol, a, sb.ptr)
        go(20)
        capi.sync()
        e0, e1, e2 = capi.Event(), capi.Event(), capi.Event()
        t0 = time.perf_counter()
        e0.record(sf.ptr)
        go(args.steps)
        e1.record(sf.ptr)
        e2.record(sb.ptr)
        capi.sync()
        wall = (time.perf_counter() - t0) / args.steps * 1e6
        ms_f = e1.elapsed_ms_since(e0)
        ms_b = e2.elapsed_ms_since(e0)
        us = max(ms_f, ms_b) / args.steps * 1e3
        print(f"force {cfg_f} bussi {cfg_b} fused={fused}: {us:7.2f} us/step (force stream {ms_f / args.steps * 1e3:.2f}, "
              f"bussi stream {ms_b / args.steps * 1e3:.2f}, host wall {wall:.2f})  frac {148 * N / (us * 1e-6) / 6454.9e9:.3f}", flush=True)

    run(dict(threads=384, ctas_per_sm=2, unroll=2), dict(threads=384, ctas_per_sm=2, unroll=2), fused=True)
    for tf, tb in ((384, 384), (512, 512), (512, 384), (384, 512), (768, 384), (768, 768), (1024, 512), (512, 1024)):
        for uf, ub in ((2, 2), (2, 4), (4, 2)):
            if (tf > 512 and uf > 2) or (tb > 512 and ub > 2):
                continue
            if (uf == 4 and tf > 512) or (ub == 4 and tb > 512):
                continue
            run(dict(threads=tf, ctas_per_sm=1, unroll=uf), dict(threads=tb, ctas_per_sm=1, unroll=ub))
    # both kernels full machine on two streams (they serialize or interleave as the scheduler likes)
    run(dict(threads=384, ctas_per_sm=2, unroll=2), dict(threads=384, ctas_per_sm=2, unroll=2))


if __name__ == "__main__":
    main()
