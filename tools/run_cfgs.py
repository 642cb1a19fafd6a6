"""tools/run_cfgs.py -- step / force / Bussi time per launch shape and step-kernel variant (CUDA events,
400 back-to-back calls rotating over 8 systems of 1M particles)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from cav_hoomd_b200 import capi, synth
from tools.prof_step import call, make
h = capi.Handle(0)
n_mol = 1_000_000
base, systems = make(h, n_mol, 8)
p = capi.Params.make(0.01, 1e-3)
dof = 3.0 * n_mol - 3
a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
st = capi.Stream()
shapes = ((2, 384, 2, 2), (3, 384, 2, 2), (3, 768, 1, 2), (3, 320, 2, 2), (3, 256, 3, 2), (3, 512, 1, 4), (3, 1024, 1, 2),
          (3, 640, 1, 2), (3, 256, 2, 4), (3, 352, 2, 2))
for variant, threads, ctas, unroll in shapes:
    try:
        h.set_tuning(variant=variant, threads=threads, ctas_per_sm=ctas, unroll=unroll)
        for kind, nb in (("step", 148),) + ((("force", 84), ("bussi", 64)) if variant == 2 else ()):
            for k in range(10):
                call(h, kind, systems[k % 8], base, n_mol, p, a, st.ptr)
            capi.sync()
            e0, e1 = capi.Event(), capi.Event()
            e0.record(st.ptr)
            for k in range(400):
                call(h, kind, systems[k % 8], base, n_mol, p, a, st.ptr)
            e1.record(st.ptr)
            us = e1.elapsed_ms_since(e0) / 400 * 1e3
            print(f"variant {variant} {threads}x{ctas} u{unroll} {kind:5s}: {us:6.2f} us  frac {nb * base.N / (us * 1e-6) / 6454.9e9:.3f}", flush=True)
    except capi.CavbError as e:
        print(variant, threads, ctas, unroll, e, flush=True)
