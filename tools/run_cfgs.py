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
for variant, threads, ctas in ((2, 384, 2),):
    h.set_tuning(variant=variant, threads=threads, ctas_per_sm=ctas, unroll=2)
    for kind, nb in (("step", 148), ("bussi", 64)):
        for k in range(10):
            call(h, kind, systems[k % 8], base, n_mol, p, a, st.ptr)
        capi.sync()
        e0, e1 = capi.Event(), capi.Event()
        e0.record(st.ptr)
        for k in range(400):
            call(h, kind, systems[k % 8], base, n_mol, p, a, st.ptr)
        e1.record(st.ptr)
        us = e1.elapsed_ms_since(e0) / 400 * 1e3
        print(f"variant {variant} {threads}x{ctas} {kind:5s}: {us:6.2f} us  frac {nb * base.N / (us * 1e-6) / 6454.9e9:.3f}", flush=True)
