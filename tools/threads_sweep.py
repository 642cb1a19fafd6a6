"""tools/threads_sweep.py -- step / force / Bussi call time for 320 / 352 / 384 threads per CTA and the automatic choice
of the step kernel, N = 262k .. 16M (CUDA events, rotating systems)."""
import os, sys
sys.path.insert(0, os.getcwd())
from cav_hoomd_b200 import capi, synth
from tools.prof_step import call, make
h = capi.Handle(0)
for n_mol in (262144, 500000, 1_000_000, 2_000_000, 4_000_000, 16_000_000):
    nsys = 8 if n_mol <= 1_000_000 else (4 if n_mol <= 4_000_000 else 2)
    base, systems = make(h, n_mol, nsys)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    st = capi.Stream()
    for threads in (320, 352, 384, 0):
        if threads:
            h.set_tuning(variant=3, threads=threads, ctas_per_sm=2, unroll=2, auto_threads=0)
        else:
            h.set_tuning(variant=3, threads=384, ctas_per_sm=2, unroll=2, auto_threads=1)
        row = []
        for kind in ("step", "force", "bussi"):
            for k in range(6):
                call(h, kind, systems[k % nsys], base, n_mol, p, a, st.ptr)
            capi.sync()
            e0, e1 = capi.Event(), capi.Event()
            e0.record(st.ptr)
            reps = 300 if n_mol <= 4_000_000 else 60
            for k in range(reps):
                call(h, kind, systems[k % nsys], base, n_mol, p, a, st.ptr)
            e1.record(st.ptr)
            row.append(f"{kind} {e1.elapsed_ms_since(e0) / reps * 1e3:8.2f}")
        print(f"N={base.N:9d} threads={threads if threads else 'auto'}: " + "  ".join(row), flush=True)
    for d in systems:
        for x in d.values():
            x.free()
