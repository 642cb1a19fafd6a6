"""tools/md_shapes.py -- cavb200_md_step_fused launch shapes (tuning md_shape), 1M and 4M particles, CUDA events."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from cav_hoomd_b200 import capi, synth
h = capi.Handle(0)
p = capi.Params.make(0.01, 1e-3)
st = capi.Stream()
for n_mol in (1_000_000, 4_000_000):
    base = synth.make_system(n_mol)
    N = base.N
    n_sys = 8 if n_mol == 1_000_000 else 3
    systems = [{f: capi.DeviceArray.from_numpy(getattr(base, f)) for f in ("pos", "charge", "image", "vel")} for _ in range(n_sys)]
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    h.bussi_ke(systems[0]["vel"], None, 0, n_mol, st.ptr)
    h.force_rank1(systems[0]["pos"], systems[0]["charge"], systems[0]["image"], N, base.box, base.L_typeid, p, st.ptr)
    for rep in range(2):
        for shape, name in ((0, "768x1"), (1, "384x2")):
            h.set_tuning(md_shape=shape)
            def one(k):
                d = systems[k % n_sys]
                h.md_step_fused(d["pos"], d["vel"], None, d["charge"], d["image"], N, synth.DT_1FS, base.box, base.L_typeid, p, 0,
                                n_mol, a, st.ptr)
            for k in range(5):
                one(k)
            capi.sync()
            e0, e1 = capi.Event(), capi.Event()
            e0.record(st.ptr)
            for k in range(200):
                one(k)
            e1.record(st.ptr)
            us = e1.elapsed_ms_since(e0) / 200 * 1e3
            assert h.bussi_read(st.ptr)["err"] == 0.0
            print(f"N={N} md_step_fused {name}: {us:7.2f} us  {N / us / 1e3:8.1f} G particle-steps/s  {148 * N / (us * 1e-6) / 6454.9e9:.3f} of 148 B roofline", flush=True)
    for d in systems:
        for x in d.values():
            x.free()
