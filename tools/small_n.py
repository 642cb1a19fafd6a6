#!/usr/bin/env python
"""tools/small_n.py -- us per call of force / Bussi / step for small systems, single-CTA kernel (tuning small_n) on / off."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402

h = capi.Handle(0)
p = capi.Params.make(0.01, 1e-3)
st = capi.Stream()
print("n_particles | small_n off: force bussi step both | small_n on: force bussi step both   (us per call, 2000 back-to-back calls)")
for n_mol in (100, 500, 1000, 1500, 2000, 3000, 4000, 6000, 8000):
    s = synth.make_system(n_mol)
    d = {f: capi.DeviceArray.from_numpy(getattr(s, f)) for f in ("pos", "charge", "image", "vel")}
    d["force"] = capi.DeviceArray((s.N, 4), np.float64)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    row = []
    for small in (0, 1 << 20):
        h.set_tuning(small_n=small)
        for kind in ("force", "bussi", "step", "both"):
            def call():
                if kind in ("force", "both"):
                    h.force(d["pos"], d["charge"], d["image"], d["force"], s.N, s.box, s.L_typeid, p, st.ptr)
                if kind in ("bussi", "both"):
                    h.bussi(d["vel"], None, 0, n_mol, a, st.ptr)
                if kind == "step":
                    h.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], s.N, s.box, s.L_typeid, p, 0, n_mol, a, st.ptr)
            for _ in range(20):
                call()
            capi.sync()
            h.debug_delay(20_000_000, st.ptr)
            e0, e1 = capi.Event(), capi.Event()
            e0.record(st.ptr)
            for _ in range(2000):
                call()
            e1.record(st.ptr)
            row.append(e1.elapsed_ms_since(e0) / 2000 * 1e3)
    print(f"{s.N:8d} | " + " ".join(f"{x:6.2f}" for x in row[:4]) + " | " + " ".join(f"{x:6.2f}" for x in row[4:]), flush=True)
