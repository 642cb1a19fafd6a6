#!/usr/bin/env python
"""tools/cluster_n.py -- us per call of force / Bussi / step for mid-size systems: the persistent multi-CTA kernels against the
single-cluster kernel k_cluster (tuning cluster_n, cluster_ctas = 8 / 16); also checks that the results are the same bits."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402

h = capi.Handle(0)
p = capi.Params.make(0.01, 1e-3)
st = capi.Stream()
sizes = [int(x) for x in sys.argv[1:]] or [500, 1000, 1500, 2000, 4000, 8000, 16000, 32000, 65000, 131000]
print("n_particles | persistent: force bussi step both | cluster of 8: force bussi step both | cluster of 16: ...   (us per call, 2000 "
      "back-to-back calls); same bits")
for n_mol in sizes:
    s = synth.make_system(n_mol)
    d = {f: capi.DeviceArray.from_numpy(getattr(s, f)) for f in ("pos", "charge", "image", "vel")}
    d["force"] = capi.DeviceArray((s.N, 4), np.float64)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
    row, outs = [], []
    for cl_n, cl_c in ((0, 8), (1 << 22, 8), (1 << 22, 16)):
        h.set_tuning(cluster_n=cl_n, cluster_ctas=cl_c)
        # results of one step from the same state
        d["vel"].upload(s.vel)
        h.bussi_reset(st.ptr)
        try:
            h.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], s.N, s.box, s.L_typeid, p, 0, n_mol, a, st.ptr)
            outs.append((d["force"].numpy(st.ptr), d["vel"].numpy(st.ptr), h.force_read(st.ptr)[0], h.bussi_read(st.ptr)["alpha"]))
        except capi.CavbError as e:
            print(f"  cluster_ctas={cl_c}: {e}")
            row += [float("nan")] * 4
            outs.append(None)
            continue
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
    same = all(o is None or (np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
                             and np.array_equal(o[2], outs[0][2]) and o[3] == outs[0][3]) for o in outs[1:])
    print(f"{s.N:8d} | " + " | ".join(" ".join(f"{x:6.2f}" for x in row[i:i + 4]) for i in (0, 4, 8)) + f" | {same}", flush=True)
