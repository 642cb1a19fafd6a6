#!/usr/bin/env python
"""tools/ring_dump.py -- per-launch device timestamps of the step kernel (tuning stamps = 2): for a run of back-to-back
cavb200_step launches prints start / end of every launch relative to the first start, the duration (first CTA start ->
last CTA end), the start-to-start period and the gap between one launch's end and the next one's start."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
h = capi.Handle(0)
base = synth.make_system(n_mol)
N = base.N
systems = []
for k in range(8):
    d = {f: capi.DeviceArray.from_numpy(getattr(base, f)) for f in ("pos", "charge", "image", "vel")}
    d["force"] = capi.DeviceArray((N, 4), np.float64)
    systems.append(d)
p = capi.Params.make(0.01, 1e-3)
dof = 3.0 * n_mol - 3
a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.1, (dof - 1) / 2)
st = capi.Stream()


def step(k):
    d = systems[k % 8]
    h.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], N, base.box, base.L_typeid, p, 0, n_mol, a, st.ptr)


for k in range(10):
    step(k)
capi.sync()
h.set_tuning(stamps=2)
h.debug_launch_ring(reset=True, read=False)
h.debug_delay(2_000_000, st.ptr)
e0, e1 = capi.Event(), capi.Event()
e0.record(st.ptr)
for k in range(K):
    step(k)
e1.record(st.ptr)
ms = e1.elapsed_ms_since(e0)
ring, ep = h.debug_launch_ring()
slots = [(ep - i) & 2047 for i in range(K)][::-1]
s0 = ring[slots, 0].astype(np.int64)
s1 = ring[slots, 1].astype(np.int64)
t0 = s0[0]
print(f"N={N} K={K} event span / K = {1e3 * ms / K:.2f} us")
print(" launch | start us | end us | duration | period | gap to next start")
for i in range(K):
    per = (s0[i + 1] - s0[i]) * 1e-3 if i + 1 < K else float("nan")
    gap = (s0[i + 1] - s1[i]) * 1e-3 if i + 1 < K else float("nan")
    print(f" {i:6d} | {(s0[i] - t0) * 1e-3:8.2f} | {(s1[i] - t0) * 1e-3:8.2f} | {(s1[i] - s0[i]) * 1e-3:8.2f} | {per:6.2f} | {gap:6.2f}")
