#!/usr/bin/env python
"""tools/sanitize_small.py -- one small invocation of every kernel of libcavb200 (for compute-sanitizer):
force, Bussi, step (variants 0/1/2/3), index-list group, sharded step (one rank), NVE / NVT / rank-1 harness steps,
md_step_one, net-force add, trackers, rhok, fkt, host-buffer step (blocking and slots).  Sizes are small so the bounded hand-off spins survive the sanitizer's slowdown."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402


def main():
    h = capi.Handle(0)
    p = capi.Params.make(0.01, 1e-3)
    for n_mol in (5, 3001):
        s = synth.make_system(n_mol)
        dof = max(3.0 * n_mol - 3.0, 0.0)
        a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.3, max(dof - 1, 0) / 2)
        dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray((s.N, 4), np.float64)
        for variant in (0, 1, 2, 3):
            h.set_tuning(variant=variant, threads=128, ctas_per_sm=1, unroll=2)
            h.force(dev["pos"], dev["charge"], dev["image"], d_f, s.N, s.box, s.L_typeid, p)
            h.bussi(dev["vel"], None, 0, n_mol, a)
            h.step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, s.box, s.L_typeid, p, 0, n_mol, a)
            capi.sync()
        idx = capi.DeviceArray.from_numpy(np.arange(0, n_mol, 2, dtype=np.uint32))
        h.bussi(dev["vel"], idx, 0, (n_mol + 1) // 2, a)
        h.shard_step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, 0, s.box, s.L_typeid, p, 0, n_mol, a)
        h.nve_kick_drift(dev["pos"], dev["vel"], d_f, s.N, 1.0)
        h.nve_half_kick(dev["vel"], d_f, s.N, 1.0)
        # harness steps (8f.1, 8f.2), trackers (8f.4)
        d_net = capi.DeviceArray.from_numpy(np.zeros((s.N, 4)))
        h.bussi_ke(dev["vel"], None, 0, n_mol)
        h.nvt_step_one(dev["pos"], dev["vel"], d_f, s.N, 1.0, 0, n_mol, a)
        h.nvt_step_two(dev["vel"], d_f, s.N, 1.0, 0, n_mol)
        h.force_rank1(dev["pos"], dev["charge"], dev["image"], s.N, s.box, s.L_typeid, p)
        h.net_force_add_rank1(d_net, dev["charge"], dev["pos"], s.N, s.L_typeid, 1e-3)
        h.nvt_step_one_rank1(dev["pos"], dev["vel"], d_net, dev["charge"], s.N, 1.0, s.L_typeid, 1e-3, 0, n_mol, a)
        h.md_step_one(dev["pos"], dev["vel"], None, dev["charge"], dev["image"], s.N, 1.0, s.box, s.L_typeid, p, 0, n_mol, a)
        h.md_step_fused(dev["pos"], dev["vel"], d_net, dev["charge"], dev["image"], s.N, 1.0, s.box, s.L_typeid, p, 0, n_mol, a)
        h.nvt_step_two_rank1(dev["vel"], d_net, dev["charge"], dev["pos"], s.N, 1.0, s.L_typeid, 1e-3, 0, n_mol)
        h.track_open(3)
        h.track_set_reference()
        for t in range(5):
            h.track_record(t, dev["vel"], s.N)
        rec, total = h.track_read(8)
        assert total == 5 and len(rec) == 3
        capi.sync()
        bufs = [dict(pos=s.pos.copy(), charge=s.charge.copy(), image=s.image.copy(), force=np.zeros((s.N, 4)), vel=s.vel.copy())
                for _ in range(3)]
        for k, q in enumerate(bufs):
            h.step_host_submit(k, q["pos"], q["charge"], q["image"], q["force"], q["vel"], s.N, s.box, s.L_typeid, p, 0, n_mol, a)
        for k in range(3):
            h.step_host_wait(k)
        en, bo = h.step_host(s.pos, s.charge, s.image, np.zeros((s.N, 4)), s.vel.copy(), s.N, s.box, s.L_typeid, p, 0, n_mol, a)
        frames = np.ascontiguousarray(np.stack([s.pos[:, :3], s.pos[:, :3] + 0.1]))
        kvec = synth.fibonacci_sphere(7)
        d_fr, d_k = capi.DeviceArray.from_numpy(frames), capi.DeviceArray.from_numpy(kvec)
        d_rho = capi.DeviceArray((2, 7, 2), np.float64)
        h.rhok(d_fr, 3, s.N * 3, s.N, 2, d_k, 7, d_rho)
        d_F = capi.DeviceArray((2, 2), np.float64)
        h.fkt(d_rho, 2, 7, 2, 2, d_F)
        capi.sync()
        print("ok", n_mol, h.force_read()[0], h.bussi_read()["alpha"], d_F.numpy()[0])
    h.close()


if __name__ == "__main__":
    main()
