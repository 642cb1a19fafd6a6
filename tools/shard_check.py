#!/usr/bin/env python
"""tools/shard_check.py -- particle-sharded step across GPUs: parity + timing (run under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/shard_check.py [--n-mol 16000000] [--steps 50]

Every rank owns a contiguous block of particles (cav_hoomd_b200.shard.shard_system); one 160-byte
record per rank crosses NVLink per step, either through ncclAllGather (mode nccl) or by direct
stores into the peers' mailboxes from inside the reduce kernel (mode nvlink).  Checks, for both
modes: forces of the local block, energies, photon index and alpha against the UNSHARDED CPU oracle
(small N), and that all ranks agree bit for bit; then times `--steps` steps of the BASELINE config
(16M particles by default) with CUDA events, max over ranks."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cav_hoomd_b200 import capi, shard, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mol", type=int, default=16_000_000)
    ap.add_argument("--check-n", type=int, default=200_000)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--modes", default="nvlink,nccl")
    ap.add_argument("--variant", type=int, default=3, help="tuning variant of the handle")
    ap.add_argument("--set", default="", help="extra tuning keys for the timed handle, e.g. pdl=0")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    out = {"world": world}
    p = capi.Params.make(0.01, 1e-3)

    def run(h, s, nsteps, check):
        sub, off, (first, n) = shard.shard_system(s, rank, world)
        n_mol_total = s.N - 1
        dof = 3.0 * n_mol_total - 3.0
        a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.3, (dof - 1) / 2)
        dev = {k: capi.DeviceArray.from_numpy(getattr(sub, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray((max(sub.N, 1), 4), np.float64)
        st = capi.Stream()
        def one():
            h.shard_step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], sub.N, off, s.box, s.L_typeid, p, first,
                         n, a, st.ptr)
        if check:
            h.bussi_reset(st.ptr)
            one()
            en, dip, ph = h.force_read(st.ptr)
            bo = h.bussi_read(st.ptr)
            return sub, off, d_f.numpy(st.ptr)[:sub.N], dev["vel"].numpy(st.ptr), en, dip, ph, bo, a
        for _ in range(5):
            one()
        st.sync()
        dist.barrier()
        e0, e1 = capi.Event(), capi.Event()
        e0.record(st.ptr)
        for _ in range(nsteps):
            one()
        e1.record(st.ptr)
        ms = e1.elapsed_ms_since(e0)
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / nsteps

    for mode in args.modes.split(","):
        h = capi.Handle(local)
        h.set_tuning(variant=args.variant)
        if args.set:
            h.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in args.set.split(","))})
        shard.bootstrap(h, dist, mode)
        # ---- parity on a size the CPU oracle finishes in a second ----
        s = synth.make_system(args.check_n)
        sub, off, f, v, en, dip, ph, bo, a = run(h, s, 1, True)
        if rank == 0:
            from oracle import oracle as O
            co = O.COracle()
            ref = co.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
            vref = s.vel.copy()
            alpha, ke = co.bussi_step(vref, np.arange(args.check_n, dtype=np.uint32), a.dof, synth.DT_1FS, synth.KT_100K,
                                      synth.TAU_5PS, a.r_normal, a.gamma_draw, np.zeros(2))
            exact = co.dipole_exact(s.pos, s.charge, s.image, s.box, ref["photon_idx"])
            ref_blob = [ref["force"], vref, ref["energies"], alpha, ref["photon_idx"], exact]
        else:
            ref_blob = None
        blob = [ref_blob]
        dist.broadcast_object_list(blob, src=0)
        rf, rv, ren, ralpha, rph, exact = blob[0]
        lo, hi = off, off + sub.N
        ok = (ph == rph and np.abs(f - rf[lo:hi]).max() <= 1e-10 * np.abs(rf).max()
              and np.allclose(en, ren, rtol=1e-10) and abs(bo["alpha"] - ralpha) <= 1e-12 * abs(ralpha)
              and np.allclose(v, rv[lo:hi], rtol=1e-12, atol=0) and np.all(np.abs(dip - exact) <= 2 * np.spacing(np.abs(exact))))
        # all ranks must hold bitwise identical scalars
        allsc = [None] * world
        dist.all_gather_object(allsc, (en.tobytes(), dip.tobytes(), bo["alpha"], ph))
        same = all(x == allsc[0] for x in allsc)
        oks = [None] * world
        dist.all_gather_object(oks, bool(ok))
        # ---- timing on the BASELINE size ----
        big = synth.make_system(args.n_mol)
        ms = run(h, big, args.steps, False)
        out[mode] = {"parity_all_ranks": all(oks), "ranks_bitwise_identical": same, "n_particles": big.N,
                     "ms_per_step": ms, "M_particle_steps_per_s": big.N / (ms * 1e-3) / 1e6,
                     "frac_of_hbm_roofline_per_gpu": 148 * big.N / world / (ms * 1e-3) / 1e9 / 6454.9}
        h.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0 if all(out[m]["parity_all_ranks"] and out[m]["ranks_bitwise_identical"] for m in args.modes.split(",")) else 1


if __name__ == "__main__":
    sys.exit(main())
