#!/usr/bin/env python
"""tools/bench_fkt.py -- BASELINE configs[4]: F(k,t) on a 1M-particle trajectory, 64 wave vectors, 1000
time origins.  rho[t][k] = sum_j exp(i k.r_j(t)) for a batch of T frames per launch (frames are
generated on the host as a random walk and rotated so the timed input is larger than L2), then
F[o][l] for all origins/lags.  Reports (particle, k) sincos pairs/s, the time 1000 origins take, the
GB/s of position traffic, and the NumPy restatement of the reference (oracle) timed on a bounded sample.

    python tools/bench_fkt.py [--n 1000000] [--K 64] [--frames 16] [--origins 1000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cav_hoomd_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--frames", type=int, default=16, help="frames per launch")
    ap.add_argument("--origins", type=int, default=1000)
    ap.add_argument("--stride", type=int, default=4, choices=[3, 4])
    ap.add_argument("--f32", action="store_true", help="float32 xyz frames (GSD layout) through cavb200_rhok_f32")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    N, K, T = args.n + 1, args.K, args.frames
    # several GPUs (torchrun): frames are independent, every rank takes its blocks of T frames of the `origins` frames
    # (cav_hoomd_b200.replicas.frames_for_rank), no data-path collective; time = max over ranks
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    h = capi.Handle(local)
    if os.environ.get("RHOK_THREADS"):
        h.set_tuning(rhok_threads=int(os.environ["RHOK_THREADS"]))
    s = synth.make_system(args.n)
    rng = np.random.default_rng(7)
    kvec = synth.fibonacci_sphere(K) * 1.0
    d_k = capi.DeviceArray.from_numpy(kvec)
    # T frames: random walk, sigma 0.05 Bohr, HOOMD Scalar4 layout or xyz
    frames = np.empty((T, N, args.stride))
    cur = s.pos.copy()
    for t in range(T):
        cur[:, :3] += 0.05 * rng.standard_normal((N, 3))
        frames[t, :, :3] = cur[:, :3]
        if args.stride == 4:
            frames[t, :, 3] = cur[:, 3]
    nbuf = max(2, int(400e6 // frames.nbytes) + 1)
    if args.f32:
        frames32 = np.ascontiguousarray(frames[:, :, :3], dtype=np.float32)
        d_frames = [capi.DeviceArray.from_numpy(frames32) for _ in range(nbuf)]
        _rhok = h.rhok
        h.rhok = lambda d, stride, fs, N_, T_, dk, K_, drho, stream=None: h.rhok_f32(d, 3 * N_, N_, T_, dk, K_, drho, stream)
    else:
        d_frames = [capi.DeviceArray.from_numpy(frames) for _ in range(nbuf)]
    d_rho = capi.DeviceArray((T, K, 2), np.float64)
    st = capi.Stream()
    for k in range(3):
        h.rhok(d_frames[k % nbuf], args.stride, N * args.stride, N, T, d_k, K, d_rho, st.ptr)
    capi.sync()
    launches = max(4, min(40, args.origins // T))
    if world > 1:
        from cav_hoomd_b200 import replicas
        launches = len(replicas.frames_for_rank(args.origins, rank, world, block=T))
        dist.barrier()
    e0, e1 = capi.Event(), capi.Event()
    l0 = h.launch_count
    e0.record(st.ptr)
    for k in range(launches):
        h.rhok(d_frames[k % nbuf], args.stride, N * args.stride, N, T, d_k, K, d_rho, st.ptr)
    e1.record(st.ptr)
    ms = e1.elapsed_ms_since(e0)
    if world > 1:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_all = float(t.item())
        if rank == 0:
            print(json.dumps({"workload": f"F(k,t) over {world} GPUs: {args.origins} frames of N={N} particles, K={K}, blocks of {T} frames "
                                          "round-robin over ranks, no collective in the data path",
                              "n_gpus": world, "seconds_for_origins_field_sum": ms_all * 1e-3,
                              "sincos_pairs_per_s": N * K * args.origins / (ms_all * 1e-3), "launches_per_rank": launches}))
        dist.destroy_process_group()
        return
    per_frame_ms = ms / (launches * T)
    pairs_per_s = N * K / (per_frame_ms * 1e-3)
    rho = d_rho.numpy(st.ptr)
    # correlation over `origins` frames x `origins` lags of synthetic rho (tiny next to the field sum)
    To = args.origins
    big = np.tile(rho, (To // T + 1, 1, 1))[:To]
    d_big = capi.DeviceArray.from_numpy(np.ascontiguousarray(big))
    d_F = capi.DeviceArray((To, To), np.float64)
    h.fkt(d_big, To, K, To, To, d_F, st.ptr)
    capi.sync()
    e0.record(st.ptr)
    h.fkt(d_big, To, K, To, To, d_F, st.ptr)
    e1.record(st.ptr)
    ms_corr = e1.elapsed_ms_since(e0)
    # check one frame against the NumPy restatement and time it (bounded sample)
    out = {
        "workload": f"F(k,t): N={N} particles, K={K} wave vectors, {T} frames per launch, stride {args.stride} doubles/particle",
        "ms_per_frame": per_frame_ms, "sincos_pairs_per_s": pairs_per_s,
        "seconds_for_origins": per_frame_ms * 1e-3 * args.origins + ms_corr * 1e-3, "origins": args.origins,
        "position_GBs": N * (12 if args.f32 else 8 * args.stride) / (per_frame_ms * 1e-3) / 1e9,
        "positions": "float32 xyz (cavb200_rhok_f32)" if args.f32 else f"float64, stride {args.stride}",
        "ms_correlation_all_origins_x_lags": ms_corr, "gpu_launches": h.launch_count - l0,
    }
    if not args.no_cpu:
        from oracle import oracle as O
        n_s = min(N, 200_000)
        t0 = time.perf_counter()
        ref = O.numpy_density_field(frames[0, :n_s, :3], kvec)
        dt = time.perf_counter() - t0
        d_one = capi.DeviceArray.from_numpy(np.ascontiguousarray(frames[0, :n_s, :3]))
        d_r1 = capi.DeviceArray((1, K, 2), np.float64)
        (_rhok if args.f32 else h.rhok)(d_one, 3, n_s * 3, n_s, 1, d_k, K, d_r1, st.ptr)
        got = d_r1.numpy(st.ptr)[0]
        err = np.abs((got[:, 0] + 1j * got[:, 1]) - ref).max() / np.abs(ref).max()
        out["cpu_numpy"] = {"sincos_pairs_per_s": n_s * K / dt, "sample": f"{n_s} particles x {K} k, 1 frame, 1 thread",
                            "max_rel_err_gpu_vs_numpy": float(err)}
    # end to end from a host-resident trajectory (what a GSD file gives): pinned frames -> H2D -> kernel, two streams
    # with a handle each so that one block's upload runs under the other block's kernel
    lib = capi.load()
    src = frames32 if args.f32 else frames
    pin = capi.PinnedArray.from_numpy(src)
    hs = [capi.Handle(local), capi.Handle(local)]
    ss = [capi.Stream(), capi.Stream()]
    dbuf = [capi.DeviceArray(src.shape, src.dtype) for _ in range(2)]
    drho = [capi.DeviceArray((T, K, 2), np.float64) for _ in range(2)]

    def block(k):
        j = k & 1
        lib.cavb200_memcpy_h2d(dbuf[j].ptr, pin.ptr, src.nbytes, ss[j].ptr)
        if args.f32:
            hs[j].rhok_f32(dbuf[j], 3 * N, N, T, d_k, K, drho[j], ss[j].ptr)
        else:
            hs[j].rhok(dbuf[j], args.stride, N * args.stride, N, T, d_k, K, drho[j], ss[j].ptr)

    for k in range(4):
        block(k)
    capi.sync()
    nblk = 16
    t0 = time.perf_counter()
    for k in range(nblk):
        block(k)
    capi.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / (nblk * T)
    out["e2e_from_pinned_host_frames"] = {"ms_per_frame": e2e_ms, "h2d_bytes_per_frame": src.nbytes // T,
                                          "sincos_pairs_per_s": N * K / (e2e_ms * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
