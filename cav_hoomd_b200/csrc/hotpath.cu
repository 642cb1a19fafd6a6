// hotpath.cu -- kernels and launchers of the single-GPU cavity force / Bussi path.
// The three phases (reduce, combine, apply) are device functions in hotpath.cuh.
#include "hotpath.cuh"

#ifndef CAVB_FOLDER_WARMUP
#define CAVB_FOLDER_WARMUP 1
#endif
#ifndef CAVB_FUSED_CHARGE_PREFETCH
#define CAVB_FUSED_CHARGE_PREFETCH 4 // charges per thread fetched before the hand-off of the force-only call
#endif
#ifndef CAVB_FUSED_VEL_PREFETCH
#define CAVB_FUSED_VEL_PREFETCH 4 // velocities per thread fetched before the hand-off of the Bussi-only call
#endif
#ifndef CAVB_VEL_PREFETCH
#define CAVB_VEL_PREFETCH 4 // velocities per thread fetched before the wait for alpha (k_split_folder)
#endif

namespace cavb
    {
// Per-launch device timestamps (tuning "stamps" = 2; bench.py's kernel_ms_device): the upper quarter of the stamps
// buffer is a ring of LAUNCH_RING {first CTA start, last CTA end} pairs in globaltimer ns, indexed by the launch's
// epoch; every CTA contributes one atomicMin and one atomicMax.  cavb200_debug_launch_ring resets / reads it.
// (Both stamps are taken by thread 0 inside short branches so that no value stays live across the kernel: the
// persistent kernels sit exactly at their 80-register budget, and a start time carried to the end cost a spill and
// 0.4-0.6 us per 1M-particle call.)
__device__ __forceinline__ void ring_start(unsigned long long* stamps, unsigned long long epoch, unsigned long long t_start)
    {
    atomicMin(stamps + LAUNCH_RING_OFFSET + 2ull * (epoch & (LAUNCH_RING - 1)), t_start);
    }
__device__ __forceinline__ void ring_end(unsigned long long* stamps, int mode, unsigned long long epoch)
    {
    if (mode == 2 && threadIdx.x == 0)
        atomicMax(stamps + LAUNCH_RING_OFFSET + 2ull * (epoch & (LAUNCH_RING - 1)) + 1, globaltimer_ns());
    }

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// variant 0, kernel 1: reduce; the last CTA to take a ticket combines and publishes Scalars + Final
// (LIST: the thermostatted group is an index list -- Bussi-only instantiations, see hotpath.cuh reduce_ke)
template<bool FORCE, bool BUSSI, int UNROLL, int LB, bool LIST>
__global__ void __launch_bounds__(LB, LB == 384 ? 2 : 1)
    k_reduce(ForceIn f, BussiIn b, Partial* recs, Scalars* scalars, Final* fin_out, unsigned long long* ticket)
    {
    __shared__ BlockScratch sc;
    __shared__ int s_last;
    pdl_wait();
    if (threadIdx.x == 0)
        sc.flags = 0u;
    reduce_phase<FORCE, BUSSI, UNROLL, LIST>(f, b, sc);
    pdl_launch_dependents();
    if (threadIdx.x == 0)
        {
        publish_record(recs + blockIdx.x, sc.rec, 0ull);
        __threadfence();
        const unsigned long long t = atom_acq_rel_add_u64(ticket, 1ull);
        s_last = (t == (unsigned long long)gridDim.x - 1);
        }
    __syncthreads();
    if (!s_last)
        return;
    combine_phase<FORCE, BUSSI, false, true>(recs, (int)gridDim.x, 0ull, f, b, sc, scalars, true);
    if (threadIdx.x == 0)
        {
        *fin_out = sc.fin;
        *ticket = 0ull; // ready for the next launch on this stream
        }
    }

// variant 0, kernel 2
template<bool FORCE, bool BUSSI, int UNROLL, int LB, bool LIST>
__global__ void __launch_bounds__(LB, LB == 384 ? 2 : 1) k_apply(ForceIn f, BussiIn b, const Final* fin_in)
    {
    __shared__ Final fin;
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0)
        fin = *fin_in;
    __syncthreads();
    apply_phase<FORCE, BUSSI, UNROLL, LIST>(fin, f, b);
    }

// variant 1: cooperative persistent kernel (co-residency guaranteed by cudaLaunchCooperativeKernel).
// Hand-off between reduce and apply: the launch's epoch is (last epoch + 1), read from device memory
// by every CTA before it publishes anything and written back by CTA 0 after its combine -- i.e. after
// every CTA has published, hence after every CTA has read it.  No host-side state, so the kernel is
// safe under CUDA-graph replay and for any sequence of grid sizes.
template<bool FORCE, bool BUSSI, int UNROLL, int LB, bool RING, bool LIST>
__global__ void __launch_bounds__(LB, LB == 384 ? 2 : 1)
    k_fused(ForceIn f, BussiIn b, Partial* recs, Scalars* scalars, unsigned long long* epoch_ctr,
            unsigned long long* stamps)
    {
    constexpr int stamp_mode = RING ? 2 : 1;
    __shared__ BlockScratch sc;
    // optional phase stamps (tools/phase_stamps.py): thread 0 of every CTA, globaltimer ns
    unsigned long long* my_stamps = (!RING && stamps && threadIdx.x == 0) ? stamps + 8ull * blockIdx.x : nullptr;
    if (my_stamps)
        my_stamps[0] = globaltimer_ns();
    // programmatic dependent launch: this grid may have been scheduled while the previous kernel of
    // the stream was still draining; nothing below may touch global memory before this returns
    pdl_wait();
    if (threadIdx.x == 0)
        {
        const unsigned long long t_start = stamp_mode == 2 ? globaltimer_ns() : 0ull;
        sc.flags = 0u;
        sc.epoch = ld_relaxed_u64(epoch_ctr) + 1ull;
        if (stamp_mode == 2)
            ring_start(stamps, sc.epoch, t_start);
        }
    reduce_phase<FORCE, BUSSI, UNROLL, LIST>(f, b, sc); // (its barriers also publish sc.epoch to the block)
    if (threadIdx.x == 0)
        publish_record(recs + blockIdx.x, sc.rec, sc.epoch);
    if (my_stamps)
        my_stamps[1] = globaltimer_ns();
    const unsigned long long epoch = sc.epoch;
    // Bussi-only call: the whole grid now waits ~3.5 us for alpha with nothing to do; fetch the first velocities of
    // the rescale pass (L2 hits) before the wait, so that the pass starts with stores (as in k_split_folder)
    constexpr bool PRE = !FORCE && BUSSI;
    // (only while the velocities still sit in L2 after the KE pass: 262k 7.24 -> 6.30 us, 1M 16.66 -> 16.03 us, but
    // 4M 63.1 -> 64.3 us, where the early loads go to HBM)
    const bool pre = PRE && !LIST && b.rescale && b.first == 0 && b.n <= 2000000u;
    VelPrefetch<PRE ? CAVB_FUSED_VEL_PREFETCH : 1> pv;
    if (pre)
        pv = prefetch_vel<PRE ? CAVB_FUSED_VEL_PREFETCH : 1>(b, whole_grid());
    constexpr bool PREF = FORCE && !BUSSI;
    ChargePrefetch<PREF ? CAVB_FUSED_CHARGE_PREFETCH : 1> cpre;
    if (PREF)
        cpre = prefetch_charge<PREF ? CAVB_FUSED_CHARGE_PREFETCH : 1>(f, whole_grid());
    combine_phase<FORCE, BUSSI, true, true>(recs, (int)gridDim.x, epoch, f, b, sc, scalars, blockIdx.x == 0, my_stamps);
    if (my_stamps)
        my_stamps[3] = globaltimer_ns();
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *epoch_ctr = epoch;
    if (sc.fin.timeout)
        {
        if (threadIdx.x == 0)
            raise_fault(scalars); // outputs not written; the host sees it at its next call (api.cu check_fault)
        return;
        }
    // from here on the next kernel's CTAs may take the SMs this grid's CTAs leave
    pdl_launch_dependents();
    if (PREF)
        forces_prefetched<PREF ? CAVB_FUSED_CHARGE_PREFETCH : 1, UNROLL>(sc.fin, f, whole_grid(), cpre);
    else if (pre)
        {
        if (sc.fin.bussi_ok && sc.fin.alpha != 1.0) // as apply_phase
            rescale_prefetched<PRE ? CAVB_FUSED_VEL_PREFETCH : 1, UNROLL>(sc.fin.alpha, b, whole_grid(), pv);
        }
    else
        apply_phase<FORCE, BUSSI, UNROLL, LIST>(sc.fin, f, b);
    if (my_stamps)
        my_stamps[4] = globaltimer_ns();
    ring_end(stamps, stamp_mode, epoch);
    }

// variant 2 ("split-phase"): the thermostat half and the force half are independent, so their two
// hand-offs can hide behind each other's streaming instead of idling the machine:
//     dipole reduce -> publish | KE reduce -> publish | take Dq (arrived long ago) -> forces |
//     take alpha (arrived during the force writes) -> rescale
// No CTA ever waits for the slowest CTA of the phase it has just finished.  Measured background
// (profiles/microwb_r1b.txt): the read half is HBM bound, the write half is bound by L2 store
// ingest with HBM idle, so a CTA that runs ahead into its write half does not slow the readers.
template<int UNROLL, int LB, bool RING>
__global__ void __launch_bounds__(LB, LB == 384 ? 2 : 1)
    k_split(ForceIn f, BussiIn b, Partial* recsF, Partial* recsB, Scalars* scalars, unsigned long long* epoch_ctr,
            unsigned long long* stamps)
    {
    constexpr int stamp_mode = RING ? 2 : 1;
    __shared__ BlockScratch sc;
    unsigned long long* my_stamps = (!RING && stamps && threadIdx.x == 0) ? stamps + 8ull * blockIdx.x : nullptr;
    if (my_stamps)
        my_stamps[0] = globaltimer_ns();
    pdl_wait();
    if (threadIdx.x == 0)
        {
        const unsigned long long t_start = stamp_mode == 2 ? globaltimer_ns() : 0ull;
        sc.flags = 0u;
        sc.epoch = ld_relaxed_u64(epoch_ctr) + 1ull;
        if (stamp_mode == 2)
            ring_start(stamps, sc.epoch, t_start);
        }
    // dipole first: its hand-off (the heavier combine, and the one whose stragglers matter) then has
    // the whole velocity pass to complete behind; the light KE hand-off hides behind the force writes
    reduce_phase<true, false, UNROLL>(f, b, sc);
    const unsigned long long epoch = sc.epoch;
    if (threadIdx.x == 0)
        publish_record(recsF + blockIdx.x, sc.rec, epoch);
    if (my_stamps)
        my_stamps[1] = globaltimer_ns();
    reduce_phase<false, true, UNROLL>(f, b, sc);
    if (threadIdx.x == 0)
        publish_record(recsB + blockIdx.x, sc.rec, epoch);
    if (my_stamps)
        my_stamps[2] = globaltimer_ns();
    // (debug stamps of the combine's inner steps go to a second bank, 1024 CTAs further on)
    combine_phase<true, false, true, true>(recsF, (int)gridDim.x, epoch, f, b, sc, scalars, blockIdx.x == 0,
                                           my_stamps ? my_stamps + 8 * 1024 : nullptr);
    if (my_stamps)
        my_stamps[3] = globaltimer_ns();
    const bool timeout_f = sc.fin.timeout != 0;
    if (!timeout_f)
        apply_phase<true, false, UNROLL>(sc.fin, f, b);
    if (my_stamps)
        my_stamps[4] = globaltimer_ns();
    __syncthreads(); // sc.fin / sc.rec are rewritten by the second combine
    combine_phase<false, true, true, true>(recsB, (int)gridDim.x, epoch, f, b, sc, scalars, blockIdx.x == 0);
    if (my_stamps)
        my_stamps[5] = globaltimer_ns();
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *epoch_ctr = epoch; // every CTA has published twice, hence read the counter
    if (sc.fin.timeout || timeout_f)
        {
        if (threadIdx.x == 0)
            raise_fault(scalars);
        return;
        }
    pdl_launch_dependents();
    apply_phase<false, true, UNROLL>(sc.fin, f, b);
    if (my_stamps)
        my_stamps[6] = globaltimer_ns();
    ring_end(stamps, stamp_mode, epoch);
    }

// variant 3 ("split-phase with a folder CTA"): the split-phase schedule, but ONE extra CTA that streams
// nothing folds the per-CTA records (it waits for them, so it runs behind the stragglers by itself) and
// publishes the Final record; the streaming CTAs only poll that record.  In k_split every CTA spends
// 2.75 + 2.0 us per step folding all records (profiles/split_r1e.txt); here that is two L2 round trips.
//     streaming CTA:  dipole reduce -> publish | KE reduce -> publish | take Dq -> forces | take alpha -> rescale
//     folder CTA:     fold dipole records -> publish Final(F) | fold KE records -> publish Final(K)
// The dipole is the correctly rounded sum of the same terms as in the other variants, folded over one
// record fewer (grid - 1 streaming CTAs): results agree to the last bit or two, not bit for bit.
template<int UNROLL, int LB, bool KE_FIRST, bool RING>
__global__ void __launch_bounds__(LB, LB == 384 ? 2 : 1)
    k_split_folder(ForceIn f, BussiIn b, Partial* recsF, Partial* recsB, Partial* finals, Scalars* scalars,
                   unsigned long long* epoch_ctr, unsigned long long* stamps)
    {
    constexpr int stamp_mode = RING ? 2 : 1;
    __shared__ BlockScratch sc;
    unsigned long long* my_stamps = (!RING && stamps && threadIdx.x == 0) ? stamps + 8ull * blockIdx.x : nullptr;
    if (my_stamps)
        my_stamps[0] = globaltimer_ns();
    pdl_wait();
    if (threadIdx.x == 0)
        {
        const unsigned long long t_start = stamp_mode == 2 ? globaltimer_ns() : 0ull;
        sc.flags = 0u;
        sc.epoch = ld_relaxed_u64(epoch_ctr) + 1ull;
        if (stamp_mode == 2)
            ring_start(stamps, sc.epoch, t_start);
        }
    StreamGrid g;
    g.nblk = gridDim.x - 1;
    g.blk = blockIdx.x - 1; // CTA 0 is the folder: it is resident first, long before it is needed
    if (blockIdx.x == 0)
        {
        // ---- folder ----
        // The fold runs ONCE per launch, on one SM, while the rest of the grid streams: its instructions
        // are cold (evicted from L2 by the 148 MB a step moves) and every instruction-cache miss costs a
        // loaded memory round trip -- measured: 6 us for a fold that takes 2.75 us hot.  So the folder
        // first runs the very same code (same loop body, pass 0) over a bank of dummy records it wrote
        // itself, while it has nothing to do anyway, and then (pass 1) over the real records.
        __syncthreads();
        const unsigned long long epoch = sc.epoch, dummy_epoch = ~epoch;
        Partial* dummy = recsF + MAX_PARTIALS / 4; // unused part of the force bank (grid <= MAX_PARTIALS / 4)
        for (int j = threadIdx.x; j < (int)g.nblk; j += blockDim.x)
            {
            Partial z = {};
            z.first_L = ~0ull;
            z.ke = 1.0;
            if (j == 0)
                z.first_L = 0ull; // one candidate, so that the photon path is warmed up too
            publish_record(dummy + j, z, dummy_epoch);
            }
        __syncthreads();
        bool timeout_f = false;
#pragma unroll 1
        for (int pass = CAVB_FOLDER_WARMUP ? 0 : 1; pass < 2; pass++)
            {
            const bool real = pass == 1;
            const unsigned long long ep = real ? epoch : dummy_epoch;
            if (my_stamps && real)
                my_stamps[1] = globaltimer_ns();
#pragma unroll
            for (int half = 0; half < 2; half++)
                {
                if ((half == 0) != KE_FIRST)
                    {
                    combine_phase<true, false, true, true, false, true>(real ? recsF : dummy, (int)g.nblk, ep, f, b, sc, scalars,
                                                                        real, (my_stamps && real) ? my_stamps + 8 * 1024 : nullptr);
                    timeout_f = timeout_f || sc.fin.timeout != 0;
                    if (threadIdx.x == 0)
                        {
                        if (timeout_f)
                            sc.fin.timeout = 1;
                        publish_final<true>(real ? finals + 0 : dummy + g.nblk, sc.fin, ep);
                        }
                    if (my_stamps && real)
                        my_stamps[3] = globaltimer_ns();
                    }
                else
                    {
                    combine_phase<false, true, true, true, false, true>(real ? recsB : dummy, (int)g.nblk, ep, f, b, sc, scalars,
                                                                        real);
                    timeout_f = timeout_f || sc.fin.timeout != 0;
                    if (threadIdx.x == 0)
                        {
                        if (timeout_f)
                            sc.fin.timeout = 1;
                        publish_final<false>(real ? finals + 1 : dummy + g.nblk + 1, sc.fin, ep);
                        }
                    if (my_stamps && real)
                        my_stamps[5] = globaltimer_ns();
                    }
                __syncthreads();
                }
            if (threadIdx.x == 0 && real)
                *epoch_ctr = epoch; // every streaming CTA has published twice, hence read the counter
            }
        pdl_launch_dependents();
        ring_end(stamps, stamp_mode, epoch);
        return;
        }
    // ---- streaming CTAs ----
    // first half: the one whose hand-off has the whole second pass to hide behind; second half: its hand-off
    // hides behind the first half's write pass.  KE_FIRST = thermostat first (see the launcher).
    if (!KE_FIRST)
        reduce_phase<true, false, UNROLL>(f, b, sc, g);
    else
        reduce_phase<false, true, UNROLL>(f, b, sc, g);
    const unsigned long long epoch = sc.epoch;
    if (threadIdx.x == 0)
        publish_record((KE_FIRST ? recsB : recsF) + g.blk, sc.rec, epoch);
    if (my_stamps)
        my_stamps[1] = globaltimer_ns();
    FinalSectors pre;
        {
        Acc a;
        // the first half's Final has normally been out since the middle of this pass: fetch it under the merge tree
        if (!KE_FIRST)
            {
            reduce_loops<false, true, UNROLL>(a, f, b, g);
            pre = prefetch_final<true>(finals + 0);
            block_merge<false, true>(a, f, sc);
            }
        else
            {
            reduce_loops<true, false, UNROLL>(a, f, b, g);
            pre = prefetch_final<false>(finals + 1);
            block_merge<true, false>(a, f, sc);
            }
        }
    if (threadIdx.x == 0)
        publish_record((KE_FIRST ? recsF : recsB) + g.blk, sc.rec, epoch);
    if (my_stamps)
        my_stamps[2] = globaltimer_ns();
    if (!KE_FIRST)
        {
        const Final finF = take_final<true>(finals + 0, epoch, pre);
        if (my_stamps)
            my_stamps[3] = globaltimer_ns();
        if (!finF.timeout)
            apply_phase<true, false, UNROLL>(finF, f, b, g);
        if (my_stamps)
            my_stamps[4] = globaltimer_ns();
        const FinalSectors preK = prefetch_final<false>(finals + 1);
        const bool contiguous = b.first == 0; // (the step takes a window of the index space, never a list)
        VelPrefetch<CAVB_VEL_PREFETCH> pv;
        if (contiguous)
            pv = prefetch_vel<CAVB_VEL_PREFETCH>(b, g);
        const Final finK = take_final<false>(finals + 1, epoch, preK);
        if (my_stamps)
            my_stamps[5] = globaltimer_ns();
        if (finF.timeout || finK.timeout)
            {
            if (threadIdx.x == 0)
                raise_fault(scalars);
            return;
            }
        pdl_launch_dependents();
        if (contiguous)
            {
            if (b.rescale && finK.bussi_ok && finK.alpha != 1.0) // as apply_phase
                rescale_prefetched<CAVB_VEL_PREFETCH, UNROLL>(finK.alpha, b, g, pv);
            }
        else
            apply_phase<false, true, UNROLL>(finK, f, b, g);
        }
    else
        {
        const Final finK = take_final<false>(finals + 1, epoch, pre);
        if (my_stamps)
            my_stamps[3] = globaltimer_ns();
        if (!finK.timeout)
            apply_phase<false, true, UNROLL>(finK, f, b, g);
        if (my_stamps)
            my_stamps[4] = globaltimer_ns();
        const Final finF = take_final<true>(finals + 0, epoch, prefetch_final<true>(finals + 0));
        if (my_stamps)
            my_stamps[5] = globaltimer_ns();
        if (finF.timeout || finK.timeout)
            {
            if (threadIdx.x == 0)
                raise_fault(scalars);
            return;
            }
        pdl_launch_dependents();
        apply_phase<true, false, UNROLL>(finF, f, b, g);
        }
    if (my_stamps)
        my_stamps[6] = globaltimer_ns();
    ring_end(stamps, stamp_mode, epoch);
    }

// Small systems (tuning "small_n", default 768 particles; the reference's own example runs 501): ONE CTA does the whole
// call -- reduce, block merge, finalize, apply -- so there is no hand-off between CTAs at all.  At 501 particles the
// two-CTA persistent kernels take 5.5 / 3.7 / 8.7 us (force / Bussi / step), nearly all of it the global-memory hand-off;
// this one takes 3.5 / 3.4 / 5.3 us and is bound by its chain of latencies (profiles/small_n_r2a.txt: it beats the
// persistent kernels below ~1000 particles and the single-cluster kernel below ~700).  Same per-thread orders and the same merge trees for the step as for the two calls, so cavb200_step stays
// bit-identical to cavb200_force + cavb200_bussi.
template<bool FORCE, bool BUSSI, bool LIST>
__global__ void __launch_bounds__(1024, 1) k_small(ForceIn f, BussiIn b, Scalars* scalars, Final* fin_out)
    {
    __shared__ BlockScratch sc;
    pdl_wait();
    if (threadIdx.x == 0)
        sc.flags = 0u;
    StreamGrid g;
    g.nblk = 1;
    g.blk = 0;
    reduce_phase<FORCE, BUSSI, 2, LIST>(f, b, sc, g);
    if (threadIdx.x == 0)
        {
        finalize<FORCE, BUSSI>(f, b, sc, scalars, true);
        if (FORCE && !BUSSI && f.force == nullptr)
            *fin_out = sc.fin; // rank-1 mode: the consumers read Dq / F_L from here
        }
    __syncthreads();
    pdl_launch_dependents();
    if (FORCE && !BUSSI && f.force == nullptr)
        return;
    apply_phase<FORCE, BUSSI, 2, LIST>(sc.fin, f, b, g);
    }

// Systems between small_n and cluster_n particles: ONE thread-block cluster does the whole call.  The CTAs of a cluster
// are co-scheduled by the hardware, can store into each other's shared memory and share a hardware barrier, so the hand-off
// needs neither global memory nor polling nor epochs nor a time-out: every CTA stores its record into every CTA's inbox
// (distributed shared memory), barrier.cluster (release / acquire), every CTA folds the <= 16 records for itself -- the same
// fold, in the same order, as the persistent kernels' -- and applies.  Measured against the persistent kernels
// (profiles/cluster_n_r2a.txt, cluster of 16): the one-launch step 5.75 against 8.1 us at 1.5k-4k particles, 6.55 against
// 8.4 us at 8k, equal at 16k; the force call 4.1 against 5.35 us (4.7 against 5.5 at 8k, slower at 16k); force + Bussi 7.8
// against 9.0 us.  (With the records going through L2 instead of shared memory, as first built: step 6.4, force 4.9 us.)
// The Bussi-only call is no better at any size -- so: calls with the force in them, up to tuning cluster_n (8192) particles.
template<bool FORCE, bool BUSSI, bool LIST>
__global__ void __launch_bounds__(1024, 1) k_cluster(ForceIn f, BussiIn b, Scalars* scalars, Final* fin_out)
    {
    __shared__ BlockScratch sc;
    __shared__ Partial inbox[16];
    // (a CTA's shared memory may be written by its peers only once it runs: first half of a barrier now, second half
    // before the first remote store)
    cluster_arrive();
    pdl_wait();
    if (threadIdx.x == 0)
        sc.flags = 0u;
    StreamGrid g;
    g.nblk = cluster_nctarank();
    g.blk = cluster_ctarank();
    reduce_phase<FORCE, BUSSI, 2, LIST>(f, b, sc, g);
    cluster_wait();
    cluster_post(inbox, sc.rec, g.blk, g.nblk);
    cluster_sync();
    combine_cluster<FORCE, BUSSI>(inbox, g.nblk, f, b, sc, scalars, g.blk == 0);
    if (FORCE && !BUSSI && f.force == nullptr)
        {
        if (g.blk == 0 && threadIdx.x == 0)
            *fin_out = sc.fin; // rank-1 mode: the consumers read Dq / F_L from here
        pdl_launch_dependents();
        return;
        }
    pdl_launch_dependents();
    apply_phase<FORCE, BUSSI, 2, LIST>(sc.fin, f, b, g);
    }

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// LB = launch bound: 1024 caps the kernel at 64 registers/thread (two 512-thread CTAs or one
// 1024-thread CTA per SM); 512 allows 128 registers for one 512-thread CTA per SM.
// `shape_threads` / `shape_ctas`: threads per CTA and CTAs per SM of the persistent grids (the tuning knobs, or what
// launch_u's rule made of them)
template<bool FORCE, bool BUSSI, int UNROLL, int LB, bool LIST = false>
static int launch_t(cavb200_handle* h, const ForceIn& f, const BussiIn& b, cudaStream_t s, int shape_threads, int shape_ctas)
    {
    if (BUSSI && !LIST && b.gidx != nullptr)
        {
        // an index list: only the Bussi-only calls take one (cavb200_bussi, cavb200_bussi_ke)
        if (FORCE)
            return (int)cudaErrorInvalidValue;
        return launch_t<FORCE, BUSSI && !FORCE, UNROLL, LB, !FORCE>(h, f, b, s, shape_threads, shape_ctas);
        }
        {
        const unsigned long long work_small = (FORCE ? (unsigned long long)f.N : 0ull) > (BUSSI ? (unsigned long long)b.n : 0ull)
                                                  ? (unsigned long long)f.N
                                                  : (unsigned long long)b.n;
        if (h->tune.small_n > 0 && work_small <= (unsigned long long)h->tune.small_n)
            {
            // one CTA, as many threads as there are particles (rounded up to a warp, at most 1024)
            int t = (int)((work_small + 31) / 32 * 32);
            t = t < 64 ? 64 : (t > 1024 ? 1024 : t);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1);
            cfg.blockDim = dim3(t);
            cfg.stream = s;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr.val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = h->tune.pdl ? 1 : 0;
            ForceIn ff = f;
            BussiIn bb = b;
            Scalars* sca = h->scalars;
            Final* fo = const_cast<Final*>(rank1_final(h));
            void* args[] = {&ff, &bb, &sca, &fo};
            CAVB_CHECK(cudaLaunchKernelExC(&cfg, (const void*)k_small<FORCE, BUSSI, LIST>, args));
            h->launches += 1;
            return 0;
            }
        // (the Bussi-only call gains nothing from it: 3.7-3.9 us either way up to 4k particles, slower above; profiles/cluster_n_r2a.txt)
        if (FORCE && h->tune.cluster_n > 0 && h->tune.cluster_ctas >= 2)
            {
            // one cluster of 16 CTAs (a non-portable size: allowed per kernel at handle creation, cluster_kernels_init; if the
            // device cannot place it, 8 from then on, which pays off for half as many particles), as many threads per CTA
            // as spread the particles one per thread
            const void* kern = (const void*)k_cluster<FORCE, BUSSI, LIST>;
            for (int attempt = 0; attempt < 2; attempt++)
                {
                const int cmax = h->tune.cluster_ctas >= 16 ? 16 : (h->tune.cluster_ctas >= 8 ? 8 : (h->tune.cluster_ctas >= 4 ? 4 : 2));
                if (work_small > (unsigned long long)h->tune.cluster_n * (unsigned long long)cmax / 16ull)
                    break;
                // (up to 2k particles 8 CTAs are as fast as 16, and steadier: 1001 particles, step 5.5 against 6.6 us)
                const int ctas = (cmax == 16 && work_small <= 2048ull) ? 8 : cmax;
                int t = (int)((work_small + ctas - 1) / ctas);
                t = (t + 31) / 32 * 32;
                t = t < 64 ? 64 : (t > 1024 ? 1024 : t);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(ctas);
                cfg.blockDim = dim3(t);
                cfg.stream = s;
                cudaLaunchAttribute attrs[2];
                attrs[0].id = cudaLaunchAttributeClusterDimension;
                attrs[0].val.clusterDim.x = ctas;
                attrs[0].val.clusterDim.y = 1;
                attrs[0].val.clusterDim.z = 1;
                attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attrs[1].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attrs;
                cfg.numAttrs = h->tune.pdl ? 2 : 1;
                ForceIn ff = f;
                BussiIn bb = b;
                Scalars* sca = h->scalars;
                Final* fo = const_cast<Final*>(rank1_final(h));
                void* args[] = {&ff, &bb, &sca, &fo};
                const cudaError_t e = cudaLaunchKernelExC(&cfg, kern, args);
                if (e == cudaSuccess)
                    {
                    h->launches += 1;
                    return 0;
                    }
                cudaGetLastError();
                if (ctas != 16)
                    return (int)e;
                h->tune.cluster_ctas = 8; // this device (or this share of it) has no room for 16 co-scheduled CTAs
                }
            }
        }
    int threads = shape_threads;
    const unsigned long long work = (FORCE ? (unsigned long long)f.N : 0ull) > (BUSSI ? (unsigned long long)b.n : 0ull)
                                        ? (unsigned long long)f.N
                                        : (unsigned long long)b.n;
    // (the Bussi-only call follows the same rule -- 1M: 15.9 us at 384 threads, 15.6 at 352, profiles/threads_r2a.txt;
    // the force-only call is flat in the CTA size)
    if (BUSSI && UNROLL == 2 && LB == 384 && threads == 384 && h->tune.auto_threads && shape_ctas == 2)
        {
        // Step kernel, default shape: a thread walks ceil(q) particles (q = work / streaming threads) two at a time, so
        // each of its four passes is ceil(ceil(q) / 2) dependent memory round trips.  Among 320 / 352 / 384 threads per
        // CTA take the fewest round trips, and on a tie the fewest threads (measured, tools/threads_sweep.py,
        // profiles/threads_r1a.txt: 1M particles 31.09 us at 384 -> 30.57 us at 352; 2M and more stay at 384; 262k and
        // 500k go to 320).
        const unsigned long long ctas = (unsigned long long)h->num_sms * 2 - ((FORCE && h->tune.variant == 3) ? 1 : 0);
        unsigned long long best_rounds = ~0ull;
        for (int t = 320; t <= 384; t += 32)
            {
            const unsigned long long per_thread = (work + ctas * t - 1) / (ctas * t);
            const unsigned long long rounds = (per_thread + 1) / 2;
            if (rounds < best_rounds)
                {
                best_rounds = rounds;
                threads = t;
                }
            }
        }
    unsigned long long want = (work + threads - 1) / threads;
    if (want < 1)
        want = 1;
    int max_grid = h->num_sms * shape_ctas;
    if (max_grid > MAX_PARTIALS)
        max_grid = MAX_PARTIALS;
    Final* fin_dev = reinterpret_cast<Final*>(h->counters + 8);

    if (FORCE && BUSSI && (h->tune.variant == 2 || h->tune.variant == 3) && h->coop_supported && b.rescale)
        {
        // the folder pays two extra L2 round trips, which only a full streaming grid hides (65k particles:
        // 11.95 us with the folder, 10.1 us without): small systems take the plain split-phase kernel
        const bool folder = h->tune.variant == 3 && want >= (unsigned long long)max_grid;
        // (the per-launch timestamp ring, tuning stamps = 2, is a separate instantiation: the kernels sit exactly at their
        // 80-register budget and even a runtime switch for it cost spills and ~0.5 us per 1M-particle call)
        const bool ring = h->tune.stamps == 2;
        const void* kern;
        int per_sm = 0;
        if (folder && h->tune.ke_first)
            {
            kern = ring ? (const void*)k_split_folder<UNROLL, LB, true, true> : (const void*)k_split_folder<UNROLL, LB, true, false>;
            CAVB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_split_folder<UNROLL, LB, true, false>, threads, 0));
            }
        else if (folder)
            {
            kern = ring ? (const void*)k_split_folder<UNROLL, LB, false, true> : (const void*)k_split_folder<UNROLL, LB, false, false>;
            CAVB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_split_folder<UNROLL, LB, false, false>, threads, 0));
            }
        else
            {
            kern = ring ? (const void*)k_split<UNROLL, LB, true> : (const void*)k_split<UNROLL, LB, false>;
            CAVB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_split<UNROLL, LB, false>, threads, 0));
            }
        if (per_sm < 1)
            return (int)cudaErrorLaunchOutOfResources;
        if (max_grid > per_sm * h->num_sms)
            max_grid = per_sm * h->num_sms;
        if (max_grid > MAX_PARTIALS / 4 - 2)
            max_grid = MAX_PARTIALS / 4 - 2;
        int grid = (int)(want < (unsigned long long)max_grid ? want : (unsigned long long)max_grid);
        if (folder) // one CTA of the co-resident grid folds instead of streaming
            grid = (grid < max_grid ? grid : max_grid - 1) + 1;
        if (grid < (folder ? 2 : 1))
            return (int)cudaErrorLaunchOutOfResources;
        Partial* recsF = h->partials;
        Partial* recsB = h->partials + MAX_PARTIALS / 2;
        Partial* finals = h->partials + MAX_PARTIALS - 2; // the last two records of the thermostat half's bank
        Scalars* sca = h->scalars;
        unsigned long long* arr = h->counters + 2;
        ForceIn ff = f;
        BussiIn bb = b;
        unsigned long long* stamps = h->tune.stamps ? h->stamps : nullptr;
        void* args_split[] = {&ff, &bb, &recsF, &recsB, &sca, &arr, &stamps};
        void* args_folder[] = {&ff, &bb, &recsF, &recsB, &finals, &sca, &arr, &stamps};
        void** args = folder ? args_folder : args_split;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.stream = s;
        cudaLaunchAttribute attrs[1];
        if (h->tune.pdl)
            {
            attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[0].val.programmaticStreamSerializationAllowed = 1;
            }
        else
            {
            attrs[0].id = cudaLaunchAttributeCooperative;
            attrs[0].val.cooperative = 1;
            }
        cfg.attrs = attrs;
        cfg.numAttrs = 1;
        CAVB_CHECK(cudaLaunchKernelExC(&cfg, kern, args));
        h->launches += 1;
        return 0;
        }

    // rank-1 mode (cavb200_force_rank1): no force array -- dipole reduce, fold, Scalars + Final
    const bool rank1 = FORCE && !BUSSI && f.force == nullptr;
    if (rank1)
        fin_dev = const_cast<Final*>(rank1_final(h));
    if (rank1 || (!FORCE && BUSSI && !b.rescale && h->tune.variant >= 1))
        {
        // kinetic energy only (cavb200_bussi_ke, what getRescalingFactorsOne needs when HOOMD applies
        // the rescale itself): nothing is applied afterwards, so no CTA has to wait for the others --
        // reduce kernel alone, the last CTA to take a ticket folds the records and publishes Scalars
        const int grid = (int)(want < (unsigned long long)max_grid ? want : (unsigned long long)max_grid);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = h->tune.pdl ? 1 : 0;
        ForceIn ff = f;
        BussiIn bb = b;
        Partial* recs = h->partials;
        Scalars* sca = h->scalars;
        unsigned long long* ticket = h->counters + 4;
        void* a1[] = {&ff, &bb, &recs, &sca, &fin_dev, &ticket};
        CAVB_CHECK(cudaLaunchKernelExC(&cfg, (const void*)k_reduce<FORCE, BUSSI, UNROLL, LB, LIST>, a1));
        h->launches += 1;
        return 0;
        }

    if (h->tune.variant >= 1 && h->coop_supported)
        {
        int per_sm = 0;
        const bool ring = h->tune.stamps == 2;
        CAVB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused<FORCE, BUSSI, UNROLL, LB, false, LIST>, threads, 0));
        if (per_sm < 1)
            return (int)cudaErrorLaunchOutOfResources;
        int resident = per_sm * h->num_sms;
        if (max_grid > resident)
            max_grid = resident;
        const int grid = (int)(want < (unsigned long long)max_grid ? want : (unsigned long long)max_grid);
        Partial* recs = h->partials;
        Scalars* sca = h->scalars;
        unsigned long long* arr = h->counters + 2;
        ForceIn ff = f;
        BussiIn bb = b;
        unsigned long long* stamps = h->tune.stamps ? h->stamps : nullptr;
        void* args[] = {&ff, &bb, &recs, &sca, &arr, &stamps};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attrs[2];
        int na = 0;
        if (h->tune.pdl)
            {
            // PDL and the cooperative attribute are mutually exclusive; the grid is sized from the
            // occupancy query above, so all CTAs are co-resident on an otherwise idle device, and the
            // hand-off spin is bounded (HANDOFF_TIMEOUT_NS) if they ever are not
            attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[na].val.programmaticStreamSerializationAllowed = 1;
            na++;
            }
        else
            {
            attrs[na].id = cudaLaunchAttributeCooperative;
            attrs[na].val.cooperative = 1;
            na++;
            }
        cfg.attrs = attrs;
        cfg.numAttrs = na;
        CAVB_CHECK(cudaLaunchKernelExC(&cfg, ring ? (const void*)k_fused<FORCE, BUSSI, UNROLL, LB, true, LIST>
                                                  : (const void*)k_fused<FORCE, BUSSI, UNROLL, LB, false, LIST>, args));
        h->launches += 1;
        return 0;
        }

    const int grid = (int)(want < (unsigned long long)max_grid ? want : (unsigned long long)max_grid);
        {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = h->tune.pdl ? 1 : 0;
        ForceIn ff = f;
        BussiIn bb = b;
        Partial* recs = h->partials;
        Scalars* sca = h->scalars;
        unsigned long long* ticket = h->counters + 4;
        void* a1[] = {&ff, &bb, &recs, &sca, &fin_dev, &ticket};
        CAVB_CHECK(cudaLaunchKernelExC(&cfg, (const void*)k_reduce<FORCE, BUSSI, UNROLL, LB, LIST>, a1));
        const Final* fin_c = fin_dev;
        void* a2[] = {&ff, &bb, &fin_c};
        CAVB_CHECK(cudaLaunchKernelExC(&cfg, (const void*)k_apply<FORCE, BUSSI, UNROLL, LB, LIST>, a2));
        }
    h->launches += 2;
    return 0;
    }

// Launch shapes: the unroll factor fixes the register budget (launch bound).  More loads in flight
// per thread (ILP) with FEWER threads keeps the same bytes in flight per SM while shrinking the
// block-wide trees, which are FP64-throughput bound when 32 warps per SM run them:
//   unroll 2 -> up to  384 threads/CTA,  80 registers (two CTAs per SM)  [default]
//               up to 1024 threads/CTA,  64 registers
//   unroll 4 -> up to  512 threads/CTA, 128 registers (one CTA per SM)
//   unroll 8 -> up to  256 threads/CTA, 255 registers (one CTA per SM)
template<bool FORCE, bool BUSSI>
static int launch_u(cavb200_handle* h, const ForceIn& f, const BussiIn& b, cudaStream_t s)
    {
    switch (h->tune.unroll)
        {
    case 8:
        if (h->tune.threads > 256)
            return (int)cudaErrorInvalidConfiguration;
        return launch_t<FORCE, BUSSI, 8, 256>(h, f, b, s, h->tune.threads, h->tune.ctas_per_sm);
    case 4:
        if (h->tune.threads > 512)
            return (int)cudaErrorInvalidConfiguration;
        return launch_t<FORCE, BUSSI, 4, 512>(h, f, b, s, h->tune.threads, h->tune.ctas_per_sm);
    default:
        {
        // two 384-thread CTAs per SM (24 warps, 80 registers): the shape that keeps the fused
        // force + KE loop out of local memory (profiles/microprod_r1b.txt); wider CTAs fall back
        // to the 64-register build
        //
        // Below ~2M particles ONE CTA per SM is the better shape for some calls -- half as many records to fold, and the
        // hand-off is a larger share of a short call (profiles/shapes_r2b.txt, us per call, default -> this rule):
        //   force call   65k 8.1 -> 7.0    131k 10.0 -> 8.3   262k 11.8 -> 10.6   524k 15.2 -> 14.2   1M 20.8 -> 20.3   2M+ equal
        //   Bussi call   16k 5.3 -> 4.0    65k 5.1 -> 4.1     131k 5.2 -> 4.8     262k 6.3 -> 6.2     400k+ no better or worse
        //   step         65k 9.6 -> 9.3    131k 13.1 -> 11.3  262k 14.4 -> 13.5   524k+ worse (the folder kernel wants two CTAs per SM)
        // Taken only while the shape knobs are at their defaults; an explicit `threads` / `ctas_per_sm` is obeyed.
        int threads = h->tune.threads, ctas = h->tune.ctas_per_sm;
        if (threads == 384 && ctas == 2 && h->tune.auto_threads)
            {
            const unsigned long long work = (FORCE ? (unsigned long long)f.N : 0ull) > (BUSSI ? (unsigned long long)b.n : 0ull)
                                                ? (unsigned long long)f.N
                                                : (unsigned long long)b.n;
            int t = 0;
            if (FORCE && !BUSSI)
                t = (work >= 32768ull && work < 400000ull) ? 704 : ((work >= 400000ull && work < 2000000ull) ? 768 : 0);
            else if (!FORCE)
                t = work < 300000ull ? 512 : 0;
            else
                t = (work >= 49152ull && work < 400000ull) ? 512 : 0;
            if (t)
                {
                threads = t;
                ctas = 1;
                }
            }
        if (threads <= 384)
            return launch_t<FORCE, BUSSI, 2, 384>(h, f, b, s, threads, ctas);
        if (threads <= 768) // one 768-thread CTA per SM: same 80-register budget, half the records
            return launch_t<FORCE, BUSSI, 2, 768>(h, f, b, s, threads, ctas);
        return launch_t<FORCE, BUSSI, 2, 1024>(h, f, b, s, threads, ctas);
        }
        }
    }

// Called once per handle (cavb200_create), outside any stream capture: allow the cluster kernels the non-portable
// cluster size of 16.  Returns the largest cluster the launcher may ask for.
int cluster_kernels_init()
    {
    const void* kerns[] = {(const void*)k_cluster<true, false, false>, (const void*)k_cluster<true, true, false>};
    for (const void* k : kerns)
        if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
            {
            cudaGetLastError();
            return 8;
            }
    return 16;
    }

int launch_hotpath(cavb200_handle* h, const ForceIn* f, const BussiIn* b, cudaStream_t s)
    {
    ForceIn fz = {};
    BussiIn bz = {};
    if (f && b)
        return launch_u<true, true>(h, *f, *b, s);
    if (f)
        return launch_u<true, false>(h, *f, bz, s);
    if (b)
        return launch_u<false, true>(h, fz, *b, s);
    return 0;
    }
    } // namespace cavb
