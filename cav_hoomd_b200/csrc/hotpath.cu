// hotpath.cu -- kernels and launchers of the single-GPU cavity force / Bussi path.
// The three phases (reduce, combine, apply) are device functions in hotpath.cuh.
#include "hotpath.cuh"

namespace cavb
    {
// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// variant 0, kernel 1: reduce; the last CTA to take a ticket combines and publishes Scalars + Final
template<bool FORCE, bool BUSSI, int UNROLL, int LB>
__global__ void __launch_bounds__(LB, 1)
    k_reduce(ForceIn f, BussiIn b, Partial* recs, Scalars* scalars, Final* fin_out, unsigned long long* ticket)
    {
    __shared__ BlockScratch sc;
    __shared__ int s_last;
    reduce_phase<FORCE, BUSSI, UNROLL>(f, b, sc);
    store_record(recs + blockIdx.x, sc.rec);
    __syncthreads();
    if (threadIdx.x == 0)
        {
        __threadfence();
        const unsigned long long t = atom_acq_rel_add_u64(ticket, 1ull);
        s_last = (t == (unsigned long long)gridDim.x - 1);
        }
    __syncthreads();
    if (!s_last)
        return;
    combine_phase<FORCE, BUSSI>(recs, (int)gridDim.x, f, b, sc, scalars, true, 0);
    if (threadIdx.x == 0)
        {
        *fin_out = sc.fin;
        *ticket = 0ull; // ready for the next launch on this stream
        }
    }

// variant 0, kernel 2
template<bool FORCE, bool BUSSI, int UNROLL, int LB>
__global__ void __launch_bounds__(LB, 1) k_apply(ForceIn f, BussiIn b, const Final* fin_in)
    {
    __shared__ Final fin;
    if (threadIdx.x == 0)
        fin = *fin_in;
    __syncthreads();
    apply_phase<FORCE, BUSSI, UNROLL>(fin, f, b);
    }

// variant 1: cooperative persistent kernel (co-residency guaranteed by cudaLaunchCooperativeKernel)
template<bool FORCE, bool BUSSI, int UNROLL, int LB>
__global__ void __launch_bounds__(LB, 1)
    k_fused(ForceIn f, BussiIn b, Partial* recs, Scalars* scalars, unsigned long long* arrivals)
    {
    __shared__ BlockScratch sc;
    __shared__ int s_timeout;
    reduce_phase<FORCE, BUSSI, UNROLL>(f, b, sc);
    store_record(recs + blockIdx.x, sc.rec);
    __syncthreads();
    if (threadIdx.x == 0)
        {
        __threadfence();
        // arrive / spin / depart.  The last CTA to depart zeroes both counters, so the next launch on
        // this handle (stream-ordered, any grid size, also under CUDA-graph replay) starts clean.
        atom_acq_rel_add_u64(arrivals, 1ull);
        int timeout = 0;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_u64(arrivals) < (unsigned long long)gridDim.x)
            {
            if (globaltimer_ns() - t0 > 2000000000ull) // 2 s: never hang the GPU
                {
                timeout = 1;
                break;
                }
            }
        if (atom_acq_rel_add_u64(arrivals + 1, 1ull) == (unsigned long long)gridDim.x - 1ull)
            {
            arrivals[0] = 0ull;
            arrivals[1] = 0ull;
            }
        s_timeout = timeout;
        }
    __syncthreads();
    combine_phase<FORCE, BUSSI>(recs, (int)gridDim.x, f, b, sc, scalars, blockIdx.x == 0, s_timeout);
    if (s_timeout)
        return;
    apply_phase<FORCE, BUSSI, UNROLL>(sc.fin, f, b);
    }

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// LB = launch bound: 1024 caps the kernel at 64 registers/thread (two 512-thread CTAs or one
// 1024-thread CTA per SM); 512 allows 128 registers for one 512-thread CTA per SM.
template<bool FORCE, bool BUSSI, int UNROLL, int LB>
static int launch_t(cavb200_handle* h, const ForceIn& f, const BussiIn& b, cudaStream_t s)
    {
    const int threads = h->tune.threads;
    const unsigned long long work = (FORCE ? (unsigned long long)f.N : 0ull) > (BUSSI ? (unsigned long long)b.n : 0ull)
                                        ? (unsigned long long)f.N
                                        : (unsigned long long)b.n;
    unsigned long long want = (work + threads - 1) / threads;
    if (want < 1)
        want = 1;
    int max_grid = h->num_sms * h->tune.ctas_per_sm;
    if (max_grid > MAX_PARTIALS)
        max_grid = MAX_PARTIALS;
    Final* fin_dev = reinterpret_cast<Final*>(h->counters + 8);

    if (h->tune.variant == 1 && h->coop_supported)
        {
        int per_sm = 0;
        CAVB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused<FORCE, BUSSI, UNROLL, LB>, threads, 0));
        if (per_sm < 1)
            return (int)cudaErrorLaunchOutOfResources;
        int resident = per_sm * h->num_sms;
        if (max_grid > resident)
            max_grid = resident;
        const int grid = (int)(want < (unsigned long long)max_grid ? want : (unsigned long long)max_grid);
        Partial* recs = h->partials;
        Scalars* sca = h->scalars;
        unsigned long long* arr = h->counters;
        ForceIn ff = f;
        BussiIn bb = b;
        void* args[] = {&ff, &bb, &recs, &sca, &arr};
        CAVB_CHECK(cudaLaunchCooperativeKernel((const void*)k_fused<FORCE, BUSSI, UNROLL, LB>, dim3(grid), dim3(threads),
                                               args, 0, s));
        h->launches += 1;
        return 0;
        }

    const int grid = (int)(want < (unsigned long long)max_grid ? want : (unsigned long long)max_grid);
    k_reduce<FORCE, BUSSI, UNROLL, LB><<<grid, threads, 0, s>>>(f, b, h->partials, h->scalars, fin_dev, h->counters + 4);
    CAVB_CHECK(cudaGetLastError());
    k_apply<FORCE, BUSSI, UNROLL, LB><<<grid, threads, 0, s>>>(f, b, fin_dev);
    CAVB_CHECK(cudaGetLastError());
    h->launches += 2;
    return 0;
    }

template<bool FORCE, bool BUSSI>
static int launch_u(cavb200_handle* h, const ForceIn& f, const BussiIn& b, cudaStream_t s)
    {
    const bool wide = h->tune.threads <= 512 && h->tune.ctas_per_sm == 1;
    switch (h->tune.unroll)
        {
    case 1:
        return wide ? launch_t<FORCE, BUSSI, 1, 512>(h, f, b, s) : launch_t<FORCE, BUSSI, 1, 1024>(h, f, b, s);
    case 4:
        return wide ? launch_t<FORCE, BUSSI, 4, 512>(h, f, b, s) : launch_t<FORCE, BUSSI, 4, 1024>(h, f, b, s);
    default:
        return wide ? launch_t<FORCE, BUSSI, 2, 512>(h, f, b, s) : launch_t<FORCE, BUSSI, 2, 1024>(h, f, b, s);
        }
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
