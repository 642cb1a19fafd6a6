// debug.cu -- measurement helpers of bench.py / tools (not on the product path):
//   cavb200_debug_delay          a one-thread kernel that holds the stream for a given time, so that a timed region can
//                                be enqueued completely before its first kernel starts (launch-latency-free timing)
//   cavb200_debug_fp64_peak      DFMA microbenchmark: the FP64 pipe's instruction rate, the roofline denominator of the
//                                F(k,t) kernel (FP64-ALU bound)
//   cavb200_debug_launch_ring    per-launch {first CTA start, last CTA end} device timestamps (tuning stamps = 2)
#include "cavb200_internal.cuh"

using namespace cavb;

namespace
    {
__global__ void k_delay(unsigned long long ns)
    {
    const unsigned long long t0 = globaltimer_ns();
    while (globaltimer_ns() - t0 < ns)
        __nanosleep(200);
    }

// 8 independent DFMA chains per thread, CHAIN instructions each per loop trip: the FP64 pipe is the only busy unit
template<int CHAINS> __global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int trips)
    {
    double x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++)
        x[k] = (double)(threadIdx.x + k);
#pragma unroll 1
    for (int t = 0; t < trips; t++)
        {
#pragma unroll
        for (int r = 0; r < 8; r++)
            {
#pragma unroll
            for (int k = 0; k < CHAINS; k++)
                x[k] = fma(x[k], a, b);
            }
        }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++)
        s += x[k];
    if (s == 123.456) // never true: keeps the chains alive
        out[0] = s;
    }
    } // namespace

extern "C" int cavb200_debug_delay(cavb200_handle* h, uint64_t ns, void* stream)
    {
    if (!h || ns > 100000000ull) // at most 100 ms
        return (int)cudaErrorInvalidValue;
    k_delay<<<1, 1, 0, (cudaStream_t)stream>>>(ns);
    CAVB_CHECK(cudaGetLastError());
    return 0;
    }

extern "C" int cavb200_debug_fp64_peak(cavb200_handle* h, double* dfma_per_s)
    {
    if (!h || !dfma_per_s)
        return (int)cudaErrorInvalidValue;
    CAVB_CHECK(cudaSetDevice(h->device));
    double* out = nullptr;
    CAVB_CHECK(cudaMalloc((void**)&out, 8));
    cudaEvent_t e0, e1;
    CAVB_CHECK(cudaEventCreate(&e0));
    CAVB_CHECK(cudaEventCreate(&e1));
    constexpr int CHAINS = 8;
    const int trips = 4096, grid = h->num_sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++)
        {
        CAVB_CHECK(cudaEventRecord(e0, 0));
        k_dfma<CHAINS><<<grid, threads>>>(out, 1.0000001, 1e-9, trips);
        CAVB_CHECK(cudaEventRecord(e1, 0));
        CAVB_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CAVB_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)grid * threads * CHAINS * 8.0 * trips / (ms * 1e-3);
        if (rep > 0 && rate > best) // the first repetition warms up
            best = rate;
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *dfma_per_s = best;
    return 0;
    }

extern "C" int cavb200_debug_launch_ring(cavb200_handle* h, int reset, uint64_t* out, uint32_t n_pairs, uint64_t* epoch)
    {
    if (!h || n_pairs > LAUNCH_RING)
        return (int)cudaErrorInvalidValue;
    CAVB_CHECK(cudaDeviceSynchronize());
    unsigned long long* ring = h->stamps + LAUNCH_RING_OFFSET;
    if (out && n_pairs)
        CAVB_CHECK(cudaMemcpy(out, ring, 16ull * n_pairs, cudaMemcpyDeviceToHost));
    if (epoch)
        CAVB_CHECK(cudaMemcpy(epoch, h->counters + 2, 8, cudaMemcpyDeviceToHost));
    if (reset)
        {
        // {start = ~0 (atomicMin target), end = 0 (atomicMax target)}
        static unsigned long long init[2 * LAUNCH_RING];
        for (unsigned long long i = 0; i < LAUNCH_RING; i++)
            {
            init[2 * i] = ~0ull;
            init[2 * i + 1] = 0ull;
            }
        CAVB_CHECK(cudaMemcpy(ring, init, sizeof(init), cudaMemcpyHostToDevice));
        }
    return 0;
    }
