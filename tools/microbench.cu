// tools/microbench.cu -- ground truth for the design of the streaming kernels on B200:
//   (1) how fast can a SHORT kernel read 32..84 MB (LDG.256 with different launch shapes, TMA bulk
//       copies into a shared-memory ring), rotating over buffers larger than L2;
//   (2) what a grid barrier costs (acquire polling vs relaxed polling) and what an empty
//       cooperative / ordinary launch costs;
//   (3) short write streams.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do                                                                                          \
        {                                                                                       \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess)                                                                   \
            {                                                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
            exit(1);                                                                            \
            }                                                                                   \
        } while (0)

__device__ __forceinline__ double4 ld256s(const double4* p)
    {
    double4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
    }
__device__ __forceinline__ void st256(double4* p, double4 v)
    {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    }

// ---- (1a) LDG.256 grid-stride read, UNROLL loads in flight per thread -------------------------
template<int U> __global__ void __launch_bounds__(1024, 1) k_read(const double4* __restrict__ a, size_t n, double* out)
    {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    for (; i + (U - 1) * stride < n; i += U * stride)
        {
        double4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            v[k] = ld256s(a + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; k++)
            s += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    for (; i < n; i += stride)
        {
        double4 v = ld256s(a + i);
        s += v.x + v.y + v.z + v.w;
        }
    if (s == 123.456)
        out[0] = s;
    }

// contiguous chunk per CTA instead of grid stride
template<int U> __global__ void __launch_bounds__(1024, 1) k_read_chunk(const double4* __restrict__ a, size_t n, double* out)
    {
    const size_t per = (n + gridDim.x - 1) / gridDim.x;
    const size_t lo = per * blockIdx.x, hi = min(n, lo + per);
    size_t i = lo + threadIdx.x;
    const size_t stride = blockDim.x;
    double s = 0;
    for (; i + (U - 1) * stride < hi; i += U * stride)
        {
        double4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            v[k] = ld256s(a + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; k++)
            s += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    for (; i < hi; i += stride)
        {
        double4 v = ld256s(a + i);
        s += v.x + v.y + v.z + v.w;
        }
    if (s == 123.456)
        out[0] = s;
    }

// ---- (1b) TMA bulk copies (cp.async.bulk) into a shared-memory ring, consumed by LDS -----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
    {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
    {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
    {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
    }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
    {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
    }

// TILE double4 elements per stage, STAGES stages; one elected thread issues, all threads consume
template<int TILE, int STAGES> __global__ void __launch_bounds__(1024, 1)
    k_read_tma(const double4* __restrict__ a, size_t n, double* out)
    {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double4* ring = reinterpret_cast<double4*>(smem_raw);
    __shared__ uint64_t full[STAGES];
    const size_t ntiles = n / TILE; // benchmark: n is a multiple of TILE
    if (threadIdx.x == 0)
        {
        for (int s = 0; s < STAGES; s++)
            mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    __syncthreads();
    // tiles of this CTA: blockIdx.x, blockIdx.x + grid, ...
    size_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && (size_t)s < my_tiles; s++)
            {
            mbar_expect_tx(&full[s], TILE * 32);
            bulk_g2s(ring + (size_t)s * TILE, a + (blockIdx.x + (size_t)s * gridDim.x) * TILE, TILE * 32, &full[s]);
            }
    double sum = 0;
    for (size_t t = 0; t < my_tiles; t++)
        {
        const int s = (int)(t % STAGES);
        const uint32_t parity = (uint32_t)((t / STAGES) & 1);
        mbar_wait(&full[s], parity);
        for (int j = threadIdx.x; j < TILE; j += blockDim.x)
            {
            const double4 v = ring[(size_t)s * TILE + j];
            sum += v.x + v.y + v.z + v.w;
            }
        __syncthreads(); // everyone done with stage s
        if (threadIdx.x == 0 && t + STAGES < my_tiles)
            {
            mbar_expect_tx(&full[s], TILE * 32);
            bulk_g2s(ring + (size_t)s * TILE, a + (blockIdx.x + (t + STAGES) * gridDim.x) * TILE, TILE * 32, &full[s]);
            }
        }
    if (sum == 123.456)
        out[0] = sum;
    }

// ---- (3) write stream ---------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) k_write(double4* a, size_t n, double v)
    {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        st256(a + i, make_double4(v, v, 0, 0));
    }

// read + write (rescale-like), L2-hot second pass can be modelled by calling it twice on the same buffer
__global__ void __launch_bounds__(1024, 1) k_scale(double4* a, size_t n, double alpha)
    {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        {
        double4 v;
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(a + i) : "memory");
        v.x *= alpha;
        v.y *= alpha;
        v.z *= alpha;
        st256(a + i, v);
        }
    }

// ---- (2) barriers ----------------------------------------------------------------------------
__global__ void k_empty(int* p)
    {
    if (p && threadIdx.x == 9999)
        *p = 1;
    }

template<int MODE> __global__ void __launch_bounds__(1024, 1) k_barrier(unsigned long long* ctr, int rounds, double* out)
    {
    // MODE 0: ld.acquire.gpu polling; MODE 1: ld.relaxed (volatile) polling + one fence; self-resetting like the product
    for (int r = 0; r < rounds; r++)
        {
        __syncthreads();
        if (threadIdx.x == 0)
            {
            __threadfence();
            unsigned long long old;
            asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(ctr + 2 * r), "l"(1ull) : "memory");
            unsigned long long v;
            do
                {
                if (MODE == 0)
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr + 2 * r) : "memory");
                else
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr + 2 * r) : "memory");
                } while (v < gridDim.x);
            if (MODE == 1)
                __threadfence();
            asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(ctr + 2 * r + 1), "l"(1ull) : "memory");
            if (old == gridDim.x - 1)
                {
                ctr[2 * r] = 0;
                ctr[2 * r + 1] = 0;
                }
            }
        __syncthreads();
        }
    if (out && threadIdx.x == 9999)
        out[0] = 1;
    }

template<typename F> static float time_us(F f, int reps)
    {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; i++)
        f(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; i++)
        f(i);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms * 1e3f / reps;
    }

int main()
    {
    const int NBUF = 12;
    const size_t MB = 1 << 20;
    const size_t bytes = 96 * MB; // each buffer; NBUF * 96 MB = 1.1 GB >> L2
    double4* buf[NBUF];
    for (int i = 0; i < NBUF; i++)
        {
        CK(cudaMalloc(&buf[i], bytes));
        CK(cudaMemset(buf[i], 0, bytes));
        }
    double* out;
    CK(cudaMalloc(&out, 64));
    unsigned long long* ctr;
    CK(cudaMalloc(&ctr, 4096));
    CK(cudaMemset(ctr, 0, 4096));
    const int reps = 120;

    printf("== empty launches ==\n");
    printf("ordinary 296x512: %.2f us\n", time_us([&](int) { k_empty<<<296, 512>>>(nullptr); }, 400));
        {
        int* np = nullptr;
        void* args[] = {&np};
        printf("cooperative 296x512: %.2f us\n",
               time_us([&](int) { cudaLaunchCooperativeKernel((const void*)k_empty, dim3(296), dim3(512), args, 0, 0); }, 400));
        printf("cooperative 148x1024: %.2f us\n",
               time_us([&](int) { cudaLaunchCooperativeKernel((const void*)k_empty, dim3(148), dim3(1024), args, 0, 0); }, 400));
        }

    printf("== grid barrier (kernel = N barrier rounds; per-round cost = (t(9) - t(1)) / 8) ==\n");
    for (int mode = 0; mode < 2; mode++)
        for (int shape = 0; shape < 2; shape++)
            {
            const int g = shape ? 148 : 296, t = shape ? 1024 : 512;
            float tt[2];
            for (int k = 0; k < 2; k++)
                {
                int rounds = k ? 9 : 1;
                double* o = nullptr;
                void* args[] = {&ctr, &rounds, &o};
                const void* fn = mode ? (const void*)k_barrier<1> : (const void*)k_barrier<0>;
                tt[k] = time_us([&](int) { cudaLaunchCooperativeKernel(fn, dim3(g), dim3(t), args, 0, 0); }, 200);
                }
            printf("mode %d (%s) grid %dx%d: 1 round %.2f us, 9 rounds %.2f us -> %.2f us per barrier\n", mode,
                   mode ? "relaxed poll" : "acquire poll", g, t, tt[0], tt[1], (tt[1] - tt[0]) / 8);
            }

    printf("== short read streams (GB/s; rotating %d buffers) ==\n", NBUF);
    for (size_t mb : {32, 52, 84})
        {
        const size_t n = mb * MB / 32;
        printf("-- %zu MB --\n", mb);
#define RUN_READ(NAME, KERNEL, G, T)                                                                   \
    {                                                                                                  \
    float us = time_us([&](int i) { KERNEL<<<G, T>>>(buf[i % NBUF], n, out); }, reps);                   \
    printf("%-34s grid %4dx%-4d : %7.2f us  %7.1f GB/s\n", NAME, G, T, us, mb * MB / (us * 1e-6) / 1e9); \
    }
        RUN_READ("ldg256 stride U1", k_read<1>, 296, 512);
        RUN_READ("ldg256 stride U2", k_read<2>, 296, 512);
        RUN_READ("ldg256 stride U4", k_read<4>, 296, 512);
        RUN_READ("ldg256 stride U8", k_read<8>, 296, 512);
        RUN_READ("ldg256 stride U4", k_read<4>, 148, 1024);
        RUN_READ("ldg256 stride U8", k_read<8>, 148, 1024);
        RUN_READ("ldg256 stride U4", k_read<4>, 592, 256);
        RUN_READ("ldg256 stride U4 (oversubscribed)", k_read<4>, 1184, 512);
        RUN_READ("ldg256 stride U2 (oversubscribed)", k_read<2>, 2368, 256);
        RUN_READ("ldg256 chunk U4", k_read_chunk<4>, 296, 512);
        RUN_READ("ldg256 chunk U8", k_read_chunk<8>, 148, 1024);
#define RUN_TMA(TILE, STAGES, G, T)                                                                              \
    {                                                                                                            \
    size_t sm = (size_t)TILE * 32 * STAGES;                                                                      \
    CK(cudaFuncSetAttribute(k_read_tma<TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
    float us = time_us([&](int i) { k_read_tma<TILE, STAGES><<<G, T, sm>>>(buf[i % NBUF], n, out); }, reps);        \
    printf("tma bulk tile %4d x %2d stages (%3zu KB) grid %4dx%-4d : %7.2f us  %7.1f GB/s\n", TILE, STAGES, sm >> 10, G, T, us, \
           mb * MB / (us * 1e-6) / 1e9);                                                                         \
    }
        RUN_TMA(256, 8, 148, 512);
        RUN_TMA(256, 16, 148, 512);
        RUN_TMA(256, 24, 148, 512);
        RUN_TMA(512, 8, 148, 512);
        RUN_TMA(512, 12, 148, 1024);
        RUN_TMA(1024, 6, 148, 1024);
        RUN_TMA(256, 8, 296, 256);
        RUN_TMA(256, 12, 296, 512);
        }

    printf("== short write / scale streams ==\n");
    for (size_t mb : {32, 64})
        {
        const size_t n = mb * MB / 32;
        float us = time_us([&](int i) { k_write<<<296, 512>>>(buf[i % NBUF], n, 1.0); }, reps);
        printf("write %zu MB grid 296x512: %.2f us  %.1f GB/s\n", mb, us, mb * MB / (us * 1e-6) / 1e9);
        us = time_us([&](int i) { k_write<<<1184, 512>>>(buf[i % NBUF], n, 1.0); }, reps);
        printf("write %zu MB grid 1184x512: %.2f us  %.1f GB/s\n", mb, us, mb * MB / (us * 1e-6) / 1e9);
        us = time_us([&](int i) { k_scale<<<296, 512>>>(buf[i % NBUF], n, 1.0000001); }, reps);
        printf("scale (cold) %zu MB grid 296x512: %.2f us  %.1f GB/s (r+w)\n", mb, us, 2 * mb * MB / (us * 1e-6) / 1e9);
        us = time_us(
            [&](int i)
            {
                k_read<4><<<296, 512>>>(buf[i % NBUF], n, out);
                k_scale<<<296, 512>>>(buf[i % NBUF], n, 1.0000001);
            },
            reps);
        printf("read then scale (L2-hot) %zu MB: %.2f us per pair\n", mb, us);
        }
    return 0;
    }
