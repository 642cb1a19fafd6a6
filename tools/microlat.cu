// tools/microlat.cu -- instruction latencies that bound the combine step (B200):
// dependent DADD / DFMA chains, 64-bit shuffle chains, __syncthreads with 16 / 32 warps,
// L2 round trip (ld.cg), relaxed polling round trip, __threadfence.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_lat(double* out, long long* cyc, const double* src, unsigned long long* flag)
    {
    __shared__ double sm[64];
    double x = src[threadIdx.x & 31], y = src[1];
    long long t0, t1;
    // DADD chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        {
        x = __dadd_rn(x, y); x = __dadd_rn(x, y); x = __dadd_rn(x, y); x = __dadd_rn(x, y);
        x = __dadd_rn(x, y); x = __dadd_rn(x, y); x = __dadd_rn(x, y); x = __dadd_rn(x, y);
        }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 512;
    // shuffle chain (64-bit = 2 x SHFL)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        {
        x = __shfl_xor_sync(0xffffffffu, x, 1); x = __shfl_xor_sync(0xffffffffu, x, 2);
        x = __shfl_xor_sync(0xffffffffu, x, 4); x = __shfl_xor_sync(0xffffffffu, x, 8);
        }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 256;
    // __syncthreads
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
    // L2 round trip: dependent ld.cg chain on a pointer-chase of zeros
    const double* p = src;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        {
        double v;
        asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
        p = src + (long long)v; // v == 0
        }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 64;
    // relaxed flag load round trip
    unsigned long long acc = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        {
        unsigned long long v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag + (acc & 1)) : "memory");
        acc += v;
        }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 64;
    // __threadfence after a store
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; i++)
        {
        if (threadIdx.x == 0) flag[8] = i;
        __threadfence();
        }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0) / 16;
    // __syncthreads_count + or
    int c = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        c += __syncthreads_count(x > 0.5 + i);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) / 64;
    // DDIV + DSQRT
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        y = __dsqrt_rn(__ddiv_rn(y + 3.0, 1.5));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0) / 64;
    sm[threadIdx.x & 63] = x + y + (double)acc + c;
    out[threadIdx.x] = sm[(threadIdx.x + 1) & 63] + (double)(long long)p;
    }
int main()
    {
    double *out, *src; long long* cyc; unsigned long long* flag;
    cudaMalloc(&out, 8192 * 8); cudaMalloc(&src, 8192); cudaMalloc(&cyc, 64); cudaMalloc(&flag, 256);
    cudaMemset(src, 0, 8192); cudaMemset(flag, 0, 256);
    for (int threads : {32, 512, 1024})
        {
        k_lat<<<148, threads>>>(out, cyc, src, flag);
        k_lat<<<148, threads>>>(out, cyc, src, flag);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d: DADD %lld cyc | SHFL64 %lld | syncthreads %lld | ld.cg L2 %lld | ld.relaxed %lld | st+threadfence %lld | syncthreads_count %lld | ddiv+dsqrt %lld\n",
               threads, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
    }
