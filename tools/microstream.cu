// tools/microstream.cu -- which ingredient of the reduce pass costs bandwidth?  In-kernel
// globaltimer span (min start .. max end over CTAs) of a grid-stride read of N particles with
//   A: pos only (LDG.256)            B: pos + vel (2 x LDG.256)
//   C: B + charge (LDG.64)           D: C + image as 3 x LDG.32 (the product's pattern)
//   E: C + image staged per warp as 3 coalesced LDG.32 + shuffles     F: D with all loads L1::no_allocate
//   G: D but software-pipelined (next batch issued before the current one is consumed)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ double4 ldn(const double4* p)
    {
    double4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
    }
__device__ __forceinline__ double4 ldp(const double4* p)
    {
    double4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
    }
__device__ __forceinline__ int ldi_na(const int* p)
    {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
    }
__device__ __forceinline__ void two_sum_acc(double& hi, double& lo, double t)
    {
    const double s = __dadd_rn(hi, t);
    const double bb = __dadd_rn(s, -hi);
    const double e = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(t, -bb));
    hi = s;
    lo = __dadd_rn(lo, e);
    }
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template<int MODE, int U>
__global__ void __launch_bounds__(1024, 1) k(const double4* pos, const double4* vel, const double* q, const int* img, unsigned int N,
                                            double* out, unsigned long long* stamps)
    {
    if (threadIdx.x == 0) stamps[2 * blockIdx.x] = gt();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    const int lane = threadIdx.x & 31;
    if (MODE >= 7)
        {
        // D + real math.  MODE 7: two-sum; 8: plain sums; 9: two-sum, no L-type logic
        double dhi[3] = {0, 0, 0}, dlo[3] = {0, 0, 0}, ke = 0;
        unsigned int cand = 0xffffffffu, nL = 0;
        const double Lx = 566.5, Ly = 566.5, Lz = 566.5;
        for (; i + (U - 1) * stride < N; i += U * stride)
            {
            double4 p[U], v[U]; double c[U]; int ix[U], iy[U], iz[U];
#pragma unroll
            for (int k2 = 0; k2 < U; k2++)
                {
                const unsigned long long j = i + k2 * stride;
                p[k2] = ldn(pos + j); v[k2] = ldp(vel + j); c[k2] = __ldg(q + j);
                ix[k2] = __ldg(img + 3 * j); iy[k2] = __ldg(img + 3 * j + 1); iz[k2] = __ldg(img + 3 * j + 2);
                }
#pragma unroll
            for (int k2 = 0; k2 < U; k2++)
                {
                const double ux = __dadd_rn(p[k2].x, __dmul_rn((double)ix[k2], Lx)), uy = __dadd_rn(p[k2].y, __dmul_rn((double)iy[k2], Ly)),
                             uz = __dadd_rn(p[k2].z, __dmul_rn((double)iz[k2], Lz));
                const double tx = __dmul_rn(c[k2], ux), ty = __dmul_rn(c[k2], uy), tz = __dmul_rn(c[k2], uz);
                bool aside = false;
                if (MODE != 9)
                    {
                    const bool isL = __double2loint(p[k2].w) == 2;
                    aside = isL && cand == 0xffffffffu;
                    if (isL) { nL++; if (aside) cand = (unsigned int)(i + k2 * stride); }
                    }
                if (!aside)
                    {
                    if (MODE == 8) { dhi[0] += tx; dhi[1] += ty; dhi[2] += tz; }
                    else { two_sum_acc(dhi[0], dlo[0], tx); two_sum_acc(dhi[1], dlo[1], ty); two_sum_acc(dhi[2], dlo[2], tz); }
                    }
                ke += v[k2].w * (v[k2].x * v[k2].x + v[k2].y * v[k2].y + v[k2].z * v[k2].z);
                }
            }
        s = dhi[0] + dhi[1] + dhi[2] + dlo[0] + dlo[1] + dlo[2] + ke + (double)cand + (double)nL;
        }
    else if (MODE != 6)
        {
        for (; i + (U - 1) * stride < N; i += U * stride)
            {
            double4 p[U], v[U]; double c[U]; int ix[U], iy[U], iz[U];
#pragma unroll
            for (int k2 = 0; k2 < U; k2++)
                {
                const unsigned long long j = i + k2 * stride;
                p[k2] = ldn(pos + j);
                if (MODE >= 1) v[k2] = (MODE == 5) ? ldn(vel + j) : ldp(vel + j);
                if (MODE >= 2) c[k2] = __ldg(q + j);
                if (MODE == 3) { ix[k2] = __ldg(img + 3 * j); iy[k2] = __ldg(img + 3 * j + 1); iz[k2] = __ldg(img + 3 * j + 2); }
                if (MODE == 5) { ix[k2] = ldi_na(img + 3 * j); iy[k2] = ldi_na(img + 3 * j + 1); iz[k2] = ldi_na(img + 3 * j + 2); }
                if (MODE == 4)
                    {
                    const unsigned long long w0 = 3 * (j - lane); // warp tile: 96 ints, 3 coalesced loads
                    ix[k2] = __ldg(img + w0 + lane); iy[k2] = __ldg(img + w0 + 32 + lane); iz[k2] = __ldg(img + w0 + 64 + lane);
                    }
                }
#pragma unroll
            for (int k2 = 0; k2 < U; k2++)
                {
                s += p[k2].x + p[k2].y + p[k2].z + p[k2].w;
                if (MODE >= 1) s += v[k2].x * v[k2].y + v[k2].z * v[k2].w;
                if (MODE >= 2) s += c[k2];
                if (MODE == 3 || MODE == 5) s += (double)ix[k2] + (double)iy[k2] + (double)iz[k2];
                if (MODE == 4)
                    {
                    // particle `lane` needs ints 3*lane .. 3*lane+2 of the tile
                    int r[3];
#pragma unroll
                    for (int cc = 0; cc < 3; cc++)
                        {
                        const int e = 3 * lane + cc, src = e & 31, reg = e >> 5;
                        const int a0 = __shfl_sync(0xffffffffu, ix[k2], src), a1 = __shfl_sync(0xffffffffu, iy[k2], src),
                                  a2 = __shfl_sync(0xffffffffu, iz[k2], src);
                        r[cc] = reg == 0 ? a0 : (reg == 1 ? a1 : a2);
                        }
                    s += (double)r[0] + (double)r[1] + (double)r[2];
                    }
                }
            }
        }
    else
        {
        // software pipeline, pattern D
        double4 p[U], v[U]; double c[U]; int ix[U], iy[U], iz[U];
        bool have = i + (U - 1) * stride < N;
        if (have)
            {
#pragma unroll
            for (int k2 = 0; k2 < U; k2++)
                {
                const unsigned long long j = i + k2 * stride;
                p[k2] = ldn(pos + j); v[k2] = ldp(vel + j); c[k2] = __ldg(q + j);
                ix[k2] = __ldg(img + 3 * j); iy[k2] = __ldg(img + 3 * j + 1); iz[k2] = __ldg(img + 3 * j + 2);
                }
            }
        while (have)
            {
            const unsigned long long nx = i + U * stride;
            const bool more = nx + (U - 1) * stride < N;
            double4 p2[U], v2[U]; double c2[U]; int jx[U], jy[U], jz[U];
            if (more)
                {
#pragma unroll
                for (int k2 = 0; k2 < U; k2++)
                    {
                    const unsigned long long j = nx + k2 * stride;
                    p2[k2] = ldn(pos + j); v2[k2] = ldp(vel + j); c2[k2] = __ldg(q + j);
                    jx[k2] = __ldg(img + 3 * j); jy[k2] = __ldg(img + 3 * j + 1); jz[k2] = __ldg(img + 3 * j + 2);
                    }
                }
#pragma unroll
            for (int k2 = 0; k2 < U; k2++)
                s += p[k2].x + p[k2].y + p[k2].z + p[k2].w + v[k2].x * v[k2].y + v[k2].z * v[k2].w + c[k2] + (double)ix[k2] + (double)iy[k2] + (double)iz[k2];
            if (more)
                {
#pragma unroll
                for (int k2 = 0; k2 < U; k2++) { p[k2] = p2[k2]; v[k2] = v2[k2]; c[k2] = c2[k2]; ix[k2] = jx[k2]; iy[k2] = jy[k2]; iz[k2] = jz[k2]; }
                }
            i = nx; have = more;
            }
        }
    if (s == 123.4567) out[0] = s;
    __syncthreads();
    if (threadIdx.x == 0) stamps[2 * blockIdx.x + 1] = gt();
    }

int main()
    {
    const unsigned int N = 1000001; const int NB = 8;
    double4 *pos[NB], *vel[NB]; double* q[NB]; int* img[NB];
    for (int b = 0; b < NB; b++)
        {
        CK(cudaMalloc(&pos[b], 32ull * N)); CK(cudaMalloc(&vel[b], 32ull * N)); CK(cudaMalloc(&q[b], 8ull * N)); CK(cudaMalloc(&img[b], 12ull * N + 512));
        const int fill = getenv("FILL") ? atoi(getenv("FILL")) : 0;
        CK(cudaMemset(pos[b], fill, 32ull * N)); CK(cudaMemset(vel[b], fill, 32ull * N)); CK(cudaMemset(q[b], fill, 8ull * N)); CK(cudaMemset(img[b], 0, 12ull * N + 512));
        }
    double* out; CK(cudaMalloc(&out, 64));
    unsigned long long* st; CK(cudaMalloc(&st, 16 * 4096));
    unsigned long long hst[2 * 1024];
    const char* names[] = {"A pos", "B pos+vel", "C +charge", "D +image 3xLDG.32", "E +image coalesced+shfl", "F D all no_allocate", "G D sw-pipelined",
                           "H D + product math", "I D + math, plain sums", "J D + math, no L logic"};
    const double bytes[] = {32, 64, 72, 84, 84, 84, 84, 84, 84, 84};
#define RUN(MODE, U, G, T)                                                                          \
    {                                                                                               \
    double best = 1e9, sum = 0;                                                                     \
    for (int r = 0; r < 12; r++)                                                                    \
        {                                                                                           \
        const int b = r % NB;                                                                       \
        k<MODE, U><<<G, T>>>(pos[b], vel[b], q[b], img[b], N, out, st);                             \
        CK(cudaMemcpy(hst, st, 16 * G, cudaMemcpyDeviceToHost));                                    \
        unsigned long long lo = ~0ull, hi = 0;                                                      \
        for (int c = 0; c < G; c++) { if (hst[2 * c] < lo) lo = hst[2 * c]; if (hst[2 * c + 1] > hi) hi = hst[2 * c + 1]; } \
        const double us = (hi - lo) * 1e-3;                                                         \
        if (r >= 4) { sum += us; if (us < best) best = us; }                                        \
        }                                                                                           \
    printf("%-26s U%d grid %dx%d : mean %6.2f us (best %6.2f)  %7.1f GB/s\n", names[MODE], U, G, T, sum / 8, best, \
           bytes[MODE] * N / (sum / 8 * 1e-6) / 1e9);                                               \
    }
    RUN(3, 4, 148, 512) RUN(7, 4, 148, 512) RUN(8, 4, 148, 512) RUN(9, 4, 148, 512) RUN(7, 2, 296, 512) RUN(8, 2, 296, 512) RUN(9, 2, 296, 512)
    RUN(7, 8, 148, 256) RUN(9, 8, 148, 256) RUN(7, 2, 148, 1024) RUN(9, 2, 148, 1024) RUN(7, 1, 296, 512)
    RUN(0, 4, 148, 512) RUN(1, 4, 148, 512) RUN(2, 4, 148, 512) RUN(3, 4, 148, 512) RUN(4, 4, 148, 512) RUN(5, 4, 148, 512) RUN(6, 2, 148, 512) RUN(6, 4, 148, 512)
    RUN(0, 8, 148, 256) RUN(1, 8, 148, 256) RUN(3, 8, 148, 256) RUN(5, 8, 148, 256)
    RUN(1, 2, 296, 512) RUN(3, 2, 296, 512) RUN(5, 2, 296, 512) RUN(3, 4, 296, 256) RUN(1, 4, 296, 256)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
    }
