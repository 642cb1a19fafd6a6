// tools/microwb.cu -- when does L2 write dirty lines back, and what does that cost the next read pass?
//
// The step is "read 84 B/particle, grid-wide hand-off, write 64 B/particle".  This harness times the
// two halves as separate kernels over 8 rotating systems (steady state, nothing survives in L2 from
// one use of a system to the next) with in-kernel globaltimer spans (first CTA start .. last CTA end):
//   R   grid-stride read of pos, charge, image, vel (84 B), trivial math
//   W   read charge + vel again (40 B, L2 hits when the system was just read), write force + vel (64 B)
// Sequences:  R only | W only | R,W alternating (the product's pattern) -- for every combination of
// load / store cache hints.  If W's dirty lines only leave L2 during the next R, R slows down by the
// write-back traffic and HBM idles during W.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template<int H> __device__ __forceinline__ double4 ld4(const double4* p)
    {
    double4 r;
    if (H == 0)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    else if (H == 1)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    else
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
    }
template<int H> __device__ __forceinline__ void st4(double4* p, const double4& v)
    {
    if (H == 0)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    else if (H == 1)
        asm volatile("st.global.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    else if (H == 2)
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    else if (H == 3)
        asm volatile("st.global.wt.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    else
        asm volatile("st.global.L2::evict_last.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    }

struct Sys
    {
    double4 *pos, *vel, *force;
    double* q;
    int* img;
    };

__device__ __forceinline__ void span(unsigned long long* slot, unsigned long long t0)
    {
    __syncthreads();
    if (threadIdx.x == 0)
        {
        atomicMin(slot, t0);
        atomicMax(slot + 1, gt());
        }
    }

template<int HP, int HV, int U>
__global__ void __launch_bounds__(384, 2) kR(Sys s, unsigned int N, double* out, unsigned long long* slot)
    {
    const unsigned long long t0 = gt();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0;
    for (; i + (U - 1) * stride < N; i += U * stride)
        {
        double4 p[U], v[U];
        double c[U];
        int a[U], b[U], d[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            const unsigned long long j = i + k * stride;
            p[k] = ld4<HP>(s.pos + j);
            v[k] = ld4<HV>(s.vel + j);
            c[k] = __ldg(s.q + j);
            a[k] = __ldg(s.img + 3 * j);
            b[k] = __ldg(s.img + 3 * j + 1);
            d[k] = __ldg(s.img + 3 * j + 2);
            }
#pragma unroll
        for (int k = 0; k < U; k++)
            acc += p[k].x + p[k].y * c[k] + p[k].z + p[k].w + v[k].x * v[k].y + v[k].z + v[k].w + (double)(a[k] + b[k] + d[k]);
        }
    for (; i < N; i += stride)
        acc += ld4<HP>(s.pos + i).x + ld4<HV>(s.vel + i).x;
    if (acc == 123.456)
        out[0] = acc;
    span(slot, t0);
    }

// REREAD: 1 = read charge + vel (what the apply half does), 0 = write constants only
template<int HF, int HW, int U, int REREAD>
__global__ void __launch_bounds__(1024, 1) kW(Sys s, unsigned int N, double alpha, unsigned long long* slot)
    {
    const unsigned long long t0 = gt();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < N; i += U * stride)
        {
        double4 v[U];
        double c[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            const unsigned long long j = i + k * stride;
            if (REREAD)
                {
                v[k] = ld4<2>(s.vel + j);
                c[k] = __ldg(s.q + j);
                }
            else
                {
                v[k] = make_double4(1.0, 2.0, 3.0, 4.0);
                c[k] = 0.5;
                }
            }
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            const unsigned long long j = i + k * stride;
            st4<HF>(s.force + j, make_double4(c[k] * alpha, c[k] * 2.0 * alpha, 0.0, 0.0));
            v[k].x *= alpha;
            v[k].y *= alpha;
            v[k].z *= alpha;
            st4<HW>(s.vel + j, v[k]);
            }
        }
    for (; i < N; i += stride)
        {
        st4<HF>(s.force + i, make_double4(alpha, alpha, 0.0, 0.0));
        double4 v = ld4<2>(s.vel + i);
        v.x *= alpha;
        st4<HW>(s.vel + i, v);
        }
    span(slot, t0);
    }

static unsigned long long* d_slots;
static unsigned long long h_slots[4096];

template<typename F> static void seq(const char* name, int nsteps, F&& launch)
    {
    // slots: 2 per kernel launch (min start, max end); launch(step, slot_base) may use up to 2 kernels
    std::vector<unsigned long long> init(4096);
    for (int k = 0; k < 2048; k++)
        {
        init[2 * k] = ~0ull;
        init[2 * k + 1] = 0;
        }
    CK(cudaMemcpy(d_slots, init.data(), 4096 * 8, cudaMemcpyHostToDevice));
    for (int st = 0; st < nsteps; st++)
        launch(st, d_slots + 4 * st);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_slots, d_slots, 4096 * 8, cudaMemcpyDeviceToHost));
    double a = 0, b = 0, period = 0;
    int n = 0;
    for (int st = nsteps / 2; st < nsteps; st++)
        {
        const unsigned long long* s = h_slots + 4 * st;
        if (s[1])
            a += (s[1] - s[0]) * 1e-3;
        if (s[3])
            b += (s[3] - s[2]) * 1e-3;
        n++;
        }
    const unsigned long long first = h_slots[4 * (nsteps / 2)] != ~0ull ? h_slots[4 * (nsteps / 2)] : h_slots[4 * (nsteps / 2) + 2];
    const unsigned long long last = h_slots[4 * (nsteps - 1)] != ~0ull ? h_slots[4 * (nsteps - 1)] : h_slots[4 * (nsteps - 1) + 2];
    period = (last - first) * 1e-3 / (nsteps - 1 - nsteps / 2);
    printf("%-58s R %6.2f us   W %6.2f us   period %6.2f us\n", name, a / n, b / n, period);
    }

int main()
    {
    const unsigned int N = 1000001;
    const int NB = 8;
    Sys sys[NB];
    std::vector<double> hp(4ull * N), hq(N);
    std::vector<int> hi(3ull * N);
    for (size_t i = 0; i < N; i++)
        {
        for (int c = 0; c < 4; c++)
            hp[4 * i + c] = 1e-3 * ((i * 7 + c * 13) % 1000);
        hq[i] = (i & 1) ? -0.5 : 0.5;
        for (int c = 0; c < 3; c++)
            hi[3 * i + c] = (int)((i + c) % 3) - 1;
        }
    for (int b = 0; b < NB; b++)
        {
        CK(cudaMalloc(&sys[b].pos, 32ull * N));
        CK(cudaMalloc(&sys[b].vel, 32ull * N));
        CK(cudaMalloc(&sys[b].force, 32ull * N));
        CK(cudaMalloc(&sys[b].q, 8ull * N));
        CK(cudaMalloc(&sys[b].img, 12ull * N + 512));
        CK(cudaMemcpy(sys[b].pos, hp.data(), 32ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(sys[b].vel, hp.data(), 32ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(sys[b].q, hq.data(), 8ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(sys[b].img, hi.data(), 12ull * N, cudaMemcpyHostToDevice));
        }
    double* out;
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&d_slots, 4096 * 8));
    const int G = 296, T = 384, S = 200;
    const double al = 1.0000001;

#define R_(HP, HV) kR<HP, HV, 2><<<G, T>>>(sys[st % NB], N, out, sl)
#define W_(HF, HW, RR) kW<HF, HW, 2, RR><<<G, T>>>(sys[st % NB], N, al, sl + 2)
    seq("R only            pos nc, vel nc", S, [&](int st, unsigned long long* sl) { R_(0, 0); });
    seq("R only            pos nc+evict_first, vel plain", S, [&](int st, unsigned long long* sl) { R_(1, 2); });
    seq("W only (no reread) force default, vel default", S, [&](int st, unsigned long long* sl) { W_(0, 0, 0); });
    seq("W only (no reread) force evict_first, vel default", S, [&](int st, unsigned long long* sl) { W_(1, 0, 0); });
    seq("W only (no reread) force .cs, vel .cs", S, [&](int st, unsigned long long* sl) { W_(2, 2, 0); });
    seq("W only (reread)    force evict_first, vel default", S, [&](int st, unsigned long long* sl) { W_(1, 0, 1); });
    seq("R,W  pos nc / force default, vel default", S, [&](int st, unsigned long long* sl) { R_(0, 2); W_(0, 0, 1); });
    seq("R,W  pos evict_first / force evict_first, vel default [product]", S, [&](int st, unsigned long long* sl) { R_(1, 2); W_(1, 0, 1); });
    seq("R,W  pos evict_first / force evict_first, vel evict_first", S, [&](int st, unsigned long long* sl) { R_(1, 2); W_(1, 1, 1); });
    seq("R,W  pos evict_first / force .cs, vel .cs", S, [&](int st, unsigned long long* sl) { R_(1, 2); W_(2, 2, 1); });
    seq("R,W  pos evict_first / force .wt, vel .wt", S, [&](int st, unsigned long long* sl) { R_(1, 2); W_(3, 3, 1); });
    seq("R,W  pos nc / force .cs, vel .cs", S, [&](int st, unsigned long long* sl) { R_(0, 2); W_(2, 2, 1); });
    seq("R,W  pos evict_first / force evict_last, vel evict_last", S, [&](int st, unsigned long long* sl) { R_(1, 2); W_(4, 4, 1); });
    seq("R,W  pos nc, vel nc(no reuse) / force default, vel default", S, [&](int st, unsigned long long* sl) { R_(0, 0); W_(0, 0, 1); });
    seq("W same system (L2 resident, no reread)", S, [&](int st, unsigned long long* sl) { kW<0, 0, 2, 0><<<G, T>>>(sys[0], N, al, sl + 2); });
    seq("W same system (L2 resident, reread)", S, [&](int st, unsigned long long* sl) { kW<0, 0, 2, 1><<<G, T>>>(sys[0], N, al, sl + 2); });
    seq("R same system (L2 resident)", S, [&](int st, unsigned long long* sl) { kR<0, 0, 2><<<G, T>>>(sys[0], N, out, sl); });
    seq("R,W  product hints, W without reread", S, [&](int st, unsigned long long* sl) { R_(1, 2); W_(1, 0, 0); });
    seq("R,W  product hints, W 148x1024 threads", S, [&](int st, unsigned long long* sl) { R_(1, 2); kW<1, 0, 2, 1><<<148, 1024>>>(sys[st % NB], N, al, sl + 2); });
    seq("R,W  product hints, W U4", S, [&](int st, unsigned long long* sl) { R_(1, 2); kW<1, 0, 4, 1><<<G, T>>>(sys[st % NB], N, al, sl + 2); });
    seq("R,W  product hints, W U1 592x384?", S, [&](int st, unsigned long long* sl) { R_(1, 2); kW<1, 0, 1, 1><<<G, T>>>(sys[st % NB], N, al, sl + 2); });
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
    }
