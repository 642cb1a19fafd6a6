// tools/microprod.cu -- the PRODUCT's reduce pass (hotpath.cuh) timed piece by piece with in-kernel
// globaltimer spans, same harness as microstream.cu:  0 = reduce_stream only, 1 = + block_merge,
// 2 = + publish_record.   Real synthetic-like data (non-zero), 8 rotating systems.
#include "../cav_hoomd_b200/csrc/hotpath.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
using namespace cavb;
template<int MODE, bool FORCE, bool KE, int U, int LB, int MINB>
__global__ void __launch_bounds__(LB, MINB) kp(ForceIn f, BussiIn b, Partial* recs, unsigned long long* stamps, double* out)
    {
    __shared__ BlockScratch sc;
    if (threadIdx.x == 0) stamps[2 * blockIdx.x] = globaltimer_ns();
    Acc a;
    acc_zero(a);
    reduce_stream<FORCE, KE, U>(a, f, b);
    if (MODE >= 1)
        {
        block_merge<FORCE, KE>(a, f, sc);
        if (MODE >= 2 && threadIdx.x == 0)
            publish_record(recs + blockIdx.x, sc.rec, 7ull);
        }
    else if (a.dhi[0] + a.dhi[1] + a.dhi[2] + a.ke + a.dlo[0] == 123.456)
        out[0] = a.dlo[1] + a.dlo[2] + a.n_L + a.cand;
    __syncthreads();
    if (threadIdx.x == 0) stamps[2 * blockIdx.x + 1] = globaltimer_ns();
    }
int main()
    {
    const unsigned int N = 1000001; const int NB = 8;
    double4 *pos[NB], *vel[NB]; double* q[NB]; int* img[NB];
    std::vector<double> hp(4ull * N), hv(4ull * N), hq(N); std::vector<int> hi(3ull * N);
    for (size_t i = 0; i < N; i++)
        {
        for (int c = 0; c < 3; c++) { hp[4 * i + c] = (double)((i * 7 + c * 13) % 1000) - 500.0; hv[4 * i + c] = 1e-3 * ((i + c) % 17); hi[3 * i + c] = (int)((i + c) % 3) - 1; }
        hp[4 * i + 3] = 0.0; hv[4 * i + 3] = 29166.0; hq[i] = (i & 1) ? -0.5 : 0.5;
        }
    long long two = 2; memcpy(&hp[4ull * (N - 1) + 3], &two, 8); hq[N - 1] = 0;
    for (int b = 0; b < NB; b++)
        {
        CK(cudaMalloc(&pos[b], 32ull * N)); CK(cudaMalloc(&vel[b], 32ull * N)); CK(cudaMalloc(&q[b], 8ull * N)); CK(cudaMalloc(&img[b], 12ull * N + 512));
        CK(cudaMemcpy(pos[b], hp.data(), 32ull * N, cudaMemcpyHostToDevice)); CK(cudaMemcpy(vel[b], hv.data(), 32ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(q[b], hq.data(), 8ull * N, cudaMemcpyHostToDevice)); CK(cudaMemcpy(img[b], hi.data(), 12ull * N, cudaMemcpyHostToDevice));
        }
    double* out; CK(cudaMalloc(&out, 64));
    Partial* recs; CK(cudaMalloc(&recs, sizeof(Partial) * 2048));
    unsigned long long* st; CK(cudaMalloc(&st, 16 * 4096));
    unsigned long long hst[2 * 2048];
#define RUN2(MODE, FORCE, KE, U, G, T, LB, MINB)                                                                 \
    {                                                                                               \
    double best = 1e9, sum = 0;                                                                     \
    for (int r = 0; r < 12; r++)                                                                    \
        {                                                                                           \
        const int bb = r % NB;                                                                      \
        ForceIn f = {}; f.pos = pos[bb]; f.charge = q[bb]; f.image = img[bb]; f.N = N; f.Lx = f.Ly = f.Lz = 566.5; f.L_typeid = 2; f.g = 1e-3; f.K = 1e-4; \
        BussiIn b = {}; b.vel = vel[bb]; b.n = N - 1; b.first = 0;                                    \
        kp<MODE, FORCE, KE, U, LB, MINB><<<G, T>>>(f, b, recs, st, out);                                       \
        CK(cudaMemcpy(hst, st, 16 * G, cudaMemcpyDeviceToHost));                                    \
        unsigned long long lo = ~0ull, hi2 = 0;                                                     \
        for (int c = 0; c < G; c++) { if (hst[2 * c] < lo) lo = hst[2 * c]; if (hst[2 * c + 1] > hi2) hi2 = hst[2 * c + 1]; } \
        const double us = (hi2 - lo) * 1e-3;                                                        \
        if (r >= 4) { sum += us; if (us < best) best = us; }                                        \
        }                                                                                           \
    printf("mode %d force %d ke %d U%d grid %dx%d bounds(%d,%d): mean %6.2f us (best %6.2f)\n", MODE, FORCE, KE, U, G, T, LB, MINB, sum / 8, best); \
    }
    RUN2(1, true, true, 2, 296, 512, 1024, 1) RUN2(1, true, true, 2, 296, 384, 384, 2) RUN2(1, true, true, 2, 444, 256, 256, 3)
    RUN2(1, true, true, 2, 296, 448, 448, 2) RUN2(1, true, true, 4, 296, 384, 384, 2) RUN2(1, true, true, 1, 296, 512, 512, 2)
    RUN2(1, true, true, 1, 592, 256, 256, 4) RUN2(1, true, true, 2, 592, 192, 192, 4) RUN2(1, true, true, 2, 444, 288, 288, 3)
    RUN2(1, true, false, 2, 296, 384, 384, 2) RUN2(1, false, true, 2, 296, 384, 384, 2) RUN2(1, true, false, 2, 296, 512, 512, 2) RUN2(1, false, true, 4, 296, 512, 512, 2)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
    }
