// tools/microtree.cu -- cycles of the pieces of the combine step (one warp active, B200)
#include "../cav_hoomd_b200/csrc/hotpath.cuh"
#include <stdio.h>
using namespace cavb;
__global__ void k(double* out, long long* cyc, const double* src)
    {
    __shared__ BlockScratch sc;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Acc a; acc_zero(a);
    for (int k2 = 0; k2 < 3; k2++) { a.dhi[k2] = src[lane + k2]; a.dlo[k2] = src[lane + 3 + k2] * 1e-17; }
    a.ke = src[lane + 7];
    long long t0 = clock64();
    warp_tree<true, true>(a);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    cross_warp_put<true, true>(a, sc, lane, warp);
    __syncthreads();
    t0 = clock64();
    Acc b;
    if (warp == 0) cross_warp_get<true, true>(b, sc, lane, blockDim.x / 32);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
    unsigned long long key = warp_min_u64((unsigned long long)lane * 77 + 5);
    unsigned int c = __reduce_add_sync(0xffffffffu, lane & 1);
    unsigned int bal = __ballot_sync(0xffffffffu, key == lane);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // finalize-like chain
    ForceIn f = {}; f.g = 1e-3; f.K = 1e-4; f.gk = 10; f.half_K = 5e-5; f.half_g2K = 5e-3; f.N = 100; 
    BussiIn bb = {}; bb.rescale = 1; bb.dof = 3e6; bb.half_kT = 1.5e-4; bb.omc = 2e-4; bb.gR2 = 3e6; bb.two_R = 0.2; bb.c = 0.9998; bb.cdof = 3e6; bb.den = 0.1; bb.r_normal = 0.1;
    if (threadIdx.x == 0)
        {
        for (int k2 = 0; k2 < 3; k2++) { sc.rec.dhi[k2] = a.dhi[k2]; sc.rec.dlo[k2] = a.dlo[k2]; sc.rec.q[k2] = 1.5 + k2; }
        sc.rec.ke = 900.0 + a.ke; sc.rec.first_L = 50; sc.flags = 0;
        t0 = clock64();
        finalize<true, true>(f, bb, sc, (Scalars*)nullptr, false);
        t1 = clock64();
        cyc[3] = t1 - t0;
        }
    __syncthreads();
    out[threadIdx.x] = b.dhi[0] + b.ke + (double)key + c + bal + sc.fin.alpha + sc.fin.Dq[0];
    }
int main()
    {
    double *out, *src; long long* cyc;
    cudaMalloc(&out, 8192 * 8); cudaMalloc(&src, 8192); cudaMalloc(&cyc, 64);
    double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1.0 + i * 0.37;
    cudaMemcpy(src, h, 8192, cudaMemcpyHostToDevice);
    for (int threads : {32, 384, 1024})
        {
        k<<<148, threads>>>(out, cyc, src); k<<<148, threads>>>(out, cyc, src);
        cudaDeviceSynchronize();
        long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d: warp_tree<force,ke> %lld cyc | cross_warp_get %lld | vote(min64+redux+ballot) %lld | finalize %lld\n", threads, c[0], c[1], c[2], c[3]);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    }
