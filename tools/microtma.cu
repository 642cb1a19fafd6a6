// tools/microtma.cu -- prototype of a TMA-fed reduce pass and a bulk-store write pass, timed like
// tools/microwb.cu (in-kernel globaltimer span, 8 rotating systems, alone and alternating R,W).
//   kR_tma   one producer warp issues cp.async.bulk (1-D bulk copies, mbarrier complete_tx) of
//            384-particle chunks of pos / vel / charge / image into a shared-memory ring; twelve
//            consumer warps take one particle per thread out of the ring and run the PRODUCT's math
//            (take_particle + KE from hotpath.cuh)
//   kR_ldg   the product's register-staged loop (reduce_stream) for comparison
//   kW_stg   STG.256 stores (force + vel), kW_bulk  the same bytes as cp.async.bulk smem->global
#include "../cav_hoomd_b200/csrc/hotpath.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
using namespace cavb;

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count)
    {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes)
    {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
    {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity)
    {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
    }
__device__ __forceinline__ unsigned long long policy_evict_first()
    {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
    }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar)
    {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
    }
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned int bytes, unsigned long long* bar,
                                              unsigned long long pol)
    {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
    }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned int bytes)
    {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template<int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int CH = 384; // particles per chunk = consumer threads

struct __align__(128) Stage
    {
    double4 pos[CH];
    double4 vel[CH];
    double charge[CH];
    int image[3 * CH];
    };

__device__ __forceinline__ void span(unsigned long long* slot, unsigned long long t0)
    {
    __syncthreads();
    if (threadIdx.x == 0)
        {
        atomicMin(slot, t0);
        atomicMax(slot + 1, globaltimer_ns());
        }
    }

template<int S, int HINT>
__global__ void __launch_bounds__(CH + 32, 2) kR_tma(ForceIn f, BussiIn b, double* out, unsigned long long* slot)
    {
    extern __shared__ __align__(128) unsigned char dyn[];
    Stage* ring = reinterpret_cast<Stage*>(dyn);
    __shared__ unsigned long long full[S], empty[S];
    __shared__ double red[16];
    const unsigned long long t0 = globaltimer_ns();
    const unsigned int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0)
        {
        for (int s = 0; s < S; s++)
            {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CH / 32);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    __syncthreads();
    const unsigned int N = f.N, nk_all = b.n;
    const unsigned int nchunks = (N + CH - 1) / CH;
    Acc a;
    acc_zero(a);
    double ke = 0.0;
    if (warp == CH / 32)
        {
        if (lane == 0)
            {
            const unsigned long long pol = policy_evict_first();
            unsigned int j = 0;
            for (unsigned int c = blockIdx.x; c < nchunks; c += gridDim.x, j++)
                {
                const unsigned int s = j % S;
                if (j >= S)
                    mbar_wait(&empty[s], ((j / S) - 1) & 1);
                const unsigned int base = c * CH;
                const unsigned int nf = min((unsigned int)CH, N - base);
                const unsigned int nb = nf & ~3u;
                const unsigned int nk = base < nk_all ? min((unsigned int)CH, nk_all - base) : 0u;
                mbar_expect_tx(&full[s], nb * 52u + nk * 32u);
                Stage& st = ring[s];
                if (nb)
                    {
                    if (HINT)
                        {
                        bulk_g2s_hint(st.pos, f.pos + base, nb * 32u, &full[s], pol);
                        bulk_g2s_hint(st.image, f.image + 3ull * base, nb * 12u, &full[s], pol);
                        }
                    else
                        {
                        bulk_g2s(st.pos, f.pos + base, nb * 32u, &full[s]);
                        bulk_g2s(st.image, f.image + 3ull * base, nb * 12u, &full[s]);
                        }
                    bulk_g2s(st.charge, f.charge + base, nb * 8u, &full[s]);
                    }
                if (nk)
                    bulk_g2s(st.vel, b.vel + base, nk * 32u, &full[s]);
                }
            }
        }
    else
        {
        unsigned int j = 0;
        for (unsigned int c = blockIdx.x; c < nchunks; c += gridDim.x, j++)
            {
            const unsigned int s = j % S;
            const unsigned int base = c * CH;
            const unsigned int nf = min((unsigned int)CH, N - base);
            const unsigned int nb = nf & ~3u;
            const unsigned int nk = base < nk_all ? min((unsigned int)CH, nk_all - base) : 0u;
            mbar_wait(&full[s], (j / S) & 1);
            const Stage& st = ring[s];
            double4 p = make_double4(0, 0, 0, 0), v = make_double4(0, 0, 0, 0);
            double q = 0;
            int ix = 0, iy = 0, iz = 0;
            if (tid < nb)
                {
                const double2 p0 = reinterpret_cast<const double2*>(st.pos)[2 * tid];
                const double2 p1 = reinterpret_cast<const double2*>(st.pos)[2 * tid + 1];
                p = make_double4(p0.x, p0.y, p1.x, p1.y);
                q = st.charge[tid];
                ix = st.image[3 * tid];
                iy = st.image[3 * tid + 1];
                iz = st.image[3 * tid + 2];
                }
            else if (tid < nf)
                {
                const unsigned long long i = (unsigned long long)base + tid;
                p = ld256_stream(f.pos + i);
                q = __ldg(f.charge + i);
                ix = __ldg(f.image + 3 * i);
                iy = __ldg(f.image + 3 * i + 1);
                iz = __ldg(f.image + 3 * i + 2);
                }
            if (tid < nk)
                {
                const double2 v0 = reinterpret_cast<const double2*>(st.vel)[2 * tid];
                const double2 v1 = reinterpret_cast<const double2*>(st.vel)[2 * tid + 1];
                v = make_double4(v0.x, v0.y, v1.x, v1.y);
                }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[s]);
            if (tid < nf)
                take_particle(a, base + tid, p, q, ix, iy, iz, f);
            ke += v.w * (v.x * v.x + v.y * v.y + v.z * v.z);
            }
        }
    // keep the results alive (a stand-in for block_merge)
    double t = a.dhi[0] + a.dhi[1] + a.dhi[2] + a.dlo[0] + a.dlo[1] + a.dlo[2] + ke + (double)a.cand + (double)a.n_L;
    for (int m = 16; m >= 1; m >>= 1)
        t += __shfl_xor_sync(0xffffffffu, t, m);
    if (lane == 0 && warp < 16)
        red[warp] = t;
    __syncthreads();
    if (tid == 0)
        {
        double r = 0;
        for (int w = 0; w < (CH + 32) / 32; w++)
            r += red[w];
        out[blockIdx.x] = r;
        }
    span(slot, t0);
    }

template<int U>
__global__ void __launch_bounds__(384, 2) kR_ldg(ForceIn f, BussiIn b, double* out, unsigned long long* slot)
    {
    __shared__ double red[16];
    const unsigned long long t0 = globaltimer_ns();
    Acc a;
    acc_zero(a);
    reduce_stream<true, true, U>(a, f, b);
    double t = a.dhi[0] + a.dhi[1] + a.dhi[2] + a.dlo[0] + a.dlo[1] + a.dlo[2] + a.ke + (double)a.cand + (double)a.n_L;
    for (int m = 16; m >= 1; m >>= 1)
        t += __shfl_xor_sync(0xffffffffu, t, m);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0)
        {
        double r = 0;
        for (int w = 0; w < 12; w++)
            r += red[w];
        out[blockIdx.x] = r;
        }
    span(slot, t0);
    }

template<int U>
__global__ void __launch_bounds__(384, 2) kW_stg(ForceIn f, BussiIn b, double alpha, unsigned long long* slot)
    {
    const unsigned long long t0 = globaltimer_ns();
    Final fin = {};
    fin.has_photon = 1;
    fin.photon_local = f.N - 1;
    fin.Dq[0] = 0.3;
    fin.Dq[1] = -0.2;
    fin.alpha = alpha;
    fin.bussi_ok = 1;
    apply_stream<true, true, U>(fin, f, b);
    span(slot, t0);
    }

// same bytes as kW_stg, written with bulk smem->global copies of constant chunks (no reads)
template<int DEPTH>
__global__ void __launch_bounds__(384, 2) kW_bulk(ForceIn f, BussiIn b, unsigned long long* slot)
    {
    extern __shared__ __align__(128) unsigned char dyn[];
    double4* buf = reinterpret_cast<double4*>(dyn); // 2 x CH double4
    const unsigned long long t0 = globaltimer_ns();
    for (int k = threadIdx.x; k < 2 * CH; k += blockDim.x)
        buf[k] = make_double4(1.0, 2.0, 3.0, 4.0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const unsigned int N = f.N, nchunks = (N + CH - 1) / CH;
    if (threadIdx.x == 0)
        {
        for (unsigned int c = blockIdx.x; c < nchunks; c += gridDim.x)
            {
            const unsigned int base = c * CH;
            const unsigned int n = min((unsigned int)CH, N - base);
            bulk_s2g(f.force + base, buf, n * 32u);
            bulk_s2g(b.vel + base, buf + CH, n * 32u);
            bulk_commit();
            bulk_wait_read<DEPTH>();
            }
        bulk_wait<0>();
        }
    span(slot, t0);
    }

static unsigned long long* d_slots;
static unsigned long long h_slots[4096];
template<typename F> static void seq(const char* name, int nsteps, F&& launch)
    {
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
    double a = 0, b = 0;
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
    printf("%-64s R %6.2f us   W %6.2f us\n", name, a / n, b / n);
    }

int main()
    {
    const unsigned int N = 1000001;
    const int NB = 8;
    double4 *pos[NB], *vel[NB], *force[NB];
    double* q[NB];
    int* img[NB];
    std::vector<double> hp(4ull * N), hv(4ull * N), hq(N);
    std::vector<int> hi(3ull * N);
    for (size_t i = 0; i < N; i++)
        {
        for (int c = 0; c < 3; c++)
            {
            hp[4 * i + c] = (double)((i * 7 + c * 13) % 1000) - 500.0;
            hv[4 * i + c] = 1e-3 * ((i + c) % 17);
            hi[3 * i + c] = (int)((i + c) % 3) - 1;
            }
        hp[4 * i + 3] = 0.0;
        hv[4 * i + 3] = 29166.0;
        hq[i] = (i & 1) ? -0.5 : 0.5;
        }
    long long two = 2;
    memcpy(&hp[4ull * (N - 1) + 3], &two, 8);
    hq[N - 1] = 0;
    for (int b = 0; b < NB; b++)
        {
        CK(cudaMalloc(&pos[b], 32ull * N));
        CK(cudaMalloc(&vel[b], 32ull * N));
        CK(cudaMalloc(&force[b], 32ull * N));
        CK(cudaMalloc(&q[b], 8ull * N));
        CK(cudaMalloc(&img[b], 12ull * N + 512));
        CK(cudaMemcpy(pos[b], hp.data(), 32ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(vel[b], hv.data(), 32ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(q[b], hq.data(), 8ull * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(img[b], hi.data(), 12ull * N, cudaMemcpyHostToDevice));
        }
    double* out;
    CK(cudaMalloc(&out, 8 * 4096));
    CK(cudaMalloc(&d_slots, 4096 * 8));
    const int G = 296, S = 200;
    auto mkf = [&](int bb) {
        ForceIn f = {};
        f.pos = pos[bb];
        f.charge = q[bb];
        f.image = img[bb];
        f.force = force[bb];
        f.N = N;
        f.Lx = f.Ly = f.Lz = 566.5;
        f.L_typeid = 2;
        f.g = 1e-3;
        f.K = 1e-4;
        fill_force_constants(f);
        return f;
    };
    auto mkb = [&](int bb) {
        BussiIn b = {};
        b.vel = vel[bb];
        b.n = N - 1;
        b.first = 0;
        b.rescale = 1;
        return b;
    };
    CK(cudaFuncSetAttribute(kR_tma<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (int)sizeof(Stage)));
    CK(cudaFuncSetAttribute(kR_tma<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (int)sizeof(Stage)));
    CK(cudaFuncSetAttribute(kR_tma<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(Stage)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kR_tma<3, 1>, CH + 32, 3 * sizeof(Stage)));
    printf("kR_tma<3> occupancy %d CTAs/SM, stage %zu B\n", occ, sizeof(Stage));
    const double al = 1.0000001;
#define RT(SS, H) kR_tma<SS, H><<<G, CH + 32, SS * sizeof(Stage)>>>(mkf(st % NB), mkb(st % NB), out, sl)
#define RL(U) kR_ldg<U><<<G, 384>>>(mkf(st % NB), mkb(st % NB), out, sl)
#define WS(U) kW_stg<U><<<G, 384>>>(mkf(st % NB), mkb(st % NB), al, sl + 2)
#define WB(D) kW_bulk<D><<<G, 384, 2 * CH * 32>>>(mkf(st % NB), mkb(st % NB), sl + 2)
    seq("R only  ldg U2 (product loop)", S, [&](int st, unsigned long long* sl) { RL(2); });
    seq("R only  tma 3 stages, evict_first pos/image", S, [&](int st, unsigned long long* sl) { RT(3, 1); });
    seq("R only  tma 3 stages, no hints", S, [&](int st, unsigned long long* sl) { RT(3, 0); });
    seq("R only  tma 2 stages, evict_first pos/image", S, [&](int st, unsigned long long* sl) { RT(2, 1); });
    seq("W only  stg U2 (product apply loop, rereads vel+charge cold)", S, [&](int st, unsigned long long* sl) { WS(2); });
    seq("W only  bulk stores depth 2", S, [&](int st, unsigned long long* sl) { WB(2); });
    seq("W only  bulk stores depth 6", S, [&](int st, unsigned long long* sl) { WB(6); });
    seq("R,W  ldg U2 / stg U2   [product]", S, [&](int st, unsigned long long* sl) { RL(2); WS(2); });
    seq("R,W  tma 3 hints / stg U2", S, [&](int st, unsigned long long* sl) { RT(3, 1); WS(2); });
    seq("R,W  tma 3 no hints / stg U2", S, [&](int st, unsigned long long* sl) { RT(3, 0); WS(2); });
    seq("R,W  tma 3 hints / bulk depth 6 (no reread)", S, [&](int st, unsigned long long* sl) { RT(3, 1); WB(6); });
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
    }
