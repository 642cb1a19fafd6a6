// rhok.cu -- F(k,t): the density field rho[t][k] = sum_j exp(i k.r_j(t)) for a batch of frames,
// and the origin/lag correlation F = mean_k Re(rho_0 conj(rho_t)).
//
// Batched form of compute_density_field / compute_field_autocorr
// (reference src/cavitymd/analysis.py:34-47, :359-364), which loops over k in Python and makes K
// NumPy passes over N every step.  Here one launch covers T frames x K wavevectors:
//   grid (P, T): CTA (p, t) owns a contiguous slice of frame t's particles; a tile of positions is
//   staged in shared memory (coalesced, stride 3 or 4 -> xyz); thread (group, k) keeps rho_k's
//   (re, im) in registers and walks the tile's particles group, group+G, ... (broadcast LDS);
//   groups are folded in shared memory; slices are folded by a second tiny kernel in slice order.
// The sum order is fixed, so results are bitwise reproducible.  The work is FP64 sincos bound
// (about 50 fp64 instructions per (particle, k) pair against 24/K bytes), not HBM bound, and the
// contraction is only k x 3 deep: no tensor-core shape here (BASELINE.json north_star (3)).
#include "cavb200_internal.cuh"

namespace cavb
    {
#ifndef RHOK_MIN_CTAS
#define RHOK_MIN_CTAS 2
#endif

#ifndef RHOK_KB
#define RHOK_KB 4
#endif
constexpr int RHOK_KB_HOST = RHOK_KB;
constexpr int RHOK_TILE = 1024; // particles per shared-memory tile (24 KB)
constexpr uint32_t RHOK_F32 = 13;  // internal `stride` code: float32 xyz positions (frame_stride then counts floats)

// sin and cos of one argument, |x| < 2^20, both to < 1 ulp-ish (abs. error ~2e-16).
// The library sincos() costs ~70 issue slots per call here, two thirds of them integer/selection/slow-path
// plumbing, and `ncu` showed the first kernel issue-bound (88 % issue slots busy, FP64 pipe at 50 %,
// profiles/ncu_rhok_r1a_summary.txt).  This one is 21 FP64 instructions and ~8 integer ones:
//   n = rint(x * 2/pi) by the 1.5*2^52 trick;  r = x - n*pio2_hi - n*pio2_lo with two FMAs (n*pio2_hi is
//   exact inside the FMA for |n| < 2^20);  fdlibm's __kernel_sin/__kernel_cos minimax polynomials on
//   [-pi/4, pi/4] (public-domain constants, < 1 ulp);  quadrant fix-up by swapping and flipping sign bits.
// (the constants sit in constant memory so that every DFMA takes its coefficient as a c[bank][offset]
// operand; as literals the compiler rebuilt each one with two UMOVs inside the loop -- 25 extra issue
// slots per pair)
__constant__ double SC[18] = {
    6755399441055744.0,          // 0  1.5 * 2^52
    6.36619772367581382433e-01,  // 1  2/pi
    1.57079632679489655800e+00,  // 2  pi/2 hi
    6.12323399573676603587e-17,  // 3  pi/2 lo
    1.58969099521155010221e-10,  // 4  S6
    -2.50507602534068634195e-08, // 5  S5
    2.75573137070700676789e-06,  // 6  S4
    -1.98412698298579493134e-04, // 7  S3
    8.33333333332248946124e-03,  // 8  S2
    -1.66666666666666324348e-01, // 9  S1
    -1.13596475577881948265e-11, // 10 C6
    2.08757232129817482790e-09,  // 11 C5
    -2.75573143513906633035e-07, // 12 C4
    2.48015872894767294178e-05,  // 13 C3
    -1.38888888888741095749e-03, // 14 C2
    4.16666666666666019037e-02,  // 15 C1
    -0.5, 1.0};
// sin and cos of the REDUCED argument r = x - q pi/2 and the quadrant count q (two's complement, only q & 3 matters)
__device__ __forceinline__ void sincos_reduced(double x, double& sr, double& cr, int& q)
    {
    const double t = __fma_rn(x, SC[1], SC[0]);
    q = __double2loint(t);
    const double n = __dadd_rn(t, -SC[0]);
    double r = __fma_rn(-n, SC[2], x);
    r = __fma_rn(-n, SC[3], r);
    const double z = __dmul_rn(r, r);
    double ps = __fma_rn(z, SC[4], SC[5]);
    ps = __fma_rn(z, ps, SC[6]);
    ps = __fma_rn(z, ps, SC[7]);
    ps = __fma_rn(z, ps, SC[8]);
    ps = __fma_rn(z, ps, SC[9]);
    sr = __fma_rn(__dmul_rn(r, z), ps, r);
    double pc = __fma_rn(z, SC[10], SC[11]);
    pc = __fma_rn(z, pc, SC[12]);
    pc = __fma_rn(z, pc, SC[13]);
    pc = __fma_rn(z, pc, SC[14]);
    pc = __fma_rn(z, pc, SC[15]);
    cr = __fma_rn(__dmul_rn(z, z), pc, __fma_rn(z, SC[16], SC[17]));
    }
__device__ __forceinline__ void sincos_lean(double x, double& s, double& c)
    {
    double sr, cr;
    int q;
    sincos_reduced(x, sr, cr, q);
    // quadrant q & 3:  0: (s, c) = (sr, cr)   1: (cr, -sr)   2: (-sr, -cr)   3: (-cr, sr)
    const bool odd = q & 1;
    const double ss = odd ? cr : sr;
    const double cc = odd ? sr : cr;
    const int fs = (q & 2) << 30;       // sign flip of sin in quadrants 2, 3
    const int fc = ((q + 1) & 2) << 30; // sign flip of cos in quadrants 1, 2
    s = __hiloint2double(__double2hiint(ss) ^ fs, __double2loint(ss));
    c = __hiloint2double(__double2hiint(cc) ^ fc, __double2loint(cc));
    }

__device__ __forceinline__ unsigned int abs_hi(double v) { return (unsigned int)__double2hiint(v) & 0x7fffffffu; }

// One thread's walk over a staged tile: particles pg, pg + groups, ... against its KB wave vectors.
// CHECK = false: the caller has bounded every |k.r| of the tile below 2^20, so the inner loop carries no range test
// (2-3 integer instructions per (particle, k) pair that share the issue port with the half-rate FP64 pipe).
// The quadrant fix-up of a pair costs ten integer / select instructions (swap sin and cos: 4 FSEL; two sign flips; the
// parity predicate), and every one of them takes an issue slot the half-rate FP64 pipe could have used.  Instead the pair
// is ROTATED by the quadrant with the FP64 pipe itself: (cq, sq) = (cos, sin)(q pi/2) in {0, +-1} comes out of a four-entry
// shared-memory table, and  re += cr cq - sr sq,  im += sr cq + cr sq  as four FMAs in place of the two additions.  One
// product of each line is an exact zero, so the sums are bit for bit what the swap-and-flip gave.
template<int KB, bool CHECK>
__device__ __forceinline__ void rhok_walk(const double* sx, const double* sy, const double* sz, const double2* rot, uint32_t pg,
                                          uint32_t n, uint32_t groups, const double (&kx)[KB], const double (&ky)[KB],
                                          const double (&kz)[KB], double (&re)[KB], double (&im)[KB])
    {
    for (uint32_t j = pg; j < n; j += groups)
        {
        const double x = sx[j], y = sy[j], z = sz[j];
        double kr[KB];
        bool huge = false;
#pragma unroll
        for (int m = 0; m < KB; m++)
            {
            // analysis.py:42 np.dot(positions, k_vec) = x kx + y ky + z kz (NumPy hands this to a BLAS
            // gemv, which contracts to FMAs on any AVX2 host; so does this)
            kr[m] = __fma_rn(z, kz[m], __fma_rn(y, ky[m], __dmul_rn(x, kx[m])));
            // |kr| >= 2^20, inf or nan: exponent test on the integer pipe (DSETP would sit on the FP64 pipe)
            if (CHECK)
                huge = huge || (__double2hiint(kr[m]) & 0x7fffffff) >= 0x41300000;
            }
        if (!CHECK || !huge)
            {
            // one basic block for the KB independent chains: coefficients are fetched once
            // (written stage by stage over the KB arguments instead -- every dependent pair 2 KB instructions apart in
            // program order -- ptxas interleaves three or four chains instead of two or three, needs 40 B of spills
            // and the frame takes 0.1277 ms instead of 0.1256: not what limits the FP64 pipe at 70 %)
#pragma unroll
            for (int m = 0; m < KB; m++)
                {
                double sr, cr;
                int q;
                sincos_reduced(kr[m], sr, cr, q);
                const double2 w = rot[q & 3];
                re[m] = __fma_rn(-sr, w.y, __fma_rn(cr, w.x, re[m]));
                im[m] = __fma_rn(cr, w.y, __fma_rn(sr, w.x, im[m]));
                }
            }
        else
            {
            for (int m = 0; m < KB; m++)
                {
                double sn, cs;
                sincos(kr[m], &sn, &cs); // huge arguments (and inf/nan): the library's Payne-Hanek path
                re[m] += cs;
                im[m] += sn;
                }
            }
        }
    }

// grid (P, T); thread (pg, kq): particle group pg walks the tile with stride `groups`, kq owns KB
// consecutive wave vectors whose components and (re, im) accumulators live in registers, so the three
// shared-memory loads and the loop bookkeeping of a particle are paid once per KB (particle, k) pairs.
template<int KB>
__global__ void __launch_bounds__(256, RHOK_MIN_CTAS)
    k_rhok(const double* __restrict__ pos, uint32_t stride, unsigned long long frame_stride, uint32_t N,
           const double* __restrict__ kvec, uint32_t k0, uint32_t Ks, uint32_t K, double* __restrict__ out, uint32_t P,
           int direct)
    {
    __shared__ double sx[RHOK_TILE], sy[RHOK_TILE], sz[RHOK_TILE];
    __shared__ unsigned int tile_maxhi[2]; // largest hi word of |coordinate| in the tile being staged (by tile parity)
    __shared__ double2 rot[4];             // (cos, sin)(q pi/2), q = 0..3
    extern __shared__ double sred[]; // [groups][KQ * KB][2]

    const uint32_t t = blockIdx.y, p = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    const uint32_t KQ = (Ks + KB - 1) / KB;      // k-quads of this slab (host guarantees KQ <= blockDim.x)
    const uint32_t groups = blockDim.x / KQ;
    const uint32_t kq = tid % KQ, pg = tid / KQ;
    const bool active = pg < groups;

    double kx[KB], ky[KB], kz[KB], re[KB], im[KB];
#pragma unroll
    for (int m = 0; m < KB; m++)
        {
        const uint32_t k = kq * KB + m;
        const bool in = active && k < Ks;
        kx[m] = in ? __ldg(kvec + 3 * (k0 + k) + 0) : 0.0;
        ky[m] = in ? __ldg(kvec + 3 * (k0 + k) + 1) : 0.0;
        kz[m] = in ? __ldg(kvec + 3 * (k0 + k) + 2) : 0.0;
        re[m] = im[m] = 0.0;
        }
    double kbound = 0.0; // max over this thread's wave vectors of |kx| + |ky| + |kz| >= |k.r| / max|coordinate|
#pragma unroll
    for (int m = 0; m < KB; m++)
        kbound = fmax(kbound, fabs(kx[m]) + fabs(ky[m]) + fabs(kz[m]));
    if (tid < 2)
        tile_maxhi[tid] = 0u;
    if (tid < 4)
        rot[tid] = make_double2(tid == 0 ? 1.0 : (tid == 2 ? -1.0 : 0.0), tid == 1 ? 1.0 : (tid == 3 ? -1.0 : 0.0));

    // contiguous slice of this frame, in whole tiles
    const uint32_t tiles = (N + RHOK_TILE - 1) / RHOK_TILE;
    const uint32_t tiles_per = (tiles + P - 1) / P;
    const uint32_t tile_lo = p * tiles_per;
    const uint32_t tile_hi = min(tiles, tile_lo + tiles_per);
    const double* frame = pos + (unsigned long long)t * frame_stride;

    for (uint32_t tile = tile_lo; tile < tile_hi; tile++)
        {
        const uint32_t base = tile * RHOK_TILE;
        const uint32_t n = min((uint32_t)RHOK_TILE, N - base);
        __syncthreads();
        unsigned int mx = 0u; // largest hi word of |coordinate| this thread stages
        if (stride == 4)
            {
            const double4* src = reinterpret_cast<const double4*>(frame) + base;
            for (uint32_t j = tid; j < n; j += blockDim.x)
                {
                const double4 r = ld256_stream(src + j);
                sx[j] = r.x;
                sy[j] = r.y;
                sz[j] = r.z;
                mx = max(mx, max(max(abs_hi(r.x), abs_hi(r.y)), abs_hi(r.z)));
                }
            }
        else if (stride == RHOK_F32)
            {
            // float32 xyz, the way GSD stores positions: widened exactly, as NumPy does in np.dot(f32, f64)
            const float* src = reinterpret_cast<const float*>(pos) + (unsigned long long)t * frame_stride + 3ull * base;
            for (uint32_t e = tid; e < 3 * n; e += blockDim.x)
                {
                const double v = (double)__ldg(src + e);
                const uint32_t j = e / 3, c = e - 3 * j;
                (c == 0 ? sx : (c == 1 ? sy : sz))[j] = v;
                mx = max(mx, abs_hi(v));
                }
            }
        else
            {
            const double* src = frame + 3ull * base;
            for (uint32_t e = tid; e < 3 * n; e += blockDim.x)
                {
                const double v = __ldg(src + e);
                const uint32_t j = e / 3, c = e - 3 * j;
                (c == 0 ? sx : (c == 1 ? sy : sz))[j] = v;
                mx = max(mx, abs_hi(v));
                }
            }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((tid & 31u) == 0)
            atomicMax(&tile_maxhi[tile & 1u], mx);
        __syncthreads();
        // upper bound of |coordinate| over the tile (hi word + 1, lo word 0 bounds every double with that hi word); with
        // the thread's own bound on |kx| + |ky| + |kz| it decides once per tile whether any argument can reach 2^20
        const unsigned int mh = tile_maxhi[tile & 1u];
        if (tid == 0)
            tile_maxhi[(tile + 1u) & 1u] = 0u;
        const bool fast = mh < 0x7ff00000u && __dmul_rn(kbound, __hiloint2double((int)(mh + 1u), 0)) < 1048575.0;
        if (active)
            {
            if (fast)
                rhok_walk<KB, false>(sx, sy, sz, rot, pg, n, groups, kx, ky, kz, re, im);
            else
                rhok_walk<KB, true>(sx, sy, sz, rot, pg, n, groups, kx, ky, kz, re, im);
            }
        }

    // fold the particle groups (fixed order)
    const uint32_t KP = KQ * KB;
    if (active)
        {
#pragma unroll
        for (int m = 0; m < KB; m++)
            {
            sred[(pg * KP + kq * KB + m) * 2 + 0] = re[m];
            sred[(pg * KP + kq * KB + m) * 2 + 1] = im[m];
            }
        }
    __syncthreads();
    for (uint32_t k = tid; k < Ks; k += blockDim.x)
        {
        double r = 0.0, i = 0.0;
        for (uint32_t g = 0; g < groups; g++)
            {
            r += sred[(g * KP + k) * 2 + 0];
            i += sred[(g * KP + k) * 2 + 1];
            }
        double* dst = direct ? out + ((unsigned long long)t * K + k0 + k) * 2
                             : out + (((unsigned long long)t * P + p) * K + k0 + k) * 2;
        dst[0] = r;
        dst[1] = i;
        }
    }

// one warp per (t, k): lane L adds slices L, L + 32, ... (independent loads), then a fixed shuffle tree
__global__ void k_rhok_fold(const double* __restrict__ part, uint32_t P, uint32_t K, uint32_t T, double* __restrict__ rho)
    {
    const unsigned long long e = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // (t, k) pair
    const uint32_t lane = threadIdx.x & 31;
    if (e >= (unsigned long long)T * K)
        return;
    const unsigned long long t = e / K, k = e % K;
    double r = 0.0, i = 0.0;
    for (uint32_t p = lane; p < P; p += 32)
        {
        const double2 v = *reinterpret_cast<const double2*>(part + ((t * P + p) * K + k) * 2);
        r += v.x;
        i += v.y;
        }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        {
        r += __shfl_xor_sync(0xffffffffu, r, m);
        i += __shfl_xor_sync(0xffffffffu, i, m);
        }
    if (lane == 0)
        {
        rho[e * 2 + 0] = r;
        rho[e * 2 + 1] = i;
        }
    }

// F[o][l] = mean_k Re(rho[o][k] conj(rho[o+l][k]))  (analysis.py:361: np.mean(np.real(f0 * np.conj(ft))))
__global__ void k_fkt(const double* __restrict__ rho, uint32_t T, uint32_t K, uint32_t n_origins, uint32_t n_lags,
                      double* __restrict__ out)
    {
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (unsigned long long)n_origins * n_lags)
        return;
    const unsigned long long o = e / n_lags, l = e % n_lags;
    if (o + l >= T)
        {
        out[e] = __longlong_as_double(0x7ff8000000000000ll);
        return;
        }
    const double* a = rho + o * K * 2;
    const double* b = rho + (o + l) * K * 2;
    double acc = 0.0;
    for (uint32_t k = 0; k < K; k++)
        {
        // Re((ar + i ai)(br - i bi)) = ar br + ai bi
        acc += a[2 * k] * b[2 * k] + a[2 * k + 1] * b[2 * k + 1];
        }
    out[e] = acc / (double)K;
    }
    } // namespace cavb

using namespace cavb;

static int rhok_launch(cavb200_handle* h, const double* pos, uint32_t stride, uint64_t frame_stride, uint32_t N,
                       uint32_t T, const double* kvec, uint32_t K, double* rho, void* stream)
    {
    if (!h)
        return (int)cudaErrorInvalidValue;
    if (T == 0 || K == 0)
        return 0;
    if (!kvec || !rho || (N > 0 && !pos) || (stride != 3 && stride != 4 && stride != RHOK_F32))
        return (int)cudaErrorInvalidValue;
    if (stride == 4 && ((reinterpret_cast<uintptr_t>(pos) & 31) || (frame_stride & 3)))
        return (int)cudaErrorMisalignedAddress;
    if (stride == RHOK_F32 && (reinterpret_cast<uintptr_t>(pos) & 3))
        return (int)cudaErrorMisalignedAddress;
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0)
        {
        CAVB_CHECK(cudaMemsetAsync(rho, 0, sizeof(double) * 2ull * K * T, s));
        return 0;
        }
    // 128-thread CTAs (four per SM) when every k-quad still gets at least four particle groups: a CTA that is staging
    // its next tile then idles a quarter of the SM's warps instead of half (8 frames per launch: 0.1239 -> 0.1213 ms per
    // frame; 64 frames: equal).  192 threads: worse (0.143).
    const int threads = (h->tune.rhok_threads == 128 || h->tune.rhok_threads == 256)
                            ? h->tune.rhok_threads
                            : ((K + RHOK_KB_HOST - 1) / RHOK_KB_HOST <= 32 ? 128 : 256);
    const uint32_t tiles = (N + RHOK_TILE - 1) / RHOK_TILE;
    uint32_t P = (24u * (uint32_t)h->num_sms + T - 1) / T;
    if (P > tiles)
        P = tiles;
    if (P < 1)
        P = 1;
    const int direct = (P == 1);
    double* target = rho;
    if (!direct)
        {
        const uint64_t need = sizeof(double) * 2ull * K * T * P;
        if (h->rhok_partials_bytes < need)
            {
            // growing the workspace is the one place a call may synchronise (first call / larger batch)
            CAVB_CHECK(cudaStreamSynchronize(s));
            cudaFree(h->rhok_partials);
            h->rhok_partials = nullptr;
            h->rhok_partials_bytes = 0;
            CAVB_CHECK(cudaMalloc((void**)&h->rhok_partials, need));
            h->rhok_partials_bytes = need;
            }
        target = h->rhok_partials;
        }
    // wave vectors in slabs of at most threads*KB = 1024 (one launch for the usual K); KB = 4 k per thread
#ifndef RHOK_KB
#define RHOK_KB 4
#endif
    constexpr int KB = RHOK_KB;
    for (uint32_t k0 = 0; k0 < K; k0 += (uint32_t)threads * KB)
        {
        const uint32_t Ks = (K - k0) < (uint32_t)threads * KB ? (K - k0) : (uint32_t)threads * KB;
        const uint32_t KQ = (Ks + KB - 1) / KB;
        const uint32_t groups = (uint32_t)threads / KQ;
        const size_t smem = sizeof(double) * 2 * groups * KQ * KB;
        k_rhok<KB><<<dim3(P, T), threads, smem, s>>>(pos, stride, frame_stride, N, kvec, k0, Ks, K, target, P, direct);
        CAVB_CHECK(cudaGetLastError());
        h->launches += 1;
        }
    if (!direct)
        {
        const unsigned long long pairs = (unsigned long long)T * K;
        k_rhok_fold<<<(unsigned int)((pairs * 32 + 255) / 256), 256, 0, s>>>(target, P, K, T, rho);
        CAVB_CHECK(cudaGetLastError());
        h->launches += 1;
        }
    return 0;
    }

extern "C" int cavb200_rhok(cavb200_handle* h, const double* pos, uint32_t stride, uint64_t frame_stride, uint32_t N,
                            uint32_t T, const double* kvec, uint32_t K, double* rho, void* stream)
    {
    if (stride != 3 && stride != 4)
        return (int)cudaErrorInvalidValue;
    return rhok_launch(h, pos, stride, frame_stride, N, T, kvec, K, rho, stream);
    }

extern "C" int cavb200_rhok_f32(cavb200_handle* h, const float* pos_xyz, uint64_t frame_stride, uint32_t N, uint32_t T,
                                const double* kvec, uint32_t K, double* rho, void* stream)
    {
    return rhok_launch(h, reinterpret_cast<const double*>(pos_xyz), RHOK_F32, frame_stride, N, T, kvec, K, rho, stream);
    }

extern "C" int cavb200_fkt(cavb200_handle* h, const double* rho, uint32_t T, uint32_t K, uint32_t n_origins,
                           uint32_t n_lags, double* out, void* stream)
    {
    if (!h)
        return (int)cudaErrorInvalidValue;
    if (n_origins == 0 || n_lags == 0)
        return 0;
    if (!rho || !out || K == 0)
        return (int)cudaErrorInvalidValue;
    const unsigned long long n = (unsigned long long)n_origins * n_lags;
    k_fkt<<<(unsigned int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rho, T, K, n_origins, n_lags, out);
    CAVB_CHECK(cudaGetLastError());
    h->launches += 1;
    return 0;
    }
