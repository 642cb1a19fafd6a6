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
constexpr int RHOK_TILE = 1024; // particles per shared-memory tile (24 KB)

__global__ void __launch_bounds__(1024)
    k_rhok(const double* __restrict__ pos, uint32_t stride, unsigned long long frame_stride, uint32_t N,
           const double* __restrict__ kvec, uint32_t k0, uint32_t KS, uint32_t K, double* __restrict__ out,
           uint32_t P, int direct)
    {
    __shared__ double sx[RHOK_TILE], sy[RHOK_TILE], sz[RHOK_TILE];
    extern __shared__ double sred[]; // [groups][KS][2]

    const uint32_t t = blockIdx.y, p = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    const uint32_t groups = blockDim.x / KS;
    const uint32_t kslot = tid % KS, group = tid / KS;
    const bool active = group < groups;

    double kx = 0, ky = 0, kz = 0;
    if (active)
        {
        kx = __ldg(kvec + 3 * (k0 + kslot) + 0);
        ky = __ldg(kvec + 3 * (k0 + kslot) + 1);
        kz = __ldg(kvec + 3 * (k0 + kslot) + 2);
        }

    // contiguous slice of this frame, in whole tiles
    const uint32_t tiles = (N + RHOK_TILE - 1) / RHOK_TILE;
    const uint32_t tiles_per = (tiles + P - 1) / P;
    const uint32_t tile_lo = p * tiles_per;
    const uint32_t tile_hi = min(tiles, tile_lo + tiles_per);
    const double* frame = pos + (unsigned long long)t * frame_stride;

    double re = 0.0, im = 0.0;
    for (uint32_t tile = tile_lo; tile < tile_hi; tile++)
        {
        const uint32_t base = tile * RHOK_TILE;
        const uint32_t n = min((uint32_t)RHOK_TILE, N - base);
        __syncthreads();
        if (stride == 4)
            {
            const double4* src = reinterpret_cast<const double4*>(frame) + base;
            for (uint32_t j = tid; j < n; j += blockDim.x)
                {
                const double4 r = ld256_stream(src + j);
                sx[j] = r.x;
                sy[j] = r.y;
                sz[j] = r.z;
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
                }
            }
        __syncthreads();
        if (active)
            {
            for (uint32_t j = group; j < n; j += groups)
                {
                // analysis.py:42 np.dot(positions, k_vec): x kx + y ky + z kz
                const double kr = __dadd_rn(__dadd_rn(__dmul_rn(sx[j], kx), __dmul_rn(sy[j], ky)), __dmul_rn(sz[j], kz));
                double s, c;
                sincos(kr, &s, &c);
                re += c;
                im += s;
                }
            }
        }

    // fold the groups (fixed order)
    if (active)
        {
        sred[(group * KS + kslot) * 2 + 0] = re;
        sred[(group * KS + kslot) * 2 + 1] = im;
        }
    __syncthreads();
    if (tid < KS)
        {
        double r = 0.0, i = 0.0;
        for (uint32_t g = 0; g < groups; g++)
            {
            r += sred[(g * KS + tid) * 2 + 0];
            i += sred[(g * KS + tid) * 2 + 1];
            }
        double* dst = direct ? out + ((unsigned long long)t * K + k0 + tid) * 2
                             : out + (((unsigned long long)t * P + p) * K + k0 + tid) * 2;
        dst[0] = r;
        dst[1] = i;
        }
    }

__global__ void k_rhok_fold(const double* __restrict__ part, uint32_t P, uint32_t K, uint32_t T, double* __restrict__ rho)
    {
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; // (t, k) pair
    if (e >= (unsigned long long)T * K)
        return;
    const unsigned long long t = e / K, k = e % K;
    double r = 0.0, i = 0.0;
    for (uint32_t p = 0; p < P; p++)
        {
        const double* src = part + ((t * P + p) * K + k) * 2;
        r += src[0];
        i += src[1];
        }
    rho[e * 2 + 0] = r;
    rho[e * 2 + 1] = i;
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

extern "C" int cavb200_rhok(cavb200_handle* h, const double* pos, uint32_t stride, uint64_t frame_stride, uint32_t N,
                            uint32_t T, const double* kvec, uint32_t K, double* rho, void* stream)
    {
    if (!h)
        return (int)cudaErrorInvalidValue;
    if (T == 0 || K == 0)
        return 0;
    if (!kvec || !rho || (N > 0 && !pos) || (stride != 3 && stride != 4))
        return (int)cudaErrorInvalidValue;
    if (stride == 4 && ((reinterpret_cast<uintptr_t>(pos) & 31) || (frame_stride & 3)))
        return (int)cudaErrorMisalignedAddress;
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0)
        {
        CAVB_CHECK(cudaMemsetAsync(rho, 0, sizeof(double) * 2ull * K * T, s));
        return 0;
        }
    const int threads = h->tune.rhok_threads;
    const uint32_t tiles = (N + RHOK_TILE - 1) / RHOK_TILE;
    uint32_t P = (32u * (uint32_t)h->num_sms + T - 1) / T;
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
    for (uint32_t k0 = 0; k0 < K; k0 += (uint32_t)threads)
        {
        const uint32_t KS = (K - k0) < (uint32_t)threads ? (K - k0) : (uint32_t)threads;
        const uint32_t groups = (uint32_t)threads / KS;
        const size_t smem = sizeof(double) * 2 * groups * KS;
        k_rhok<<<dim3(P, T), threads, smem, s>>>(pos, stride, frame_stride, N, kvec, k0, KS, K, target, P, direct);
        CAVB_CHECK(cudaGetLastError());
        h->launches += 1;
        }
    if (!direct)
        {
        const unsigned long long pairs = (unsigned long long)T * K;
        k_rhok_fold<<<(unsigned int)((pairs + 255) / 256), 256, 0, s>>>(target, P, K, T, rho);
        CAVB_CHECK(cudaGetLastError());
        h->launches += 1;
        }
    return 0;
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
