// hotpath.cuh -- device functions of the streaming kernels of the cavity force and the Bussi thermostat (sm_100a).
//
// Every call is "reduce, combine, apply":
//   reduce   stream pos/charge/image (52 B/particle) and/or vel (32 B) once, accumulating the dipole
//            d = sum c_i (r_i + n_i L) as compensated (hi, lo) pairs, sum m|v|^2, and the first
//            particle of type 'L' (index, unwrapped position, dipole term);
//   combine  G per-CTA records (128 B each) -> d, KE, photon, Dq, F_L, energies, alpha.  Fixed
//            order, so results are bitwise reproducible for a given launch shape;
//   apply    force_i = {(-g c_i) Dq.x, (-g c_i) Dq.y, 0, 0} (8 B read + 32 B write per particle) and
//            v_i <- alpha v_i (32 B read + 32 B write).
// Two launch shapes of the same three device functions:
//   variant 0  reduce kernel (last CTA to finish does the combine) + apply kernel
//   variant 1  one cooperative persistent kernel: reduce, grid barrier, every CTA combines the
//              same records in the same order, apply.
// The arithmetic of the per-particle terms, Dq, F_L, the energies and alpha follows the reference
// operation by operation (src/CavityForceCompute.cc:91-129,169-207 and
// src/BussiReservoirThermostat.h:177-225): explicit __dmul_rn/__dadd_rn so nothing is contracted
// into an FMA (a baseline x86-64 build of the reference has none).  Only the ORDER of the dipole
// sum differs: the reference adds index-ascending in plain doubles; here each thread adds its
// strided subset with error-free two-sums and the pairs are merged in a fixed tree, which is the
// correctly rounded sum of the same terms to ~1 ulp (SURVEY.md Appendix A bounds the reference's
// own serial-sum error at (N-1) eps sum|c_i u_i|).
#pragma once
#include "cavb200_internal.cuh"

namespace cavb
    {
struct Final
    {
    double Dq[2];
    double FL[3];
    double alpha;
    long long photon_local; // index within this call's arrays, or -1
    unsigned int n_L;
    int has_photon;
    int bussi_ok;
    };

struct __align__(16) BlockScratch
    {
    double red[32][8];
    unsigned int red_u[32];
    unsigned long long red_ull[32];
    unsigned int min_cand;
    Partial rec;
    Final fin;
    };

// ------------------------------------------------------------------------------------------
// reduce
// ------------------------------------------------------------------------------------------
struct Acc
    {
    double dhi[3], dlo[3];
    double ke;
    unsigned int cand; // first 'L' particle this thread met (local index)
    unsigned int n_L;
    };

__device__ __forceinline__ void unwrap_term(const double4& p, double c, int ix, int iy, int iz, const ForceIn& f,
                                            double u[3], double t[3])
    {
    // CavityForceCompute.cc:107-109 and :124 -- multiply, then add, each rounded
    u[0] = __dadd_rn(p.x, __dmul_rn((double)ix, f.Lx));
    u[1] = __dadd_rn(p.y, __dmul_rn((double)iy, f.Ly));
    u[2] = __dadd_rn(p.z, __dmul_rn((double)iz, f.Lz));
    t[0] = __dmul_rn(c, u[0]);
    t[1] = __dmul_rn(c, u[1]);
    t[2] = __dmul_rn(c, u[2]);
    }

__device__ __forceinline__ void take_particle(Acc& a, unsigned int i, const double4& p, double c, int ix, int iy,
                                              int iz, const ForceIn& f)
    {
    double u[3], t[3];
    unwrap_term(p, c, ix, iy, iz, f, u, t);
    const bool isL = __double2loint(p.w) == (int)f.L_typeid;
    // the first 'L' of this thread is set aside: whether it is THE photon (and so skipped by the
    // sum, CavityForceCompute.cc:122) is only known once the block has voted
    const bool aside = isL && a.cand == NO_INDEX;
    if (isL)
        {
        a.n_L++;
        if (aside)
            a.cand = i;
        }
    if (!aside)
        {
        two_sum_acc(a.dhi[0], a.dlo[0], t[0]);
        two_sum_acc(a.dhi[1], a.dlo[1], t[1]);
        two_sum_acc(a.dhi[2], a.dlo[2], t[2]);
        }
    }

template<int UNROLL> __device__ __forceinline__ void reduce_force(Acc& a, const ForceIn& f)
    {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long N = f.N;
    if (UNROLL > 1)
        {
        for (; i + (UNROLL - 1) * stride < N; i += UNROLL * stride)
            {
            double4 p[UNROLL];
            double c[UNROLL];
            int ix[UNROLL], iy[UNROLL], iz[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                {
                const unsigned long long j = i + k * stride;
                p[k] = ld256_stream(f.pos + j);
                c[k] = __ldg(f.charge + j);
                ix[k] = __ldg(f.image + 3 * j + 0);
                iy[k] = __ldg(f.image + 3 * j + 1);
                iz[k] = __ldg(f.image + 3 * j + 2);
                }
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                take_particle(a, (unsigned int)(i + k * stride), p[k], c[k], ix[k], iy[k], iz[k], f);
            }
        }
    for (; i < N; i += stride)
        {
        const double4 p = ld256_stream(f.pos + i);
        const double c = __ldg(f.charge + i);
        const int ix = __ldg(f.image + 3 * i + 0);
        const int iy = __ldg(f.image + 3 * i + 1);
        const int iz = __ldg(f.image + 3 * i + 2);
        take_particle(a, (unsigned int)i, p, c, ix, iy, iz, f);
        }
    }

template<int UNROLL> __device__ __forceinline__ void reduce_ke(Acc& a, const BussiIn& b)
    {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = b.n;
    double ke[UNROLL > 1 ? UNROLL : 1];
#pragma unroll
    for (int k = 0; k < (UNROLL > 1 ? UNROLL : 1); k++)
        ke[k] = 0.0;
    if (UNROLL > 1)
        {
        for (; j + (UNROLL - 1) * stride < n; j += UNROLL * stride)
            {
            double4 v[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                {
                const unsigned long long jj = j + k * stride;
                const unsigned long long idx = b.gidx ? (unsigned long long)__ldg(b.gidx + jj) : b.first + jj;
                v[k] = ld256(b.vel + idx);
                }
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                ke[k] += v[k].w * (v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z);
            }
        }
    for (; j < n; j += stride)
        {
        const unsigned long long idx = b.gidx ? (unsigned long long)__ldg(b.gidx + j) : b.first + j;
        const double4 v = ld256(b.vel + idx);
        ke[0] += v.w * (v.x * v.x + v.y * v.y + v.z * v.z);
        }
#pragma unroll
    for (int k = 0; k < (UNROLL > 1 ? UNROLL : 1); k++)
        a.ke += ke[k];
    }

// Block-wide merge of the per-thread accumulators into `sc.rec` (valid after the call in thread 0,
// and in shared memory for everyone after the trailing __syncthreads()).
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void block_merge(Acc& a, const ForceIn& f, BlockScratch& sc)
    {
    const unsigned int tid = threadIdx.x;
    const unsigned int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;

    if (FORCE)
        {
        if (tid == 0)
            {
            sc.min_cand = NO_INDEX;
            sc.rec.first_L = ~0ull;
            sc.rec.q[0] = sc.rec.q[1] = sc.rec.q[2] = 0.0;
            sc.rec.t[0] = sc.rec.t[1] = sc.rec.t[2] = 0.0;
            }
        __syncthreads();
        if (a.cand != NO_INDEX)
            atomicMin(&sc.min_cand, a.cand);
        __syncthreads();
        if (a.cand != NO_INDEX)
            {
            // rare path (one thread per system in practice): fetch the particle again
            const unsigned long long i = a.cand;
            const double4 p = ld256(f.pos + i);
            const double c = f.charge[i];
            double u[3], t[3];
            unwrap_term(p, c, f.image[3 * i + 0], f.image[3 * i + 1], f.image[3 * i + 2], f, u, t);
            if (a.cand == sc.min_cand)
                {
                sc.rec.first_L = f.index_offset + i;
                for (int k = 0; k < 3; k++)
                    {
                    sc.rec.q[k] = u[k];
                    sc.rec.t[k] = t[k];
                    }
                }
            else
                {
                for (int k = 0; k < 3; k++)
                    two_sum_acc(a.dhi[k], a.dlo[k], t[k]);
                }
            }
        }

    // warp tree
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        {
        if (FORCE)
            {
#pragma unroll
            for (int k = 0; k < 3; k++)
                {
                const double h2 = shfl_xor_d(a.dhi[k], m);
                const double l2 = shfl_xor_d(a.dlo[k], m);
                pair_add(a.dhi[k], a.dlo[k], h2, l2);
                }
            a.n_L += __shfl_xor_sync(0xffffffffu, a.n_L, m);
            }
        if (BUSSI)
            a.ke += shfl_xor_d(a.ke, m);
        }
    if (lane == 0)
        {
        for (int k = 0; k < 3; k++)
            {
            sc.red[warp][k] = a.dhi[k];
            sc.red[warp][3 + k] = a.dlo[k];
            }
        sc.red[warp][6] = a.ke;
        sc.red_u[warp] = a.n_L;
        }
    __syncthreads();
    if (warp == 0)
        {
        Acc b;
        for (int k = 0; k < 3; k++)
            {
            b.dhi[k] = lane < nwarps ? sc.red[lane][k] : 0.0;
            b.dlo[k] = lane < nwarps ? sc.red[lane][3 + k] : 0.0;
            }
        b.ke = lane < nwarps ? sc.red[lane][6] : 0.0;
        b.n_L = lane < nwarps ? sc.red_u[lane] : 0u;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1)
            {
            if (FORCE)
                {
#pragma unroll
                for (int k = 0; k < 3; k++)
                    {
                    const double h2 = shfl_xor_d(b.dhi[k], m);
                    const double l2 = shfl_xor_d(b.dlo[k], m);
                    pair_add(b.dhi[k], b.dlo[k], h2, l2);
                    }
                b.n_L += __shfl_xor_sync(0xffffffffu, b.n_L, m);
                }
            if (BUSSI)
                b.ke += shfl_xor_d(b.ke, m);
            }
        if (lane == 0)
            {
            for (int k = 0; k < 3; k++)
                {
                sc.rec.dhi[k] = b.dhi[k];
                sc.rec.dlo[k] = b.dlo[k];
                }
            sc.rec.ke = b.ke;
            sc.rec.n_L = b.n_L;
            sc.rec.pad = 0;
            sc.rec.pad_d = 0.0;
            if (!FORCE)
                {
                sc.rec.first_L = ~0ull;
                sc.rec.q[0] = sc.rec.q[1] = sc.rec.q[2] = 0.0;
                sc.rec.t[0] = sc.rec.t[1] = sc.rec.t[2] = 0.0;
                }
            }
        }
    __syncthreads();
    }

template<bool FORCE, bool BUSSI, int UNROLL>
__device__ __forceinline__ void reduce_phase(const ForceIn& f, const BussiIn& b, BlockScratch& sc)
    {
    Acc a;
    for (int k = 0; k < 3; k++)
        a.dhi[k] = a.dlo[k] = 0.0;
    a.ke = 0.0;
    a.cand = NO_INDEX;
    a.n_L = 0;
    if (FORCE)
        reduce_force<UNROLL>(a, f);
    if (BUSSI)
        reduce_ke<UNROLL>(a, b);
    block_merge<FORCE, BUSSI>(a, f, sc);
    }

__device__ __forceinline__ void store_record(Partial* dst, const Partial& src)
    {
    // 128 B as eight 16-B stores by the first 8 threads
    const double2* s = reinterpret_cast<const double2*>(&src);
    double2* d = reinterpret_cast<double2*>(dst);
    if (threadIdx.x < 8)
        __stcg(d + threadIdx.x, s[threadIdx.x]);
    }

// ------------------------------------------------------------------------------------------
// combine = merge (G records -> one record, block-wide, fixed order) + finalize (thread 0)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double bussi_alpha(double KE, const BussiIn& b, int& ok)
    {
    // BussiReservoirThermostat.h:177-225 with c = exp(-dt/tau) evaluated by the host
    ok = 1;
    if (b.dof == 0.0)
        return 1.0;
    if (KE == 0.0)
        {
        ok = 0; // the reference throws (:57-61)
        return 1.0;
        }
    const double c = b.c, R = b.r_normal;
    const double omc = __dadd_rn(1.0, -c);
    const double v = __ddiv_rn(__ddiv_rn(b.kT, 2.0), KE);
    const double vomc = __dmul_rn(v, omc);
    const double term1 = __dmul_rn(vomc, __dadd_rn(b.r_gamma, __dmul_rn(R, R)));
    const double term2 = __dmul_rn(__dmul_rn(2.0, R), __dsqrt_rn(__dmul_rn(vomc, c)));
    const double alpha2 = __dadd_rn(__dadd_rn(c, term1), term2);
    const double mag = __dsqrt_rn(alpha2);
    const double K_bar = __ddiv_rn(__dmul_rn(b.kT, b.dof), 2.0);
    const double sign_term
        = __dadd_rn(R, __dsqrt_rn(__ddiv_rn(__dmul_rn(__dmul_rn(c, b.dof), KE), __dmul_rn(omc, K_bar))));
    return sign_term >= 0.0 ? mag : -mag;
    }

// Merge G records (read through L2) into sc.rec.  Every field of sc.rec is valid for ALL threads
// after the call.  The winner of the first-'L' vote keeps q and t; the losers' terms go back into d
// (they are ordinary members of the dipole sum, CavityForceCompute.cc:120-126).
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void merge_phase(const Partial* __restrict__ recs, int G, BlockScratch& sc)
    {
    const unsigned int tid = threadIdx.x;
    const unsigned int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;

    // pass 1: global first 'L'
    unsigned long long gmin = ~0ull;
    if (FORCE)
        {
        for (int r = tid; r < G; r += blockDim.x)
            {
            const unsigned long long v = __ldcg(&recs[r].first_L);
            gmin = v < gmin ? v : gmin;
            }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1)
            {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, gmin, m);
            gmin = o < gmin ? o : gmin;
            }
        __syncthreads(); // sc.rec may still be being read by store_record of the caller
        if (lane == 0)
            sc.red_ull[warp] = gmin;
        if (tid == 0)
            {
            sc.rec.q[0] = sc.rec.q[1] = sc.rec.q[2] = 0.0;
            sc.rec.t[0] = sc.rec.t[1] = sc.rec.t[2] = 0.0;
            }
        __syncthreads();
        gmin = ~0ull;
        for (unsigned int w = 0; w < nwarps; w++)
            gmin = sc.red_ull[w] < gmin ? sc.red_ull[w] : gmin;
        }
    else
        __syncthreads();

    // pass 2: pairs in record order
    Acc a;
    for (int k = 0; k < 3; k++)
        a.dhi[k] = a.dlo[k] = 0.0;
    a.ke = 0.0;
    a.n_L = 0;
    for (int r = tid; r < G; r += blockDim.x)
        {
        // record layout in 16-byte words: {dhi0,dhi1} {dhi2,dlo0} {dlo1,dlo2} {ke,q0} {q1,q2} {t0,t1}
        // {t2,pad} {first_L, n_L|pad}
        const double2* p = reinterpret_cast<const double2*>(recs + r);
        const double2 v3 = __ldcg(p + 3);
        if (FORCE)
            {
            const double2 v0 = __ldcg(p + 0), v1 = __ldcg(p + 1), v2 = __ldcg(p + 2);
            pair_add(a.dhi[0], a.dlo[0], v0.x, v1.y);
            pair_add(a.dhi[1], a.dlo[1], v0.y, v2.x);
            pair_add(a.dhi[2], a.dlo[2], v1.x, v2.y);
            const ulonglong2 v7 = __ldcg(reinterpret_cast<const ulonglong2*>(p + 7));
            a.n_L += (unsigned int)(v7.y & 0xffffffffull);
            if (v7.x != ~0ull)
                {
                const double2 v4 = __ldcg(p + 4), v5 = __ldcg(p + 5), v6 = __ldcg(p + 6);
                if (v7.x == gmin)
                    {
                    sc.rec.q[0] = v3.y;
                    sc.rec.q[1] = v4.x;
                    sc.rec.q[2] = v4.y;
                    sc.rec.t[0] = v5.x;
                    sc.rec.t[1] = v5.y;
                    sc.rec.t[2] = v6.x;
                    }
                else
                    {
                    two_sum_acc(a.dhi[0], a.dlo[0], v5.x);
                    two_sum_acc(a.dhi[1], a.dlo[1], v5.y);
                    two_sum_acc(a.dhi[2], a.dlo[2], v6.x);
                    }
                }
            }
        if (BUSSI)
            a.ke += v3.x;
        }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        {
        if (FORCE)
            {
#pragma unroll
            for (int k = 0; k < 3; k++)
                {
                const double h2 = shfl_xor_d(a.dhi[k], m);
                const double l2 = shfl_xor_d(a.dlo[k], m);
                pair_add(a.dhi[k], a.dlo[k], h2, l2);
                }
            a.n_L += __shfl_xor_sync(0xffffffffu, a.n_L, m);
            }
        if (BUSSI)
            a.ke += shfl_xor_d(a.ke, m);
        }
    if (lane == 0)
        {
        for (int k = 0; k < 3; k++)
            {
            sc.red[warp][k] = a.dhi[k];
            sc.red[warp][3 + k] = a.dlo[k];
            }
        sc.red[warp][6] = a.ke;
        sc.red_u[warp] = a.n_L;
        }
    __syncthreads();
    if (tid == 0)
        {
        double dh[3] = {0, 0, 0}, dl[3] = {0, 0, 0}, kes = 0.0;
        unsigned int nL = 0;
        for (unsigned int w = 0; w < nwarps; w++)
            {
            for (int k = 0; k < 3; k++)
                pair_add(dh[k], dl[k], sc.red[w][k], sc.red[w][3 + k]);
            kes += sc.red[w][6];
            nL += sc.red_u[w];
            }
        for (int k = 0; k < 3; k++)
            {
            sc.rec.dhi[k] = dh[k];
            sc.rec.dlo[k] = dl[k];
            }
        sc.rec.ke = kes;
        sc.rec.n_L = nL;
        sc.rec.pad = 0;
        sc.rec.pad_d = 0.0;
        sc.rec.first_L = gmin;
        }
    __syncthreads();
    }

// thread 0 only: merged record -> Final (shared) and, when `publish`, the device-resident Scalars
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void finalize(const ForceIn& f, const BussiIn& b, BlockScratch& sc, Scalars* scalars,
                                         bool publish, int barrier_timeout)
    {
    const Partial& m = sc.rec;
    Final fin;
    fin.n_L = m.n_L;
    fin.has_photon = 0;
    fin.photon_local = -1;
    fin.Dq[0] = fin.Dq[1] = 0.0;
    fin.FL[0] = fin.FL[1] = fin.FL[2] = 0.0;
    fin.alpha = 1.0;
    fin.bussi_ok = 1;
    double en[3] = {0, 0, 0}, d[3] = {0, 0, 0}, q[3] = {0, 0, 0};
    if (FORCE && m.first_L != ~0ull)
        {
        fin.has_photon = 1;
        const unsigned long long lo = f.index_offset;
        if (m.first_L >= lo && m.first_L < lo + f.N)
            fin.photon_local = (long long)(m.first_L - lo);
        for (int k = 0; k < 3; k++)
            {
            d[k] = __dadd_rn(m.dhi[k], m.dlo[k]);
            q[k] = m.q[k];
            }
        const double g = f.g, K = f.K;
        // CavityForceCompute.cc:174-176
        const double qq = __dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2]));
        const double dq = __dadd_rn(__dadd_rn(__dmul_rn(d[0], q[0]), __dmul_rn(d[1], q[1])), 0.0);
        const double dd = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), 0.0);
        en[0] = __dmul_rn(__dmul_rn(0.5, K), qq);
        en[1] = __dmul_rn(g, dq);
        en[2] = __dmul_rn(__dmul_rn(0.5, __ddiv_rn(__dmul_rn(g, g), K)), dd);
        // :183
        const double gk = __ddiv_rn(g, K);
        fin.Dq[0] = __dadd_rn(q[0], __dmul_rn(gk, d[0]));
        fin.Dq[1] = __dadd_rn(q[1], __dmul_rn(gk, d[1]));
        // :203
        fin.FL[0] = __dadd_rn(__dmul_rn(-K, q[0]), -__dmul_rn(g, d[0]));
        fin.FL[1] = __dadd_rn(__dmul_rn(-K, q[1]), -__dmul_rn(g, d[1]));
        fin.FL[2] = __dadd_rn(__dmul_rn(-K, q[2]), -__dmul_rn(g, 0.0));
        }
    double KE = 0.0, inst = 0.0;
    if (BUSSI)
        {
        KE = __dmul_rn(0.5, m.ke);
        if (b.rescale)
            {
            fin.alpha = bussi_alpha(KE, b, fin.bussi_ok);
            // BussiReservoirThermostat.h:86
            inst = __dmul_rn(KE, __dadd_rn(1.0, -__dmul_rn(fin.alpha, fin.alpha)));
            }
        }
    sc.fin = fin;
    if (publish)
        {
        if (FORCE)
            {
            for (int k = 0; k < 3; k++)
                {
                scalars->energies[k] = en[k];
                scalars->dipole[k] = d[k];
                scalars->q[k] = q[k];
                scalars->FL[k] = fin.FL[k];
                }
            scalars->Dq[0] = fin.Dq[0];
            scalars->Dq[1] = fin.Dq[1];
            scalars->photon_idx = fin.has_photon ? (long long)m.first_L : -1;
            scalars->n_L = m.n_L;
            }
        if (BUSSI)
            {
            scalars->ke = KE;
            if (b.rescale)
                {
                scalars->alpha = fin.alpha;
                scalars->inst = inst;
                scalars->cumulative = __dadd_rn(scalars->cumulative, inst); // :90
                if (!fin.bussi_ok)
                    scalars->err = 1.0;
                }
            }
        if (barrier_timeout)
            scalars->err = 2.0;
        }
    }

template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void combine_phase(const Partial* __restrict__ recs, int G, const ForceIn& f,
                                              const BussiIn& b, BlockScratch& sc, Scalars* scalars, bool publish,
                                              int barrier_timeout)
    {
    merge_phase<FORCE, BUSSI>(recs, G, sc);
    if (threadIdx.x == 0)
        finalize<FORCE, BUSSI>(f, b, sc, scalars, publish, barrier_timeout);
    __syncthreads();
    }

// ------------------------------------------------------------------------------------------
// apply
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 force_of(unsigned long long i, double c, const Final& fin, const ForceIn& f)
    {
    double4 o = make_double4(0.0, 0.0, 0.0, 0.0);
    if (!fin.has_photon)
        return o; // CavityForceCompute.cc:149-156
    bool isL;
    if (fin.n_L <= 1)
        isL = (long long)i == fin.photon_local;
    else // several 'L' particles: the type has to be looked at again (32 B/particle, never in practice)
        isL = __double2loint(f.pos[i].w) == (int)f.L_typeid;
    if (!isL)
        {
        const double s = __dmul_rn(-f.g, c); // :193
        o.x = __dmul_rn(s, fin.Dq[0]);
        o.y = __dmul_rn(s, fin.Dq[1]);
        }
    else if ((long long)i == fin.photon_local)
        {
        o.x = fin.FL[0];
        o.y = fin.FL[1];
        o.z = fin.FL[2];
        }
    return o;
    }

template<int UNROLL> __device__ __forceinline__ void apply_force(const Final& fin, const ForceIn& f)
    {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long N = f.N;
    if (UNROLL > 1)
        {
        for (; i + (UNROLL - 1) * stride < N; i += UNROLL * stride)
            {
            double c[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                c[k] = __ldg(f.charge + i + k * stride);
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                st256(f.force + i + k * stride, force_of(i + k * stride, c[k], fin, f));
            }
        }
    for (; i < N; i += stride)
        st256(f.force + i, force_of(i, __ldg(f.charge + i), fin, f));
    }

template<int UNROLL> __device__ __forceinline__ void apply_rescale(double alpha, const BussiIn& b)
    {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = b.n;
    if (UNROLL > 1)
        {
        for (; j + (UNROLL - 1) * stride < n; j += UNROLL * stride)
            {
            double4 v[UNROLL];
            unsigned long long idx[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                {
                const unsigned long long jj = j + k * stride;
                idx[k] = b.gidx ? (unsigned long long)__ldg(b.gidx + jj) : b.first + jj;
                v[k] = ld256(b.vel + idx[k]);
                }
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                {
                v[k].x = __dmul_rn(v[k].x, alpha);
                v[k].y = __dmul_rn(v[k].y, alpha);
                v[k].z = __dmul_rn(v[k].z, alpha);
                st256(b.vel + idx[k], v[k]);
                }
            }
        }
    for (; j < n; j += stride)
        {
        const unsigned long long idx = b.gidx ? (unsigned long long)__ldg(b.gidx + j) : b.first + j;
        double4 v = ld256(b.vel + idx);
        v.x = __dmul_rn(v.x, alpha);
        v.y = __dmul_rn(v.y, alpha);
        v.z = __dmul_rn(v.z, alpha);
        st256(b.vel + idx, v);
        }
    }

template<bool FORCE, bool BUSSI, int UNROLL>
__device__ __forceinline__ void apply_phase(const Final& fin, const ForceIn& f, const BussiIn& b)
    {
    if (FORCE)
        apply_force<UNROLL>(fin, f);
    if (BUSSI && b.rescale && fin.bussi_ok && fin.alpha != 1.0)
        apply_rescale<UNROLL>(fin.alpha, b);
    }

    } // namespace cavb
