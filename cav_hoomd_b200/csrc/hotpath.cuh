// hotpath.cuh -- device functions of the streaming kernels of the cavity force and the Bussi
// thermostat (sm_100a).
//
// Every call is "reduce, combine, apply":
//   reduce   stream pos/charge/image (52 B/particle) and/or vel (32 B) once, accumulating the dipole
//            d = sum c_i (r_i + n_i L) as compensated (hi, lo) pairs, sum m|v|^2, and the first
//            particle of type 'L' (index, unwrapped position, dipole term);
//   combine  G per-CTA records -> d, KE, photon, Dq, F_L, energies, alpha.  Fixed order, so results
//            are bitwise reproducible for a given launch shape;
//   apply    force_i = {(-g c_i) Dq.x, (-g c_i) Dq.y, 0, 0} (8 B read + 32 B write per particle) and
//            v_i <- alpha v_i (32 B read + 32 B write).
// Launch shapes built from these functions (hotpath.cu, shard.cu):
//   variant 1  ONE cooperative persistent kernel: reduce, grid-wide hand-off, every CTA combines the
//              same records in the same order, apply.  The hand-off has no counter and no atomic:
//              each CTA publishes its record and then an epoch flag inside the record; thread j of
//              every CTA polls record j's flag and reads the record the moment it is there.
//   variant 0  reduce kernel (last CTA to take a ticket combines) + apply kernel.
//
// Why the combine looks the way it does (profiles/ round-1 notes): with one 128-byte record per CTA
// read as eight 16-byte requests by all 296 CTAs, the hand-off took 5.5 us -- 700 k requests
// hammering 296 L2 lines.  Records are therefore laid out in 32-byte sectors read with LDG.256 and
// the common path touches three sectors.
//
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
    int many_L;             // more than one particle of type 'L' exists
    int has_photon;
    int bussi_ok;
    int timeout;
    };

struct __align__(32) BlockScratch
    {
    double red[8][32]; // [component][warp]: conflict-free when lane == warp index
    unsigned long long red_ull[32];
    unsigned int red_cnt[32]; // per-warp count of records holding an 'L' candidate (bit 31: multi flag)
    unsigned int red_nl[32];  // per-warp number of 'L' particles
    double wq[32][3];         // per-warp: q of the warp's best candidate
    double wt[32][3];         // per-warp: its dipole term (kept only by rank-level merges)
    unsigned int min_cand;
    unsigned int flags; // bit 0: hand-off timed out
    unsigned long long epoch;
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

__device__ __forceinline__ void acc_zero(Acc& a)
    {
#pragma unroll
    for (int k = 0; k < 3; k++)
        a.dhi[k] = a.dlo[k] = 0.0;
    a.ke = 0.0;
    a.cand = NO_INDEX;
    a.n_L = 0;
    }

// The trees.  A thread's Acc holds four (hi, lo) pairs worth summing -- d.x, d.y, d.z and (KE, nothing) -- and a
// butterfly over all four costs 20 pair additions and 40 64-bit shuffles per warp, every lane computing every total.
// With 24 warps of an SM arriving at the end of the reduce together that is ~3000 DADD and ~2000 SHFL warp instructions
// queued on two pipes: measured 0.8 us of the 21.6 us force call for the z component alone (profiles/tree4_r2a.txt).
// So the tree is TRANSPOSED: after the first exchange (lanes l and l^16) a lane carries on with two of the four pairs,
// after the second (l^8) with one, and lanes 8s..8s+7 end up holding the warp's total of pair s (Slot): 6 pair additions
// and 12 shuffles.  The pairings (16, 8, 4, 2, 1) are the butterfly's, so the plain-double KE -- the hi word of pair 3,
// whose lo word is ignored -- keeps the bits of a plain butterfly sum, which is what the Bussi-only kernels use.
// (Measured before: the whole-grid fold of k_fused takes 4.25 us the first time through and 2.25 us when it is run a
// second time in the same launch; rolling the level loop to shrink the code -- 5664 -> 3208 instructions --
// made it slower, 23.7 -> 24.95 us per force call, so instruction fetch is not what the first pass pays for.)
struct Slot
    {
    double hi, lo; // lanes 8s..8s+7: total of pair s (0..2: dipole x, y, z; 3: KE in hi).  !FORCE: KE in hi, every lane
    };

__device__ __forceinline__ void slot_butterfly(Slot& v, int m)
    {
    const double h2 = shfl_xor_d(v.hi, m), l2 = shfl_xor_d(v.lo, m);
    pair_add(v.hi, v.lo, h2, l2);
    }

// warp total; a.n_L is summed in place (every lane)
template<bool FORCE, bool BUSSI> __device__ __forceinline__ Slot warp_tree(Acc& a, unsigned int lane)
    {
    Slot v;
    if (FORCE)
        {
        const bool up = (lane & 16u) != 0; // upper half keeps (z, KE) and sends (x, y)
        const double ke = BUSSI ? a.ke : 0.0;
        double k0h = up ? a.dhi[2] : a.dhi[0], k0l = up ? a.dlo[2] : a.dlo[0];
        double k1h = up ? ke : a.dhi[1], k1l = up ? 0.0 : a.dlo[1];
        const double s0h = up ? a.dhi[0] : a.dhi[2], s0l = up ? a.dlo[0] : a.dlo[2];
        const double s1h = up ? a.dhi[1] : ke, s1l = up ? a.dlo[1] : 0.0;
        const double r0h = shfl_xor_d(s0h, 16), r0l = shfl_xor_d(s0l, 16);
        const double r1h = shfl_xor_d(s1h, 16), r1l = shfl_xor_d(s1l, 16);
        pair_add(k0h, k0l, r0h, r0l);
        pair_add(k1h, k1l, r1h, r1l);
        const bool up8 = (lane & 8u) != 0; // of its two pairs a lane keeps the first (bit 3 clear) or the second
        v.hi = up8 ? k1h : k0h;
        v.lo = up8 ? k1l : k0l;
        const double sh = up8 ? k0h : k1h, sl = up8 ? k0l : k1l;
        const double rh = shfl_xor_d(sh, 8), rl = shfl_xor_d(sl, 8);
        pair_add(v.hi, v.lo, rh, rl);
        slot_butterfly(v, 4);
        slot_butterfly(v, 2);
        slot_butterfly(v, 1);
        a.n_L = __reduce_add_sync(0xffffffffu, a.n_L);
        }
    else
        {
        v.hi = a.ke;
        v.lo = 0.0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1)
            v.hi += shfl_xor_d(v.hi, m);
        }
    return v;
    }

// Second tree level: the warps' totals go through shared memory, pair s in row s (hi) and 4 + s (lo; row 7 unused),
// columns skewed by 8 per row so that the read below is conflict-free.
__device__ __forceinline__ unsigned int red_col(unsigned int slot, unsigned int warp) { return (warp + 8u * slot) & 31u; }

template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void cross_warp_put(const Slot& v, unsigned int n_L, BlockScratch& sc, unsigned int lane,
                                               unsigned int warp)
    {
    if (FORCE)
        {
        if ((lane & 7u) == 0)
            {
            const unsigned int slot = lane >> 3;
            sc.red[slot][red_col(slot, warp)] = v.hi;
            if (slot < 3)
                sc.red[4 + slot][red_col(slot, warp)] = v.lo;
            }
        if (lane == 0)
            sc.red_nl[warp] = n_L;
        }
    else if (lane == 0)
        sc.red[3][red_col(3, warp)] = v.hi;
    }
// warp 0 only.  Lane 8s + j sums pair s of warps j, j + 8, j + 16, j + 24 -- paired (j, j+16), (j+8, j+24), then those two,
// the butterfly's order again -- and three exchanges finish it.  Returns the CTA totals as a Slot; n_L in every lane.
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ Slot cross_warp_get(unsigned int& n_L, const BlockScratch& sc, unsigned int lane,
                                               unsigned int nwarps)
    {
    Slot v;
    n_L = 0;
    if (FORCE)
        {
        const unsigned int slot = lane >> 3, j = lane & 7u;
        Slot w[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            {
            const unsigned int wi = j + 8u * c;
            const bool have = wi < nwarps;
            w[c].hi = have ? sc.red[slot][red_col(slot, wi)] : 0.0;
            w[c].lo = (have && slot < 3) ? sc.red[4 + slot][red_col(slot, wi)] : 0.0;
            }
        if (nwarps > 16)
            {
            pair_add(w[0].hi, w[0].lo, w[2].hi, w[2].lo);
            pair_add(w[1].hi, w[1].lo, w[3].hi, w[3].lo);
            }
        if (nwarps > 8)
            pair_add(w[0].hi, w[0].lo, w[1].hi, w[1].lo);
        v = w[0];
        slot_butterfly(v, 4);
        slot_butterfly(v, 2);
        slot_butterfly(v, 1);
        n_L = __reduce_add_sync(0xffffffffu, lane < nwarps ? sc.red_nl[lane] : 0u);
        }
    else
        {
        v.hi = lane < nwarps ? sc.red[3][red_col(3, lane)] : 0.0;
        v.lo = 0.0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1)
            v.hi += shfl_xor_d(v.hi, m);
        }
    return v;
    }
// the lanes that hold a total write it into the merged record (lane 0 of each group of eight; KE: lane 24, or lane 0 of
// a Bussi-only fold); a __syncwarp() must follow before another lane reads sc.rec
template<bool FORCE, bool BUSSI> __device__ __forceinline__ void slot_to_record(const Slot& v, Partial& rec, unsigned int lane)
    {
    if (FORCE)
        {
        if ((lane & 7u) == 0)
            {
            const unsigned int slot = lane >> 3;
            if (slot < 3)
                {
                rec.dhi[slot] = v.hi;
                rec.dlo[slot] = v.lo;
                }
            else
                rec.ke = BUSSI ? v.hi : 0.0;
            }
        }
    else if (lane == 0)
        {
        for (int k = 0; k < 3; k++)
            rec.dhi[k] = rec.dlo[k] = 0.0;
        rec.ke = v.hi;
        }
    }

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

// Which CTAs stream particles: all of the grid (default), or a sub-range when some CTAs have another
// role (the folder CTA of k_split_folder).
struct StreamGrid
    {
    unsigned int nblk, blk;
    };
__device__ __forceinline__ StreamGrid whole_grid()
    {
    StreamGrid g;
    g.nblk = gridDim.x;
    g.blk = blockIdx.x;
    return g;
    }

// Kinetic-energy pass over a group given as an index list (what HOOMD's ParticleGroup hands out) or as a window
// [first, first + n).  With a list, element j needs gidx[j] before its velocity can be fetched: two dependent memory
// round trips per loop iteration as first written (18.0 vs 16.1 us per 1M-particle Bussi call, profiles/ab_r2b.txt).
// The index loads are therefore software-pipelined one iteration ahead -- the indices of iteration i+1 are requested
// together with the velocities of iteration i -- so only the first iteration pays the second round trip.
// (The list loops are compiled only into the kernel instantiations that take an index list -- template parameter LIST,
// the Bussi-only kernels of cavb200_bussi / cavb200_bussi_ke.  Sharing one kernel with the window path changed the
// register allocation of the persistent kernels around the loops, which sit exactly at their 80-register budget, and
// cost the contiguous step 0.3 us at 1M particles; as non-inlined functions they miscompiled, nvcc 12.9: wrong alpha
// from the second call on in tests/test_bussi_gpu.py::test_bussi_index_list_group_and_cumulative.)
template<int UNROLL, bool LIST>
__device__ __forceinline__ void reduce_ke(Acc& a, const BussiIn& b, const StreamGrid g = whole_grid())
    {
    constexpr int U = UNROLL > 1 ? UNROLL : 1;
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    unsigned long long j = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
    const unsigned long long n = b.n;
    double ke[U];
#pragma unroll
    for (int k = 0; k < U; k++)
        ke[k] = 0.0;
    if (LIST)
        {
        unsigned int nxt[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            nxt[k] = (j + k * stride < n) ? __ldg(b.gidx + j + k * stride) : 0u;
        for (; j < n; j += U * stride)
            {
            double4 v[U];
            unsigned int cur[U];
#pragma unroll
            for (int k = 0; k < U; k++)
                {
                cur[k] = nxt[k];
                v[k] = make_double4(0.0, 0.0, 0.0, 0.0);
                if (j + k * stride < n)
                    v[k] = ld256(b.vel + cur[k]);
                }
#pragma unroll
            for (int k = 0; k < U; k++)
                {
                const unsigned long long jn = j + (U + k) * stride;
                nxt[k] = jn < n ? __ldg(b.gidx + jn) : 0u;
                }
#pragma unroll
            for (int k = 0; k < U; k++)
                ke[k] += v[k].w * (v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z); // (+0 for the padding slots)
            }
        }
    else
        {
        for (; j + (U - 1) * stride < n; j += U * stride)
            {
            double4 v[U];
#pragma unroll
            for (int k = 0; k < U; k++)
                v[k] = ld256(b.vel + b.first + j + k * stride);
#pragma unroll
            for (int k = 0; k < U; k++)
                ke[k] += v[k].w * (v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z);
            }
        for (; j < n; j += stride)
            {
            const double4 v = ld256(b.vel + b.first + j);
            ke[0] += v.w * (v.x * v.x + v.y * v.y + v.z * v.z);
            }
        }
#pragma unroll
    for (int k = 0; k < U; k++)
        a.ke += ke[k];
    }

// ---- fused stream (contiguous Bussi group starting at particle 0) ------------------------------
// ONE loop reads pos/charge/image AND vel.  Measured with tools/microstream.cu
// (profiles/microstream_r1b.txt): this access pattern plus this arithmetic streams 84 MB in 13.1 us
// (6.4 TB/s) at 1024 threads/SM with two particles in flight per thread, but only 4.4 TB/s at 512
// and 2.1 TB/s at 256 threads/SM -- the per-particle FP64 work needs warps, not unrolling, to hide.
// Separate force and KE loops (the first version) paid the memory ramp twice.  A software-pipelined
// variant with predicated loads was slower than this plain "U loads, U consumes" loop and was dropped.
template<bool FORCE, bool KE, int U>
__device__ __forceinline__ void reduce_stream(Acc& a, const ForceIn& f, const BussiIn& b, const StreamGrid g = whole_grid())
    {
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    const unsigned long long nf = FORCE ? (unsigned long long)f.N : 0ull;
    const unsigned long long nk = KE ? (unsigned long long)b.n : 0ull; // group = [0, n)
    const unsigned long long common = (FORCE && KE) ? (nf < nk ? nf : nk) : (FORCE ? nf : nk);
    const unsigned long long all = nf > nk ? nf : nk;
    unsigned long long i = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
    double ke[U];
#pragma unroll
    for (int k = 0; k < U; k++)
        ke[k] = 0.0;
    for (; i + (U - 1) * stride < common; i += U * stride)
        {
        double4 p[FORCE ? U : 1], v[KE ? U : 1];
        double c[FORCE ? U : 1];
        int ix[FORCE ? U : 1], iy[FORCE ? U : 1], iz[FORCE ? U : 1];
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            const unsigned long long j = i + k * stride;
            if (FORCE)
                {
                p[k] = ld256_stream(f.pos + j);
                c[k] = __ldg(f.charge + j);
                ix[k] = ld_image(f.image + 3 * j + 0);
                iy[k] = ld_image(f.image + 3 * j + 1);
                iz[k] = ld_image(f.image + 3 * j + 2);
                }
            if (KE)
                v[k] = ld256_na(b.vel + j);
            }
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            if (FORCE)
                take_particle(a, (unsigned int)(i + k * stride), p[k], c[k], ix[k], iy[k], iz[k], f);
            if (KE)
                ke[k] += v[k].w * (v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z);
            }
        }
    // tail: fewer than U strides left, and the indices only one of the two ranges covers
    for (; i < all; i += stride)
        {
        if (FORCE && i < nf)
            {
            const double4 p = ld256_stream(f.pos + i);
            const double c = __ldg(f.charge + i);
            const int ix = ld_image(f.image + 3 * i + 0);
            const int iy = ld_image(f.image + 3 * i + 1);
            const int iz = ld_image(f.image + 3 * i + 2);
            take_particle(a, (unsigned int)i, p, c, ix, iy, iz, f);
            }
        if (KE && i < nk)
            {
            const double4 v = ld256_na(b.vel + i);
            ke[0] += v.w * (v.x * v.x + v.y * v.y + v.z * v.z);
            }
        }
    if (KE)
        {
#pragma unroll
        for (int k = 0; k < U; k++)
            a.ke += ke[k];
        }
    }

// Block-wide merge of the per-thread accumulators into sc.rec (shared; complete after the
// trailing __syncthreads()).
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void block_merge(Acc& a, const ForceIn& f, BlockScratch& sc)
    {
    const unsigned int tid = threadIdx.x;
    const unsigned int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;

    if (tid == 0)
        {
        sc.min_cand = NO_INDEX;
        sc.rec.first_L = ~0ull;
        sc.rec.q[0] = sc.rec.q[1] = sc.rec.q[2] = 0.0;
        sc.rec.t[0] = sc.rec.t[1] = sc.rec.t[2] = 0.0;
        }
    if (FORCE)
        {
        // the vote is only needed when somebody in the block met an 'L' particle
        const int any = __syncthreads_or(a.cand != NO_INDEX);
        if (any)
            {
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
        }
    const Slot v = warp_tree<FORCE, BUSSI>(a, lane);
    cross_warp_put<FORCE, BUSSI>(v, a.n_L, sc, lane, warp);
    __syncthreads();
    if (warp == 0)
        {
        unsigned int n_L;
        const Slot b = cross_warp_get<FORCE, BUSSI>(n_L, sc, lane, nwarps);
        slot_to_record<FORCE, BUSSI>(b, sc.rec, lane);
        if (lane == 0)
            {
            sc.rec.n_L = n_L;
            if (n_L > 1 && sc.rec.first_L != ~0ull)
                sc.rec.first_L |= MULTI_L_BIT;
            }
        }
    __syncthreads();
    }

template<bool FORCE, bool BUSSI, int UNROLL, bool LIST = false>
__device__ __forceinline__ void reduce_loops(Acc& a, const ForceIn& f, const BussiIn& b, const StreamGrid g = whole_grid())
    {
    acc_zero(a);
    if (BUSSI && (LIST || b.first != 0))
        {
        // index-list group or a group that does not start at particle 0: the velocity pass cannot
        // share the particle loop
        if (FORCE)
            reduce_stream<true, false, UNROLL>(a, f, b, g);
        reduce_ke<UNROLL, LIST>(a, b, g);
        }
    else
        reduce_stream<FORCE, BUSSI, UNROLL>(a, f, b, g);
    }

template<bool FORCE, bool BUSSI, int UNROLL, bool LIST = false>
__device__ __forceinline__ void reduce_phase(const ForceIn& f, const BussiIn& b, BlockScratch& sc,
                                             const StreamGrid g = whole_grid())
    {
    Acc a;
    reduce_loops<FORCE, BUSSI, UNROLL, LIST>(a, f, b, g);
    block_merge<FORCE, BUSSI>(a, f, sc);
    }

// ---- record sectors: payload x3 + check word --------------------------------------------------
__device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ double bitsd(unsigned long long x) { return __longlong_as_double((long long)x); }
__device__ __forceinline__ double4 make_sector(double a, double b, double c, unsigned long long epoch)
    {
    return make_double4(a, b, c, bitsd(epoch ^ dbits(a) ^ dbits(b) ^ dbits(c)));
    }
__device__ __forceinline__ bool sector_ok(const double4& s, unsigned long long epoch)
    {
    return (dbits(s.x) ^ dbits(s.y) ^ dbits(s.z) ^ dbits(s.w)) == epoch;
    }

// ONE thread publishes the record: five STG.256, no fence, no flag (see the Partial comment).
__device__ __forceinline__ void publish_record(Partial* dst, const Partial& src, unsigned long long epoch)
    {
    double4* d = reinterpret_cast<double4*>(dst);
    st256(d + 0, make_sector(src.dhi[0], src.dhi[1], src.dhi[2], epoch));
    st256(d + 1, make_sector(src.dlo[0], src.dlo[1], src.dlo[2], epoch));
    st256(d + 2, make_sector(src.ke, bitsd(src.first_L), bitsd(src.n_L), epoch));
    st256(d + 3, make_sector(src.q[0], src.q[1], src.q[2], epoch));
    st256(d + 4, make_sector(src.t[0], src.t[1], src.t[2], epoch));
    }

// Strong (L2) read of one sector; WAIT: poll until it carries this launch's epoch (bounded).
// SYS: the records are written by PEER GPUs over NVLink into this GPU's memory (sharded mode):
// system-scope loads, and a timeout that tolerates ranks starting late.
// Sleep between two polls of a sector that has not arrived (A/B on one box, 1M particles, profiles/poll_r1a.txt):
// whole-grid folds (k_fused, k_split) want it short -- 64 -> 16 ns: Bussi call 17.7 -> 16.7 us, force call
// 23.6 -> 23.35 us (0 ns is no better); the streaming CTAs of k_split_folder waiting for a Final record want it
// long -- every warp of the grid polls the same sector: 0 / 64 / 300 ns give 31.88 / 31.87 / 31.63 us per step.
// (Round 2, with the L2 policies of profiles/cache_policy_r2a.txt: 100 / 300 / 600 ns give 29.91 / 29.85 / 29.84 us; letting
// ONE warp per CTA poll Final(K) with a 16 or 64 ns sleep and hand alpha on through shared memory: 30.02 us.)
#ifndef CAVB_POLL_NS
#define CAVB_POLL_NS 16
#endif
#ifndef CAVB_FINAL_POLL_NS
#define CAVB_FINAL_POLL_NS 300
#endif
template<bool SYS> __device__ __forceinline__ double4 ld_rec(const double4* p) { return SYS ? ld256_sys(p) : ld256_cg(p); }
// (Looking at the timeout clock only every 16th or 256th poll was measured: force call 22.1 -> 22.5 us, no gain elsewhere.)
template<bool WAIT, bool SYS = false, int POLL_NS = CAVB_POLL_NS>
__device__ __forceinline__ double4 read_sector(const double4* p, unsigned long long epoch, bool& late)
    {
    double4 s = ld_rec<SYS>(p);
    if (WAIT)
        {
        if (!sector_ok(s, epoch))
            {
            const unsigned long long t0 = globaltimer_ns();
            do
                {
                if (POLL_NS > 0)
                    __nanosleep(POLL_NS);
                s = ld_rec<SYS>(p);
                if (globaltimer_ns() - t0 > (SYS ? PEER_TIMEOUT_NS : HANDOFF_TIMEOUT_NS)) // never hang the GPU
                    {
                    late = true;
                    break;
                    }
                } while (!sector_ok(s, epoch));
            }
        }
    return s;
    }

// ------------------------------------------------------------------------------------------
// combine = merge (G records -> one record, fixed order) + finalize
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double bussi_alpha(double KE, const BussiIn& b, int& ok)
    {
    // BussiReservoirThermostat.h:177-225.  Everything that does not depend on KE was evaluated by
    // the host in the same operation order (api.cu fill_bussi_constants): c = exp(-dt/tau),
    // omc = 1 - c, half_kT = kT/2, gR2 = r_gamma + R*R, two_R = 2*R, cdof = c*dof, den = omc*(kT*dof/2).
    ok = 1;
    if (b.dof == 0.0)
        return 1.0;
    if (KE == 0.0)
        {
        ok = 0; // the reference throws (:57-61)
        return 1.0;
        }
    const double v = __ddiv_rn(b.half_kT, KE);
    const double vomc = __dmul_rn(v, b.omc);
    const double term1 = __dmul_rn(vomc, b.gR2);
    const double term2 = __dmul_rn(b.two_R, __dsqrt_rn(__dmul_rn(vomc, b.c)));
    const double alpha2 = __dadd_rn(__dadd_rn(b.c, term1), term2);
    const double mag = __dsqrt_rn(alpha2);
    const double sign_term = __dadd_rn(b.r_normal, __dsqrt_rn(__ddiv_rn(__dmul_rn(b.cdof, KE), b.den)));
    return sign_term >= 0.0 ? mag : -mag;
    }

template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void finalize(const ForceIn& f, const BussiIn& b, BlockScratch& sc, Scalars* scalars,
                                         bool publish);

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
    {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
        v = o < v ? o : v;
        }
    return v;
    }

// Merge G records into sc.rec and (FINALIZE) form sc.fin / Scalars, all valid for every thread after
// the call.  Latency is what matters here -- the whole grid is waiting -- so:
//   * only the warps that hold records work (thread j takes records j, j + blockDim, ...);
//   * the vote for the first 'L' particle, the pair trees and the finalize need TWO block barriers;
//   * the photon position (sector 3) is fetched by the one lane whose record holds the candidate,
//     concurrently with the tree; sector 4 is touched only when several 'L' particles exist.
//   * WIDE_POLL (folder CTA of k_split_folder): the photon sector is fetched eagerly with the others and a
//     thread whose record has not arrived re-reads ALL its sectors on every attempt, so the fold is one L2
//     round trip behind the last record.  When the WHOLE grid polls (k_fused, k_split) that is 4x the
//     traffic on the same few lines and delays the publishers (force call 23.9 -> 25.4 us, measured), so
//     there a thread polls one sector at a time.
template<bool FORCE, bool BUSSI, bool WAIT, bool FINALIZE, bool SYS = false, bool WIDE_POLL = false>
__device__ __forceinline__ void combine_phase(const Partial* __restrict__ recs, int G, unsigned long long epoch,
                                              const ForceIn& f, const BussiIn& b, BlockScratch& sc, Scalars* scalars,
                                              bool publish, unsigned long long* dbg = nullptr)
    {
    const unsigned int tid = threadIdx.x;
    const unsigned int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    // (Measured, profiles/fold_threads_r2a.txt: letting a thread take two or more records so that fewer warps run the
    // compensated tree -- 5, 3 or 2 warps instead of 10 -- makes the force call SLOWER, 22.2 -> 23.1 / 23.7 / 24.2 us:
    // the fold is bound by the chain of round trips per thread, not by FP64 throughput.)
    const unsigned int work_warps = min(nwarps, (unsigned int)((G + 31) >> 5));
    bool late = false;

    if (warp < work_warps)
        {
        Acc a;
        Slot v;
        acc_zero(a);
        unsigned long long mykey = ~0ull;
        unsigned int mycnt = 0, mymulti = 0;
        int myrec = -1;
        double4 mys3 = make_double4(0.0, 0.0, 0.0, 0.0);
        for (int j = tid; j < G; j += blockDim.x)
            {
            const double4* p = reinterpret_cast<const double4*>(recs + j);
            // issue all sector reads before looking at any of them
            double4 s2 = ld_rec<SYS>(p + 2), s0, s1, s3;
            if (FORCE)
                {
                s0 = ld_rec<SYS>(p + 0);
                s1 = ld_rec<SYS>(p + 1);
                // folder: the photon's position travels with the other sectors, so the one lane that turns
                // out to hold the candidate has it without a second, dependent L2 round trip
                if (WIDE_POLL)
                    s3 = ld_rec<SYS>(p + 3);
                }
            if (!WIDE_POLL)
                {
                // whole-grid fold: poll ONE sector (every CTA polls every record); once it is there, fetch the
                // other sectors that were stale at first look TOGETHER -- and, for the rare record that holds an
                // 'L' candidate, its position sector with them -- instead of one dependent round trip each
                bool refetch = false;
                if (WAIT && !sector_ok(s2, epoch))
                    {
                    s2 = read_sector<true, SYS>(p + 2, epoch, late);
                    refetch = true;
                    }
                if (FORCE)
                    {
                    const bool cand = dbits(s2.y) != ~0ull;
                    if (WAIT && (refetch || !sector_ok(s0, epoch) || !sector_ok(s1, epoch)))
                        {
                        s0 = ld_rec<SYS>(p + 0);
                        s1 = ld_rec<SYS>(p + 1);
                        }
                    if (cand)
                        s3 = ld_rec<SYS>(p + 3);
                    if (WAIT)
                        {
                        if (!sector_ok(s0, epoch))
                            s0 = read_sector<true, SYS>(p + 0, epoch, late);
                        if (!sector_ok(s1, epoch))
                            s1 = read_sector<true, SYS>(p + 1, epoch, late);
                        if (cand && !sector_ok(s3, epoch))
                            s3 = read_sector<true, SYS>(p + 3, epoch, late);
                        }
                    }
                }
            if (WAIT && WIDE_POLL)
                {
                // poll ALL the sectors this thread needs with one round trip per attempt (checking them one
                // after the other costs a dependent L2 round trip each -- 1-2 us apiece while the grid streams)
                bool ok = sector_ok(s2, epoch);
                if (FORCE)
                    ok = ok && sector_ok(s0, epoch) && sector_ok(s1, epoch) && sector_ok(s3, epoch);
                if (!ok)
                    {
                    const unsigned long long t0 = globaltimer_ns();
                    do
                        {
                        __nanosleep(32);
                        s2 = ld_rec<SYS>(p + 2);
                        if (FORCE)
                            {
                            s0 = ld_rec<SYS>(p + 0);
                            s1 = ld_rec<SYS>(p + 1);
                            s3 = ld_rec<SYS>(p + 3);
                            }
                        ok = sector_ok(s2, epoch);
                        if (FORCE)
                            ok = ok && sector_ok(s0, epoch) && sector_ok(s1, epoch) && sector_ok(s3, epoch);
                        if (!ok && globaltimer_ns() - t0 > (SYS ? PEER_TIMEOUT_NS : HANDOFF_TIMEOUT_NS))
                            {
                            late = true; // never hang the GPU
                            break;
                            }
                        } while (!ok);
                    }
                }
            if (FORCE)
                {
                pair_add(a.dhi[0], a.dlo[0], s0.x, s1.x);
                pair_add(a.dhi[1], a.dlo[1], s0.y, s1.y);
                pair_add(a.dhi[2], a.dlo[2], s0.z, s1.z);
                const unsigned long long key = dbits(s2.y);
                if (key != ~0ull)
                    {
                    mycnt++;
                    mymulti |= (key & MULTI_L_BIT) ? 1u : 0u;
                    const unsigned long long idx = key & ~MULTI_L_BIT;
                    if (idx < mykey)
                        {
                        mykey = idx;
                        myrec = j;
                        mys3 = s3;
                        }
                    }
                }
            if (BUSSI)
                a.ke += s2.x;
            }
        if (dbg && tid == 0)
            dbg[2] = globaltimer_ns();
        if (FORCE)
            {
            // warp-level vote; the winning lane starts fetching q right away
            const unsigned long long wmin = warp_min_u64(mykey);
            const unsigned int wcnt = __reduce_add_sync(0xffffffffu, mycnt);
            const unsigned int wmulti = __any_sync(0xffffffffu, mymulti != 0);
            double4 s3 = make_double4(0.0, 0.0, 0.0, 0.0), s4 = s3;
            const bool i_hold = (wmin != ~0ull) && (mykey == wmin);
            if (i_hold)
                {
                s3 = mys3; // fetched with the record's other sectors
                if (!FINALIZE) // a rank-level merge may lose the global vote later: keep the term too
                    s4 = read_sector<WAIT, SYS>(reinterpret_cast<const double4*>(recs + myrec) + 4, epoch, late);
                }
            v = warp_tree<FORCE, BUSSI>(a, lane);
            if (i_hold)
                {
                sc.wq[warp][0] = s3.x;
                sc.wq[warp][1] = s3.y;
                sc.wq[warp][2] = s3.z;
                sc.wt[warp][0] = s4.x;
                sc.wt[warp][1] = s4.y;
                sc.wt[warp][2] = s4.z;
                }
            if (lane == 0)
                {
                sc.red_ull[warp] = wmin;
                sc.red_cnt[warp] = wcnt | (wmulti ? 0x80000000u : 0u);
                }
            }
        else
            v = warp_tree<FORCE, BUSSI>(a, lane);
        cross_warp_put<FORCE, BUSSI>(v, 0u, sc, lane, warp);
        if (late)
            atomicOr(&sc.flags, 1u);
        }
    __syncthreads();
    if (dbg && tid == 0)
        dbg[5] = globaltimer_ns();
    if (warp == 0)
        {
        unsigned int unused_n_L;
        Slot t = cross_warp_get<FORCE, BUSSI>(unused_n_L, sc, lane, work_warps);
        unsigned long long gmin = ~0ull;
        int many = 0;
        if (FORCE)
            {
            const unsigned long long wkey = lane < work_warps ? sc.red_ull[lane] : ~0ull;
            const unsigned int wc = lane < work_warps ? sc.red_cnt[lane] : 0u;
            gmin = warp_min_u64(wkey);
            const unsigned int cnt = __reduce_add_sync(0xffffffffu, wc & 0x7fffffffu);
            many = (cnt > 1) || __any_sync(0xffffffffu, (wc & 0x80000000u) != 0);
            // q of the global winner: the lowest warp whose vote equals gmin
            const unsigned int holders = __ballot_sync(0xffffffffu, gmin != ~0ull && wkey == gmin);
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0;
            if (holders)
                {
                const int w = __ffs(holders) - 1;
                q0 = sc.wq[w][0];
                q1 = sc.wq[w][1];
                q2 = sc.wq[w][2];
                t0 = sc.wt[w][0];
                t1 = sc.wt[w][1];
                t2 = sc.wt[w][2];
                }
            if (many)
                {
                // several 'L' particles: every candidate that is not the global first is an ordinary
                // member of the dipole sum (CavityForceCompute.cc:120-126) -- add its term back
                Acc extra;
                acc_zero(extra);
                bool late2 = false;
                for (int j = lane; j < G; j += 32)
                    {
                    const double4* p = reinterpret_cast<const double4*>(recs + j);
                    const double4 s2 = read_sector<WAIT, SYS>(p + 2, epoch, late2);
                    const unsigned long long key = dbits(s2.y);
                    if (key != ~0ull && (key & ~MULTI_L_BIT) != gmin)
                        {
                        const double4 s4 = read_sector<WAIT, SYS>(p + 4, epoch, late2);
                        two_sum_acc(extra.dhi[0], extra.dlo[0], s4.x);
                        two_sum_acc(extra.dhi[1], extra.dlo[1], s4.y);
                        two_sum_acc(extra.dhi[2], extra.dlo[2], s4.z);
                        }
                    }
                const Slot e = warp_tree<true, false>(extra, lane); // pair 3 of it is zero
                pair_add(t.hi, t.lo, e.hi, e.lo);
                if (late2)
                    atomicOr(&sc.flags, 1u);
                }
            if (lane == 0)
                {
                sc.rec.q[0] = q0;
                sc.rec.q[1] = q1;
                sc.rec.q[2] = q2;
                sc.rec.t[0] = t0; // (zero after a final merge: the global first's own term never enters d)
                sc.rec.t[1] = t1;
                sc.rec.t[2] = t2;
                }
            }
        slot_to_record<FORCE, BUSSI>(t, sc.rec, lane);
        __syncwarp();
        if (lane == 0)
            {
            sc.rec.n_L = many ? 2ull : (gmin != ~0ull ? 1ull : 0ull);
            sc.rec.first_L = gmin == ~0ull ? gmin : (gmin | (many ? MULTI_L_BIT : 0ull));
            if (dbg)
                dbg[6] = globaltimer_ns();
            if (FINALIZE)
                finalize<FORCE, BUSSI>(f, b, sc, scalars, publish);
            if (dbg)
                dbg[7] = globaltimer_ns();
            }
        }
    __syncthreads();
    }

// thread 0 only: merged record -> Final (shared) and, when `publish`, the device-resident Scalars
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void finalize(const ForceIn& f, const BussiIn& b, BlockScratch& sc, Scalars* scalars,
                                         bool publish)
    {
    const Partial& m = sc.rec;
    Final fin;
    fin.many_L = 0;
    fin.has_photon = 0;
    fin.photon_local = -1;
    fin.Dq[0] = fin.Dq[1] = 0.0;
    fin.FL[0] = fin.FL[1] = fin.FL[2] = 0.0;
    fin.alpha = 1.0;
    fin.bussi_ok = 1;
    fin.timeout = (int)(sc.flags & 1u);
    double en[3] = {0, 0, 0}, d[3] = {0, 0, 0}, q[3] = {0, 0, 0};
    unsigned long long first = ~0ull;
    if (FORCE && m.first_L != ~0ull)
        {
        fin.has_photon = 1;
        fin.many_L = (m.first_L & MULTI_L_BIT) != 0;
        first = m.first_L & ~MULTI_L_BIT;
        const unsigned long long lo = f.index_offset;
        if (first >= lo && first < lo + f.N)
            fin.photon_local = (long long)(first - lo);
#pragma unroll
        for (int k = 0; k < 3; k++)
            {
            d[k] = __dadd_rn(m.dhi[k], m.dlo[k]);
            q[k] = m.q[k];
            }
        // CavityForceCompute.cc:174-176; half_K = 0.5*K and half_g2K = 0.5*(g*g/K) from the host
        const double qq = __dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2]));
        const double dq = __dadd_rn(__dadd_rn(__dmul_rn(d[0], q[0]), __dmul_rn(d[1], q[1])), 0.0);
        const double dd = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), 0.0);
        en[0] = __dmul_rn(f.half_K, qq);
        en[1] = __dmul_rn(f.g, dq);
        en[2] = __dmul_rn(f.half_g2K, dd);
        // :183, gk = g/K from the host
        fin.Dq[0] = __dadd_rn(q[0], __dmul_rn(f.gk, d[0]));
        fin.Dq[1] = __dadd_rn(q[1], __dmul_rn(f.gk, d[1]));
        // :203
        fin.FL[0] = __dadd_rn(__dmul_rn(-f.K, q[0]), -__dmul_rn(f.g, d[0]));
        fin.FL[1] = __dadd_rn(__dmul_rn(-f.K, q[1]), -__dmul_rn(f.g, d[1]));
        fin.FL[2] = __dadd_rn(__dmul_rn(-f.K, q[2]), -__dmul_rn(f.g, 0.0));
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
            scalars->photon_idx = fin.has_photon ? (long long)first : -1;
            scalars->n_L = (unsigned int)m.n_L;
            scalars->err_force = fin.timeout ? 2.0 : 0.0; // written by every call, so one failure does not stick
            }
        if (BUSSI)
            {
            scalars->ke = KE;
            if (b.rescale)
                {
                scalars->alpha = fin.alpha;
                scalars->inst = inst;
                scalars->cumulative = __dadd_rn(scalars->cumulative, inst); // :90
                }
            scalars->err = fin.timeout ? 2.0 : (fin.bussi_ok ? 0.0 : 1.0);
            }
        if (fin.timeout)
            raise_fault(scalars);
        }
    }

// ---- hand-off inside ONE thread-block cluster (k_cluster): the records travel through distributed shared memory ---------
// Every CTA stores its record into the inbox of every CTA of the cluster (st.shared::cluster, 20 words per peer), the cluster
// barrier orders the stores, and warp 0 of every CTA folds its own inbox: the fold of combine_phase for G <= 32 records -- one
// record per lane, the same transposed tree, the same vote for the first 'L' particle, the same put-back of the other
// candidates' terms -- without a global-memory round trip.
__device__ __forceinline__ void cluster_post(Partial* inbox, const Partial& rec, unsigned int rank, unsigned int nctas)
    {
    const double* src = reinterpret_cast<const double*>(&rec);
    double* slot = reinterpret_cast<double*>(inbox + rank);
    constexpr unsigned int W = sizeof(Partial) / 8;
    for (unsigned int w = threadIdx.x; w < W * nctas; w += blockDim.x)
        st_dsmem_f64(slot + w % W, w / W, src[w % W]);
    }
template<bool FORCE, bool BUSSI>
__device__ __forceinline__ void combine_cluster(const Partial* inbox, unsigned int C, const ForceIn& f, const BussiIn& b,
                                                BlockScratch& sc, Scalars* scalars, bool publish)
    {
    const unsigned int lane = threadIdx.x & 31u;
    if (threadIdx.x < 32u)
        {
        const bool have = lane < C;
        const Partial& r = inbox[have ? lane : 0u];
        Acc a;
        acc_zero(a);
        unsigned long long key = ~0ull;
        if (have)
            {
            if (FORCE)
                {
#pragma unroll
                for (int k = 0; k < 3; k++)
                    pair_add(a.dhi[k], a.dlo[k], r.dhi[k], r.dlo[k]);
                key = r.first_L;
                }
            if (BUSSI)
                a.ke += r.ke;
            }
        Slot t = warp_tree<FORCE, BUSSI>(a, lane);
        unsigned long long gmin = ~0ull;
        int many = 0;
        if (FORCE)
            {
            const bool cand = key != ~0ull;
            const unsigned long long idx = cand ? (key & ~MULTI_L_BIT) : ~0ull;
            gmin = warp_min_u64(idx);
            const unsigned int cnt = __popc(__ballot_sync(0xffffffffu, cand));
            many = (cnt > 1) || __any_sync(0xffffffffu, cand && (key & MULTI_L_BIT) != 0);
            const unsigned int holders = __ballot_sync(0xffffffffu, cand && idx == gmin);
            double q0 = 0.0, q1 = 0.0, q2 = 0.0;
            if (holders)
                {
                const int w = __ffs(holders) - 1;
                q0 = __shfl_sync(0xffffffffu, r.q[0], w);
                q1 = __shfl_sync(0xffffffffu, r.q[1], w);
                q2 = __shfl_sync(0xffffffffu, r.q[2], w);
                }
            if (many)
                {
                // every candidate that is not the global first is an ordinary member of the dipole sum (as combine_phase)
                Acc extra;
                acc_zero(extra);
                if (cand && idx != gmin)
                    {
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        two_sum_acc(extra.dhi[k], extra.dlo[k], r.t[k]);
                    }
                const Slot e = warp_tree<true, false>(extra, lane);
                pair_add(t.hi, t.lo, e.hi, e.lo);
                }
            if (lane == 0)
                {
                sc.rec.q[0] = q0;
                sc.rec.q[1] = q1;
                sc.rec.q[2] = q2;
                sc.rec.t[0] = sc.rec.t[1] = sc.rec.t[2] = 0.0;
                }
            }
        slot_to_record<FORCE, BUSSI>(t, sc.rec, lane);
        __syncwarp();
        if (lane == 0)
            {
            sc.rec.n_L = many ? 2ull : (gmin != ~0ull ? 1ull : 0ull);
            sc.rec.first_L = gmin == ~0ull ? gmin : (gmin | (many ? MULTI_L_BIT : 0ull));
            finalize<FORCE, BUSSI>(f, b, sc, scalars, publish);
            }
        }
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
    if (!fin.many_L)
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

template<int UNROLL>
__device__ __forceinline__ void apply_force(const Final& fin, const ForceIn& f, const StreamGrid g = whole_grid())
    {
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    unsigned long long i = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
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

template<int UNROLL, bool LIST>
__device__ __forceinline__ void apply_rescale(double alpha, const BussiIn& b, const StreamGrid g = whole_grid())
    {
    constexpr int U = UNROLL > 1 ? UNROLL : 1;
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    unsigned long long j = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
    const unsigned long long n = b.n;
    if (LIST)
        {
        // index list: indices fetched one iteration ahead (see reduce_ke)
        unsigned int nxt[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            nxt[k] = (j + k * stride < n) ? __ldg(b.gidx + j + k * stride) : 0u;
        for (; j < n; j += U * stride)
            {
            double4 v[U];
            unsigned int cur[U];
#pragma unroll
            for (int k = 0; k < U; k++)
                {
                cur[k] = nxt[k];
                if (j + k * stride < n)
                    v[k] = ld256(b.vel + cur[k]);
                }
#pragma unroll
            for (int k = 0; k < U; k++)
                {
                const unsigned long long jn = j + (U + k) * stride;
                nxt[k] = jn < n ? __ldg(b.gidx + jn) : 0u;
                }
#pragma unroll
            for (int k = 0; k < U; k++)
                if (j + k * stride < n)
                    {
                    v[k].x = __dmul_rn(v[k].x, alpha);
                    v[k].y = __dmul_rn(v[k].y, alpha);
                    v[k].z = __dmul_rn(v[k].z, alpha);
                    st256_vel(b.vel + cur[k], v[k], b.stream_st != 0);
                    }
            }
        }
    else
        {
        for (; j + (U - 1) * stride < n; j += U * stride)
            {
            double4 v[U];
#pragma unroll
            for (int k = 0; k < U; k++)
                v[k] = ld256(b.vel + b.first + j + k * stride);
#pragma unroll
            for (int k = 0; k < U; k++)
                {
                v[k].x = __dmul_rn(v[k].x, alpha);
                v[k].y = __dmul_rn(v[k].y, alpha);
                v[k].z = __dmul_rn(v[k].z, alpha);
                st256_vel(b.vel + b.first + j + k * stride, v[k], b.stream_st != 0);
                }
            }
        for (; j < n; j += stride)
            {
            double4 v = ld256(b.vel + b.first + j);
            v.x = __dmul_rn(v.x, alpha);
            v.y = __dmul_rn(v.y, alpha);
            v.z = __dmul_rn(v.z, alpha);
            st256_vel(b.vel + b.first + j, v, b.stream_st != 0);
            }
        }
    }

// fused apply for a group [0, n): charge and velocity loads of U particles first, then the stores
template<bool FORCE, bool RESCALE, int U>
__device__ __forceinline__ void apply_stream(const Final& fin, const ForceIn& f, const BussiIn& b,
                                             const StreamGrid g = whole_grid())
    {
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    const unsigned long long nf = FORCE ? (unsigned long long)f.N : 0ull;
    const unsigned long long nk = RESCALE ? (unsigned long long)b.n : 0ull;
    const unsigned long long common = (FORCE && RESCALE) ? (nf < nk ? nf : nk) : (FORCE ? nf : nk);
    const unsigned long long all = nf > nk ? nf : nk;
    const double alpha = fin.alpha;
    unsigned long long i = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < common; i += U * stride)
        {
        double c[FORCE ? U : 1];
        double4 v[RESCALE ? U : 1];
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            if (FORCE)
                c[k] = ld_charge_last(f.charge + i + k * stride);
            if (RESCALE)
                v[k] = ld256(b.vel + i + k * stride);
            }
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            const unsigned long long j = i + k * stride;
            if (FORCE)
                st256_stream(f.force + j, force_of(j, c[k], fin, f));
            if (RESCALE)
                {
                v[k].x = __dmul_rn(v[k].x, alpha);
                v[k].y = __dmul_rn(v[k].y, alpha);
                v[k].z = __dmul_rn(v[k].z, alpha);
                st256_vel(b.vel + j, v[k], b.stream_st != 0);
                }
            }
        }
    for (; i < all; i += stride)
        {
        if (FORCE && i < nf)
            st256_stream(f.force + i, force_of(i, ld_charge_last(f.charge + i), fin, f));
        if (RESCALE && i < nk)
            {
            double4 v = ld256(b.vel + i);
            v.x = __dmul_rn(v.x, alpha);
            v.y = __dmul_rn(v.y, alpha);
            v.z = __dmul_rn(v.z, alpha);
            st256_vel(b.vel + i, v, b.stream_st != 0);
            }
        }
    }

template<bool FORCE, bool BUSSI, int UNROLL, bool LIST = false>
__device__ __forceinline__ void apply_phase(const Final& fin, const ForceIn& f, const BussiIn& b,
                                            const StreamGrid g = whole_grid())
    {
    const bool rescale = BUSSI && b.rescale && fin.bussi_ok && fin.alpha != 1.0;
    if (BUSSI && rescale && !LIST && b.first == 0)
        apply_stream<FORCE, true, UNROLL>(fin, f, b, g);
    else
        {
        if (FORCE)
            apply_stream<true, false, UNROLL>(fin, f, b, g);
        if (BUSSI && rescale)
            apply_rescale<UNROLL, LIST>(fin.alpha, b, g);
        }
    }

// ---- rescale with the first velocities already in registers (k_split_folder) ----------------------
// The streaming CTAs wait ~2.5 us for alpha with nothing to do; the rescale pass that follows re-reads the
// velocities (L2 hits) before it can store anything.  So the first PRE strided velocities of every thread
// are fetched BEFORE the wait and only multiplied and stored after it.  Contiguous group [0, n) only.
template<int PRE> struct VelPrefetch
    {
    double4 v[PRE];
    };
template<int PRE> __device__ __forceinline__ VelPrefetch<PRE> prefetch_vel(const BussiIn& b, const StreamGrid g)
    {
    VelPrefetch<PRE> p;
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    const unsigned long long i0 = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < PRE; k++)
        {
        p.v[k] = make_double4(0.0, 0.0, 0.0, 0.0);
        if (i0 + k * stride < (unsigned long long)b.n)
            p.v[k] = ld256(b.vel + i0 + k * stride);
        }
    return p;
    }
template<int PRE, int U>
__device__ __forceinline__ void rescale_prefetched(double alpha, const BussiIn& b, const StreamGrid g, const VelPrefetch<PRE>& p)
    {
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    const unsigned long long n = b.n;
    unsigned long long i = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < PRE; k++)
        {
        if (i + k * stride < n)
            {
            double4 v = p.v[k];
            v.x = __dmul_rn(v.x, alpha);
            v.y = __dmul_rn(v.y, alpha);
            v.z = __dmul_rn(v.z, alpha);
            st256_vel(b.vel + i + k * stride, v, b.stream_st != 0);
            }
        }
    i += (unsigned long long)PRE * stride;
    for (; i + (U - 1) * stride < n; i += U * stride)
        {
        double4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            v[k] = ld256(b.vel + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; k++)
            {
            v[k].x = __dmul_rn(v[k].x, alpha);
            v[k].y = __dmul_rn(v[k].y, alpha);
            v[k].z = __dmul_rn(v[k].z, alpha);
            st256_vel(b.vel + i + k * stride, v[k], b.stream_st != 0);
            }
        }
    for (; i < n; i += stride)
        {
        double4 v = ld256(b.vel + i);
        v.x = __dmul_rn(v.x, alpha);
        v.y = __dmul_rn(v.y, alpha);
        v.z = __dmul_rn(v.z, alpha);
        st256_vel(b.vel + i, v, b.stream_st != 0);
        }
    }

// The same for the force pass of the force-only call (k_fused<1,0>): the first PRE charges of every thread are
// fetched before the hand-off.  (In k_split_folder, where they would be fetched under the KE merge tree, it measured
// slightly worse, 31.37 vs 31.25 us, and is not used there.)
template<int PRE> struct ChargePrefetch
    {
    double c[PRE];
    };
template<int PRE> __device__ __forceinline__ ChargePrefetch<PRE> prefetch_charge(const ForceIn& f, const StreamGrid g)
    {
    ChargePrefetch<PRE> p;
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    const unsigned long long i0 = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < PRE; k++)
        p.c[k] = (i0 + k * stride < (unsigned long long)f.N) ? ld_charge_last(f.charge + i0 + k * stride) : 0.0;
    return p;
    }
template<int PRE, int U>
__device__ __forceinline__ void forces_prefetched(const Final& fin, const ForceIn& f, const StreamGrid g,
                                                  const ChargePrefetch<PRE>& p)
    {
    const unsigned long long stride = (unsigned long long)g.nblk * blockDim.x;
    const unsigned long long N = f.N;
    unsigned long long i = (unsigned long long)g.blk * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < PRE; k++)
        if (i + k * stride < N)
            st256_stream(f.force + i + k * stride, force_of(i + k * stride, p.c[k], fin, f));
    i += (unsigned long long)PRE * stride;
    for (; i + (U - 1) * stride < N; i += U * stride)
        {
        double c[U];
#pragma unroll
        for (int k = 0; k < U; k++)
            c[k] = ld_charge_last(f.charge + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; k++)
            st256_stream(f.force + i + k * stride, force_of(i + k * stride, c[k], fin, f));
        }
    for (; i < N; i += stride)
        st256_stream(f.force + i, force_of(i, ld_charge_last(f.charge + i), fin, f));
    }


// ---- Final record hand-off (k_split_folder) -----------------------------------------------------
// The folder CTA publishes what finalize() formed in the sector format of the reduce records
// (payload x3 + check word, one STG.256 each); every thread of a streaming CTA polls the sectors it
// needs (the loads of a warp coalesce into one request) -- no block barrier, no shared memory.
//   force half:      sector 0 {Dq.x, Dq.y, FL.x}   sector 1 {FL.y, FL.z, -}   sector 2 {photon_local, flags, -}
//   thermostat half: sector 2 {alpha, flags, -}
// flags: bit 0 many_L, bit 1 has_photon, bit 2 bussi_ok, bit 3 timeout
__device__ __forceinline__ unsigned long long final_flags(const Final& fin)
    {
    return (fin.many_L ? 1ull : 0ull) | (fin.has_photon ? 2ull : 0ull) | (fin.bussi_ok ? 4ull : 0ull)
           | (fin.timeout ? 8ull : 0ull);
    }
template<bool FORCE> __device__ __forceinline__ void publish_final(Partial* dst, const Final& fin, unsigned long long epoch)
    {
    double4* d = reinterpret_cast<double4*>(dst);
    if (FORCE)
        {
        st256(d + 0, make_sector(fin.Dq[0], fin.Dq[1], fin.FL[0], epoch));
        st256(d + 1, make_sector(fin.FL[1], fin.FL[2], 0.0, epoch));
        st256(d + 2, make_sector(bitsd((unsigned long long)fin.photon_local), bitsd(final_flags(fin)), 0.0, epoch));
        }
    else
        st256(d + 2, make_sector(fin.alpha, bitsd(final_flags(fin)), 0.0, epoch));
    }
// The sector loads are issued by prefetch_final (before work that does not depend on them, so the L2
// round trip -- 1-2 us while the grid is streaming -- overlaps it) and checked / re-polled by take_final.
struct FinalSectors
    {
    double4 s0, s1, s2;
    };
template<bool FORCE> __device__ __forceinline__ FinalSectors prefetch_final(const Partial* src)
    {
    const double4* p = reinterpret_cast<const double4*>(src);
    FinalSectors r;
    r.s2 = ld_rec<false>(p + 2);
    if (FORCE)
        {
        r.s0 = ld_rec<false>(p + 0);
        r.s1 = ld_rec<false>(p + 1);
        }
    return r;
    }
template<bool FORCE, int POLL_NS = CAVB_FINAL_POLL_NS>
__device__ __forceinline__ Final take_final(const Partial* src, unsigned long long epoch, const FinalSectors& pre)
    {
    const double4* p = reinterpret_cast<const double4*>(src);
    bool late = false;
    Final fin;
    fin.Dq[0] = fin.Dq[1] = fin.FL[0] = fin.FL[1] = fin.FL[2] = 0.0;
    fin.alpha = 1.0;
    fin.photon_local = -1;
    double4 s0 = pre.s0, s1 = pre.s1, s2 = pre.s2;
    if (!sector_ok(s2, epoch))
        s2 = read_sector<true, false, POLL_NS>(p + 2, epoch, late);
    if (FORCE)
        {
        if (!sector_ok(s0, epoch))
            s0 = read_sector<true, false, POLL_NS>(p + 0, epoch, late);
        if (!sector_ok(s1, epoch))
            s1 = read_sector<true, false, POLL_NS>(p + 1, epoch, late);
        fin.Dq[0] = s0.x;
        fin.Dq[1] = s0.y;
        fin.FL[0] = s0.z;
        fin.FL[1] = s1.x;
        fin.FL[2] = s1.y;
        fin.photon_local = (long long)dbits(s2.x);
        }
    else
        fin.alpha = s2.x;
    const unsigned long long fl = dbits(s2.y);
    fin.many_L = (int)(fl & 1ull);
    fin.has_photon = (int)((fl >> 1) & 1ull);
    fin.bussi_ok = (int)((fl >> 2) & 1ull);
    fin.timeout = (int)((fl >> 3) & 1ull) | (late ? 1 : 0);
    return fin;
    }
    } // namespace cavb
