// cavb200_internal.cuh -- handle, device-resident scalar blocks and small device helpers shared
// by the translation units of libcavb200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cavb200.h"

#define CAVB_CHECK(expr)                       \
    do                                         \
        {                                      \
        cudaError_t _e = (expr);               \
        if (_e != cudaSuccess)                 \
            return (int)_e;                    \
        } while (0)

namespace cavb
    {
constexpr uint32_t NO_INDEX = 0xFFFFFFFFu;
constexpr int MAX_PARTIALS = 2048; // upper bound on the grid of a reduce pass

// One reduce-pass record per CTA (and, in sharded mode, per rank): 128 bytes.
// d is carried as unevaluated (hi, lo) pairs so that combining is (nearly) exact and therefore
// independent of how many CTAs / ranks took part up to the final rounding.
struct __align__(16) Partial
    {
    double dhi[3];
    double dlo[3];
    double ke;           // sum m |v|^2 (NOT halved)
    double q[3];         // unwrapped position of this CTA's first 'L' particle
    double t[3];         // its dipole term charge*u (added back if it is not the global first)
    double pad_d;
    unsigned long long first_L; // GLOBAL index of that particle, ~0ull if none
    unsigned int n_L;    // number of 'L' particles seen
    unsigned int pad;
    };
static_assert(sizeof(Partial) == CAVB200_SHARD_RECORD_BYTES, "record size is part of the ABI");

// Device-resident results of the last call; host getters copy this back lazily.
struct __align__(16) Scalars
    {
    // cavity force
    double energies[3];  // harmonic, coupling, dipole self
    double dipole[3];
    double Dq[2];
    double q[3];
    double FL[3];
    long long photon_idx; // global index, -1 if none
    unsigned int n_L;
    unsigned int pad0;
    // Bussi
    double ke;           // 1/2 sum m|v|^2 seen by the last call
    double alpha;
    double inst;         // instantaneous reservoir energy  KE (1 - alpha^2)
    double cumulative;   // running sum of inst
    double err;          // 0 ok, 1 = zero kinetic energy with dof != 0, 2 = barrier timeout
    double pad1[3];
    };

struct Tuning
    {
    int variant;         // 0 = reduce kernel + apply kernel, 1 = one cooperative persistent kernel
    int threads;         // CTA size of the streaming kernels (256 / 512 / 1024)
    int ctas_per_sm;     // resident CTAs per SM the grid is sized for
    int unroll;          // independent particle loads in flight per thread (1, 2 or 4)
    int prefetch;        // cooperative variant: issue the apply pass's first loads before the barrier
    int rhok_threads;
    int rhok_kblock;     // k-vectors handled per thread in the F(k,t) kernel
    };

struct ShardState
    {
    int mode;            // 0 nccl, 1 nvlink mailbox
    int rank, nranks;
    void* nccl_comm;
    Partial* gather;     // device, nranks records (nccl all-gather target / local mailbox)
    Partial* peer_mailbox[16]; // IPC-mapped mailbox base of every rank (own entry = local)
    unsigned long long* flags;      // device, per-rank arrival sequence numbers (inside mailbox allocation)
    unsigned long long* peer_flags[16];
    void* mailbox_alloc;
    unsigned long long seq;
    };
    } // namespace cavb

struct cavb200_handle
    {
    int device;
    int num_sms;
    int coop_supported;
    cavb::Partial* partials;          // MAX_PARTIALS records
    cavb::Scalars* scalars;           // 1 record
    unsigned long long* counters;     // [0] barrier arrivals, [1] departures, [4] reduce-pass ticket, [8..] Final
    uint64_t launches;
    cavb::Tuning tune;
    cavb::ShardState shard;
    // host-buffer pipeline staging (grown on demand, owned)
    void* stage[5];
    uint64_t stage_bytes[5];
    cudaStream_t copy_streams[3];
    cudaEvent_t copy_events[64];
    int copy_ready;
    // F(k,t) workspace
    double* rhok_partials;
    uint64_t rhok_partials_bytes;
    };

namespace cavb
    {
// ---- 256-bit global accesses (PTX ISA 8.8, sm_100+): one LDG.E.256 / STG.E.256 per Scalar4 ----
__device__ __forceinline__ double4 ld256_stream(const double4* p)
    {
    double4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
    }
__device__ __forceinline__ double4 ld256(const double4* p)
    {
    double4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
    }
__device__ __forceinline__ void st256(double4* p, const double4& v)
    {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w)
                 : "memory");
    }
__device__ __forceinline__ void st256_stream(double4* p, const double4& v)
    {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y),
                 "d"(v.z), "d"(v.w)
                 : "memory");
    }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
    {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
    }
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v)
    {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
__device__ __forceinline__ unsigned long long atom_acq_rel_add_u64(unsigned long long* p, unsigned long long v)
    {
    unsigned long long old;
    asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
    }
__device__ __forceinline__ unsigned long long globaltimer_ns()
    {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
    }

// ---- compensated accumulation: (hi, lo) is an unevaluated sum; error-free transforms only ----
__device__ __forceinline__ void two_sum_acc(double& hi, double& lo, double t)
    {
    const double s = __dadd_rn(hi, t);
    const double bb = __dadd_rn(s, -hi);
    const double e = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(t, -bb));
    hi = s;
    lo = __dadd_rn(lo, e);
    }
__device__ __forceinline__ void pair_add(double& hi, double& lo, double hi2, double lo2)
    {
    two_sum_acc(hi, lo, hi2);
    lo = __dadd_rn(lo, lo2);
    }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
    } // namespace cavb

// teardown hooks of the sharded and host-buffer paths (shard.cu, host.cu)
void cavb_shard_release(cavb200_handle* h);
void cavb_host_release(cavb200_handle* h);

// launch helpers implemented in the .cu files
namespace cavb
    {
struct ForceIn
    {
    const double4* pos;
    const double* charge;
    const int* image;
    double4* force;
    uint32_t N;
    unsigned long long index_offset; // global index of particle 0 (sharded mode)
    double Lx, Ly, Lz;
    uint32_t L_typeid;
    double g, K;
    };
struct BussiIn
    {
    double4* vel;
    const uint32_t* gidx;
    uint32_t first, n;
    double kT, c /* exp(-dt/tau), host-evaluated */, dof, r_normal, r_gamma /* already doubled */;
    int rescale;         // 0: KE only
    };

int launch_hotpath(cavb200_handle* h, const ForceIn* f, const BussiIn* b, cudaStream_t s);
int launch_shard_step(cavb200_handle* h, const ForceIn* f, const BussiIn* b, cudaStream_t s);
    } // namespace cavb
