// nve.cu -- minimal velocity-Verlet harness around the cavity force (energy-drift comparison).
// Same arithmetic as oracle/cavity_oracle.c orc_nve_step: hm = 0.5*dt/m; v = v + hm*f; r = r + dt*v,
// each product and sum rounded separately (no FMA) so that both arms integrate identically.
#include "cavb200_internal.cuh"

namespace cavb
    {
template<bool DRIFT>
__global__ void __launch_bounds__(256) k_nve(double4* pos, double4* vel, const double4* force, uint32_t N, double dt)
    {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
        {
        double4 v = ld256(vel + i);
        const double4 f = ld256(force + i);
        const double hm = __ddiv_rn(__dmul_rn(0.5, dt), v.w);
        v.x = __dadd_rn(v.x, __dmul_rn(hm, f.x));
        v.y = __dadd_rn(v.y, __dmul_rn(hm, f.y));
        v.z = __dadd_rn(v.z, __dmul_rn(hm, f.z));
        st256(vel + i, v);
        if (DRIFT)
            {
            double4 p = ld256(pos + i);
            p.x = __dadd_rn(p.x, __dmul_rn(dt, v.x));
            p.y = __dadd_rn(p.y, __dmul_rn(dt, v.y));
            p.z = __dadd_rn(p.z, __dmul_rn(dt, v.z));
            st256(pos + i, p);
            }
        }
    }
    } // namespace cavb

using namespace cavb;

static int nve_launch(cavb200_handle* h, double* pos, double* vel, const double* force, uint32_t N, double dt,
                      cudaStream_t s, bool drift)
    {
    if (!h)
        return (int)cudaErrorInvalidValue;
    if (N == 0)
        return 0;
    if (!vel || !force || (drift && !pos))
        return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(vel) | reinterpret_cast<uintptr_t>(force)) & 31)
        return (int)cudaErrorMisalignedAddress;
    unsigned long long want = ((unsigned long long)N + 255) / 256;
    const unsigned long long cap = (unsigned long long)h->num_sms * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (drift)
        k_nve<true><<<grid, 256, 0, s>>>((double4*)pos, (double4*)vel, (const double4*)force, N, dt);
    else
        k_nve<false><<<grid, 256, 0, s>>>(nullptr, (double4*)vel, (const double4*)force, N, dt);
    CAVB_CHECK(cudaGetLastError());
    h->launches += 1;
    return 0;
    }

extern "C" int cavb200_nve_kick_drift(cavb200_handle* h, double* pos, double* vel, const double* force, uint32_t N,
                                      double dt, void* stream)
    {
    return nve_launch(h, pos, vel, force, N, dt, (cudaStream_t)stream, true);
    }

extern "C" int cavb200_nve_half_kick(cavb200_handle* h, double* vel, const double* force, uint32_t N, double dt,
                                     void* stream)
    {
    return nve_launch(h, nullptr, vel, force, N, dt, (cudaStream_t)stream, false);
    }
