// host.cu -- host-buffer entry point (the end-to-end path bench.py reports as "e2e").
//
// A HOOMD GPU run keeps every particle array resident on the device, so the device-pointer entry
// points are what the plugin calls.  This entry point exists for callers whose data lives in host
// memory: it stages through device buffers owned by the handle and orders the copies so that the
// two PCIe directions overlap where the data dependencies allow:
//     stream A:  H2D vel ............ Bussi kernel ... D2H vel
//     stream B:  (after H2D vel) H2D pos, charge, image ... (after Bussi) force kernel ... D2H force
// The force of particle i needs the dipole of ALL particles, so no force byte can leave before the
// last position byte has arrived; the velocity round trip is the only transfer that can hide
// behind the position upload.  Per step PCIe carries 84 B/particle in and 64 B/particle out.
#include "cavb200_internal.cuh"

#include <math.h>
#include <string.h>

using namespace cavb;

namespace
    {
int ensure_stage(cavb200_handle* h, int slot, uint64_t bytes)
    {
    if (h->stage_bytes[slot] >= bytes)
        return 0;
    cudaFree(h->stage[slot]);
    h->stage[slot] = nullptr;
    h->stage_bytes[slot] = 0;
    CAVB_CHECK(cudaMalloc(&h->stage[slot], bytes));
    h->stage_bytes[slot] = bytes;
    return 0;
    }

int ensure_streams(cavb200_handle* h)
    {
    if (h->copy_ready)
        return 0;
    for (int i = 0; i < 3; i++)
        CAVB_CHECK(cudaStreamCreateWithFlags(&h->copy_streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++)
        CAVB_CHECK(cudaEventCreateWithFlags(&h->copy_events[i], cudaEventDisableTiming));
    h->copy_ready = 1;
    return 0;
    }
    } // namespace

void cavb_host_release(cavb200_handle* h)
    {
    for (int i = 0; i < 5; i++)
        cudaFree(h->stage[i]);
    if (h->copy_ready)
        {
        for (int i = 0; i < 3; i++)
            cudaStreamDestroy(h->copy_streams[i]);
        for (int i = 0; i < 4; i++)
            cudaEventDestroy(h->copy_events[i]);
        }
    h->copy_ready = 0;
    }

extern "C" int cavb200_step_host(cavb200_handle* h, const double* pos, const double* charge, const int32_t* image,
                                 double* force, double* vel, uint32_t N, double Lx, double Ly, double Lz,
                                 uint32_t L_typeid, const cavb200_params* params, uint32_t group_first,
                                 uint32_t n_group, const cavb200_bussi_args* bussi, double energies[3],
                                 double bussi_out[5])
    {
    if (!h || !params || !bussi)
        return (int)cudaErrorInvalidValue;
    if (N == 0)
        return 0;
    if (!pos || !charge || !image || !force || !vel)
        return (int)cudaErrorInvalidValue;
    if ((unsigned long long)group_first + n_group > N)
        return (int)cudaErrorInvalidValue;
    CAVB_CHECK(cudaSetDevice(h->device));
    int rc;
    const uint64_t n = N;
    if ((rc = ensure_streams(h)) || (rc = ensure_stage(h, 0, 32 * n)) || (rc = ensure_stage(h, 1, 8 * n))
        || (rc = ensure_stage(h, 2, 12 * n)) || (rc = ensure_stage(h, 3, 32 * n)) || (rc = ensure_stage(h, 4, 32 * n)))
        return rc;
    double* d_pos = (double*)h->stage[0];
    double* d_charge = (double*)h->stage[1];
    int32_t* d_image = (int32_t*)h->stage[2];
    double* d_force = (double*)h->stage[3];
    double* d_vel = (double*)h->stage[4];
    cudaStream_t sA = h->copy_streams[0], sB = h->copy_streams[1];
    cudaEvent_t vel_in = h->copy_events[0], bussi_done = h->copy_events[1], b_done = h->copy_events[2];

    const bool thermostat = bussi->deltaT != 0.0 && n_group > 0;
    // stream A: velocities in, thermostat, velocities out (only the group's range moves)
    const uint64_t voff = 4ull * group_first;
    if (thermostat)
        {
        CAVB_CHECK(cudaMemcpyAsync(d_vel + voff, vel + voff, 32ull * n_group, cudaMemcpyHostToDevice, sA));
        }
    CAVB_CHECK(cudaEventRecord(vel_in, sA));
    if (thermostat)
        {
        rc = cavb200_bussi(h, d_vel, nullptr, group_first, n_group, bussi, sA);
        if (rc)
            return rc;
        }
    CAVB_CHECK(cudaEventRecord(bussi_done, sA));
    if (thermostat)
        CAVB_CHECK(cudaMemcpyAsync(vel + voff, d_vel + voff, 32ull * n_group, cudaMemcpyDeviceToHost, sA));

    // stream B: positions in behind the velocities, force kernel behind the thermostat kernel
    CAVB_CHECK(cudaStreamWaitEvent(sB, vel_in, 0));
    CAVB_CHECK(cudaMemcpyAsync(d_pos, pos, 32 * n, cudaMemcpyHostToDevice, sB));
    CAVB_CHECK(cudaMemcpyAsync(d_charge, charge, 8 * n, cudaMemcpyHostToDevice, sB));
    CAVB_CHECK(cudaMemcpyAsync(d_image, image, 12 * n, cudaMemcpyHostToDevice, sB));
    CAVB_CHECK(cudaStreamWaitEvent(sB, bussi_done, 0));
    rc = cavb200_force(h, d_pos, d_charge, d_image, d_force, N, Lx, Ly, Lz, L_typeid, params, sB);
    if (rc)
        return rc;
    CAVB_CHECK(cudaMemcpyAsync(force, d_force, 32 * n, cudaMemcpyDeviceToHost, sB));
    CAVB_CHECK(cudaEventRecord(b_done, sB));
    CAVB_CHECK(cudaStreamWaitEvent(sA, b_done, 0));

    Scalars s;
    CAVB_CHECK(cudaMemcpyAsync(&s, h->scalars, sizeof(Scalars), cudaMemcpyDeviceToHost, sA));
    CAVB_CHECK(cudaStreamSynchronize(sA));
    if (energies)
        for (int k = 0; k < 3; k++)
            energies[k] = s.energies[k];
    if (bussi_out)
        {
        bussi_out[0] = s.ke;
        bussi_out[1] = s.alpha;
        bussi_out[2] = s.inst;
        bussi_out[3] = s.cumulative;
        bussi_out[4] = s.err;
        }
    if (s.err == 2.0)
        return (int)cudaErrorLaunchTimeout;
    return 0;
    }
