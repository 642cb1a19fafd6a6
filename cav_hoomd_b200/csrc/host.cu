// host.cu -- host-buffer entry points (the end-to-end path bench.py reports as "e2e").
//
// A HOOMD GPU run keeps every particle array resident on the device, so the device-pointer entry
// points are what the plugin calls.  These entry points exist for callers whose data lives in host
// memory (a driver that holds several replicas on the host): cavb200_step_host_submit stages one
// system through device buffers owned by the handle, cavb200_step_host_wait returns its results.
// Two staging slots, three streams, so that both PCIe directions stay busy across steps:
//     copy-in  stream:  H2D vel(k) | H2D pos, charge, image(k) | H2D vel(k+1) | ...
//     kernel   stream:  Bussi(k) after vel(k) | force(k) after pos(k) | Bussi(k+1) | ...
//     copy-out stream:  D2H vel(k) after Bussi(k) | D2H force(k) after force(k) | scalars(k) | ...
// All kernels are on ONE stream: the persistent kernels size their grids to the whole device and
// hand off between co-resident CTAs, so two of them must never run side by side.
// The force of particle i needs the dipole of ALL particles, so within one step no force byte can
// leave before the last position byte has arrived; the overlap comes from the velocity round trip and,
// with two slots, from the next system's upload running under this system's download.
// Per step PCIe carries 84 B/particle in and 64 B/particle out; cavb200_step_host_submit_ex lets the caller leave out
// what has not changed since the slot's last submit (charges never change in an MD run, images rarely) and take the
// rank-1 result {Dq, F_L, energies} instead of the 32 B/particle force array.
#include "cavb200_internal.cuh"

#include <math.h>
#include <string.h>

using namespace cavb;

namespace
    {
enum { EV_VEL_IN = 0, EV_POS_IN, EV_BUSSI_DONE, EV_FORCE_DONE, EV_DONE, EV_PER_SLOT };

int ensure_stage(cavb200_handle* h, int slot, int arr, uint64_t bytes)
    {
    if (h->stage_bytes[slot][arr] >= bytes)
        return 0;
    // the slot may still be in flight from an earlier, smaller use
    if (h->copy_ready)
        CAVB_CHECK(cudaEventSynchronize(h->copy_events[slot * EV_PER_SLOT + EV_DONE]));
    cudaFree(h->stage[slot][arr]);
    h->stage[slot][arr] = nullptr;
    h->stage_bytes[slot][arr] = 0;
    CAVB_CHECK(cudaMalloc(&h->stage[slot][arr], bytes));
    h->stage_bytes[slot][arr] = bytes;
    return 0;
    }

int ensure_streams(cavb200_handle* h)
    {
    if (h->copy_ready)
        return 0;
    for (int i = 0; i < 3; i++)
        CAVB_CHECK(cudaStreamCreateWithFlags(&h->copy_streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < CAVB_HOST_SLOTS * EV_PER_SLOT; i++)
        CAVB_CHECK(cudaEventCreateWithFlags(&h->copy_events[i], cudaEventDisableTiming));
    CAVB_CHECK(cudaMalloc((void**)&h->slot_scalars, CAVB_HOST_SLOTS * sizeof(Scalars)));
    CAVB_CHECK(cudaMallocHost((void**)&h->slot_scalars_host, CAVB_HOST_SLOTS * sizeof(Scalars)));
    h->copy_ready = 1;
    return 0;
    }
    } // namespace

void cavb_host_release(cavb200_handle* h)
    {
    if (h->copy_ready)
        {
        for (int i = 0; i < 3; i++)
            {
            cudaStreamSynchronize(h->copy_streams[i]);
            cudaStreamDestroy(h->copy_streams[i]);
            }
        for (int i = 0; i < CAVB_HOST_SLOTS * EV_PER_SLOT; i++)
            cudaEventDestroy(h->copy_events[i]);
        cudaFree(h->slot_scalars);
        cudaFreeHost(h->slot_scalars_host);
        }
    for (int sl = 0; sl < CAVB_HOST_SLOTS; sl++)
        for (int i = 0; i < 5; i++)
            cudaFree(h->stage[sl][i]);
    h->copy_ready = 0;
    }

extern "C" int cavb200_step_host_submit_ex(cavb200_handle* h, uint32_t slot, const double* pos, const double* charge,
                                           const int32_t* image, double* force, double* vel, uint32_t N, double Lx,
                                           double Ly, double Lz, uint32_t L_typeid, const cavb200_params* params,
                                           uint32_t group_first, uint32_t n_group, const cavb200_bussi_args* bussi,
                                           uint32_t flags)
    {
    if (!h || !params || !bussi || slot >= CAVB_HOST_SLOTS)
        return (int)cudaErrorInvalidValue;
    if (N == 0)
        return 0;
    const bool keep_charge = (flags & CAVB200_HOST_KEEP_CHARGE) != 0, keep_image = (flags & CAVB200_HOST_KEEP_IMAGE) != 0;
    const bool rank1 = (flags & CAVB200_HOST_RANK1_RESULT) != 0;
    if (!pos || (!charge && !keep_charge) || (!image && !keep_image) || (!force && !rank1) || !vel)
        return (int)cudaErrorInvalidValue;
    // "unchanged since the last submit on this slot": the slot's device copy must exist and be of this size
    if ((keep_charge || keep_image) && (!h->slot_used[slot] || h->slot_N[slot] != N))
        return (int)cudaErrorInvalidValue;
    if ((unsigned long long)group_first + n_group > N)
        return (int)cudaErrorInvalidValue;
    CAVB_CHECK(cudaSetDevice(h->device));
    int rc;
    const uint64_t n = N;
    if ((rc = ensure_streams(h)) || (rc = ensure_stage(h, slot, 0, 32 * n)) || (rc = ensure_stage(h, slot, 1, 8 * n))
        || (rc = ensure_stage(h, slot, 2, 12 * n)) || (rc = ensure_stage(h, slot, 3, 32 * n))
        || (rc = ensure_stage(h, slot, 4, 32 * n)))
        return rc;
    double* d_pos = (double*)h->stage[slot][0];
    double* d_charge = (double*)h->stage[slot][1];
    int32_t* d_image = (int32_t*)h->stage[slot][2];
    double* d_force = (double*)h->stage[slot][3];
    double* d_vel = (double*)h->stage[slot][4];
    cudaStream_t sIn = h->copy_streams[0], sK = h->copy_streams[1], sOut = h->copy_streams[2];
    cudaEvent_t* ev = h->copy_events + slot * EV_PER_SLOT;

    const bool thermostat = bussi->deltaT != 0.0 && n_group > 0;
    const uint64_t voff = 4ull * group_first;
    // copy-in: this slot's previous results must have left its buffers (only the group's range of the
    // velocities moves)
    if (h->slot_used[slot])
        CAVB_CHECK(cudaStreamWaitEvent(sIn, ev[EV_DONE], 0));
    if (thermostat)
        CAVB_CHECK(cudaMemcpyAsync(d_vel + voff, vel + voff, 32ull * n_group, cudaMemcpyHostToDevice, sIn));
    CAVB_CHECK(cudaEventRecord(ev[EV_VEL_IN], sIn));
    CAVB_CHECK(cudaMemcpyAsync(d_pos, pos, 32 * n, cudaMemcpyHostToDevice, sIn));
    if (!keep_charge)
        CAVB_CHECK(cudaMemcpyAsync(d_charge, charge, 8 * n, cudaMemcpyHostToDevice, sIn));
    if (!keep_image)
        CAVB_CHECK(cudaMemcpyAsync(d_image, image, 12 * n, cudaMemcpyHostToDevice, sIn));
    CAVB_CHECK(cudaEventRecord(ev[EV_POS_IN], sIn));
    // kernels, one stream
    CAVB_CHECK(cudaStreamWaitEvent(sK, ev[EV_VEL_IN], 0));
    if (thermostat)
        {
        rc = cavb200_bussi(h, d_vel, nullptr, group_first, n_group, bussi, sK);
        if (rc)
            return rc;
        }
    CAVB_CHECK(cudaEventRecord(ev[EV_BUSSI_DONE], sK));
    CAVB_CHECK(cudaStreamWaitEvent(sK, ev[EV_POS_IN], 0));
    // rank-1 result: the dipole reduce alone (nothing written per particle); the caller forms
    // F_i = (-g c_i) Dq from the charges it already holds (reference src/CavityForceCompute.cc:188-200)
    rc = rank1 ? cavb200_force_rank1(h, d_pos, d_charge, d_image, N, Lx, Ly, Lz, L_typeid, params, sK)
               : cavb200_force(h, d_pos, d_charge, d_image, d_force, N, Lx, Ly, Lz, L_typeid, params, sK);
    if (rc)
        return rc;
    // the Scalars block is per handle: keep this step's copy before the next submit's kernels run
    CAVB_CHECK(cudaMemcpyAsync(h->slot_scalars + slot, h->scalars, sizeof(Scalars), cudaMemcpyDeviceToDevice, sK));
    CAVB_CHECK(cudaEventRecord(ev[EV_FORCE_DONE], sK));
    // copy-out
    CAVB_CHECK(cudaStreamWaitEvent(sOut, ev[EV_BUSSI_DONE], 0));
    if (thermostat)
        CAVB_CHECK(cudaMemcpyAsync(vel + voff, d_vel + voff, 32ull * n_group, cudaMemcpyDeviceToHost, sOut));
    CAVB_CHECK(cudaStreamWaitEvent(sOut, ev[EV_FORCE_DONE], 0));
    if (!rank1)
        CAVB_CHECK(cudaMemcpyAsync(force, d_force, 32 * n, cudaMemcpyDeviceToHost, sOut));
    CAVB_CHECK(cudaMemcpyAsync(h->slot_scalars_host + slot, h->slot_scalars + slot, sizeof(Scalars),
                               cudaMemcpyDeviceToHost, sOut));
    CAVB_CHECK(cudaEventRecord(ev[EV_DONE], sOut));
    h->slot_used[slot] = 1;
    h->slot_N[slot] = N;
    return 0;
    }

extern "C" int cavb200_step_host_submit(cavb200_handle* h, uint32_t slot, const double* pos, const double* charge,
                                        const int32_t* image, double* force, double* vel, uint32_t N, double Lx,
                                        double Ly, double Lz, uint32_t L_typeid, const cavb200_params* params,
                                        uint32_t group_first, uint32_t n_group, const cavb200_bussi_args* bussi)
    {
    return cavb200_step_host_submit_ex(h, slot, pos, charge, image, force, vel, N, Lx, Ly, Lz, L_typeid, params, group_first,
                                       n_group, bussi, 0u);
    }

extern "C" int cavb200_step_host_wait_ex(cavb200_handle* h, uint32_t slot, double energies[3], double bussi_out[5],
                                         double rank1_out[6])
    {
    const int rc = cavb200_step_host_wait(h, slot, energies, bussi_out);
    if (rc)
        return rc;
    if (rank1_out)
        {
        const Scalars& s = h->slot_scalars_host[slot];
        rank1_out[0] = s.Dq[0];
        rank1_out[1] = s.Dq[1];
        rank1_out[2] = s.FL[0];
        rank1_out[3] = s.FL[1];
        rank1_out[4] = s.FL[2];
        rank1_out[5] = (double)s.photon_idx;
        }
    return 0;
    }

extern "C" int cavb200_step_host_wait(cavb200_handle* h, uint32_t slot, double energies[3], double bussi_out[5])
    {
    if (!h || slot >= CAVB_HOST_SLOTS)
        return (int)cudaErrorInvalidValue;
    if (!h->copy_ready || !h->slot_used[slot])
        return (int)cudaErrorNotReady; // nothing was submitted on this slot
    CAVB_CHECK(cudaEventSynchronize(h->copy_events[slot * EV_PER_SLOT + EV_DONE]));
    const Scalars& s = h->slot_scalars_host[slot];
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
    if (s.err == 2.0 || s.err_force == 2.0)
        return (int)cudaErrorLaunchTimeout;
    return 0;
    }

extern "C" int cavb200_step_host(cavb200_handle* h, const double* pos, const double* charge, const int32_t* image,
                                 double* force, double* vel, uint32_t N, double Lx, double Ly, double Lz,
                                 uint32_t L_typeid, const cavb200_params* params, uint32_t group_first,
                                 uint32_t n_group, const cavb200_bussi_args* bussi, double energies[3],
                                 double bussi_out[5])
    {
    if (h && N == 0 && params && bussi)
        return 0;
    const int rc = cavb200_step_host_submit(h, 0, pos, charge, image, force, vel, N, Lx, Ly, Lz, L_typeid, params,
                                            group_first, n_group, bussi);
    if (rc)
        return rc;
    return cavb200_step_host_wait(h, 0, energies, bussi_out);
    }
