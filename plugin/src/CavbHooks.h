// CavbHooks.h -- the two small interfaces through which TwoStepConstantVolumeCavity (the fused integration method)
// talks to the thermostat and force classes of this plugin.  They live in different extension modules
// (hoomd.bussi_reservoir._bussi_reservoir / hoomd.cavitymd._cavitymd), so the coupling is a pair of abstract
// classes with virtual methods only: a call lands in the module that constructed the object.
#ifndef CAVB200_HOOKS_H
#define CAVB200_HOOKS_H

#include <cavb200.h>
#include <cstdint>
#include <pybind11/pybind11.h>

namespace hoomd
    {
//! What the fused integration method needs from a Bussi-type thermostat
class PYBIND11_EXPORT CavbBussiSource
    {
    public:
    virtual ~CavbBussiSource() { }
    //! kT(timestep), tau, deltaT, translational dof and the two random draws of this step, made with the generator,
    //! seeding and draw order of reference src/BussiReservoirThermostat.h:63-67,192-200
    virtual cavb200_bussi_args drawBussiArgs(uint64_t timestep, double deltaT) = 0;
    //! from now on alpha / the reservoir energies are formed by kernels running on `h` (the method's handle): the
    //! thermostat's getters read them from there
    virtual void adoptHandle(cavb200_handle* h) = 0;
    };
    } // namespace hoomd
#endif
