// BussiReservoirThermostat.cc -- pybind11 registration of the B200 thermostat host class.
// The Python wrapper (reference src/bussi_reservoir/thermostats.py) looks these names up, so the class name,
// constructor signature, the two read/write properties and the seven methods are those of the reference's
// export (src/BussiReservoirThermostat.cc:9-30); `fused_rescale` is the one addition.
#include "BussiReservoirThermostat.h"

namespace hoomd::md
    {
namespace
    {
using Self = BussiReservoirThermostat;

// the six reservoir read-outs share one signature: register them from a table
struct Getter
    {
    const char* name;
    Scalar (Self::*fn)();
    };
const Getter reservoir_getters[] = {
    {"getReservoirEnergyTranslational", &Self::getReservoirEnergyTranslational},
    {"getReservoirEnergyRotational", &Self::getReservoirEnergyRotational},
    {"getTotalReservoirEnergy", &Self::getTotalReservoirEnergy},
    {"getInstantaneousReservoirTranslational", &Self::getInstantaneousReservoirTranslational},
    {"getInstantaneousReservoirRotational", &Self::getInstantaneousReservoirRotational},
    {"getInstantaneousReservoirTotal", &Self::getInstantaneousReservoirTotal},
};
    } // namespace

void export_BussiReservoirThermostat(pybind11::module& m)
    {
    namespace py = pybind11;
    py::class_<Self, Thermostat, std::shared_ptr<Self>> cls(m, "BussiReservoirThermostat");
    cls.def(py::init<std::shared_ptr<Variant>, std::shared_ptr<ParticleGroup>, std::shared_ptr<ComputeThermo>,
                     std::shared_ptr<SystemDefinition>, Scalar>());
    cls.def_property("kT", &Self::getT, &Self::setT);
    cls.def_property("tau", &Self::getTau, &Self::setTau);
    cls.def_property("fused_rescale", &Self::getFusedRescale, &Self::setFusedRescale);
    cls.def_property("cooperative_launch", &Self::getCooperativeLaunch, &Self::setCooperativeLaunch);
    cls.def("getFaultCount", &Self::getFaultCount);
    for (const Getter& g : reservoir_getters)
        cls.def(g.name, g.fn);
    cls.def("resetReservoirEnergy", &Self::resetReservoirEnergy);
    }
    } // namespace hoomd::md
