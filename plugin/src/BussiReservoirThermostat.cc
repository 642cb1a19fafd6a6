// BussiReservoirThermostat.cc -- pybind11 export, same names as the reference
// (reference src/BussiReservoirThermostat.cc:9-30) plus the opt-in fused_rescale property.
#include "BussiReservoirThermostat.h"

namespace hoomd::md
    {
void export_BussiReservoirThermostat(pybind11::module& m)
    {
    pybind11::class_<BussiReservoirThermostat, Thermostat, std::shared_ptr<BussiReservoirThermostat>>(
        m, "BussiReservoirThermostat")
        .def(pybind11::init<std::shared_ptr<Variant>, std::shared_ptr<ParticleGroup>, std::shared_ptr<ComputeThermo>,
                            std::shared_ptr<SystemDefinition>, Scalar>())
        .def_property("tau", &BussiReservoirThermostat::getTau, &BussiReservoirThermostat::setTau)
        .def_property("kT", &BussiReservoirThermostat::getT, &BussiReservoirThermostat::setT)
        .def_property("fused_rescale", &BussiReservoirThermostat::getFusedRescale, &BussiReservoirThermostat::setFusedRescale)
        .def("getReservoirEnergyTranslational", &BussiReservoirThermostat::getReservoirEnergyTranslational)
        .def("getReservoirEnergyRotational", &BussiReservoirThermostat::getReservoirEnergyRotational)
        .def("getTotalReservoirEnergy", &BussiReservoirThermostat::getTotalReservoirEnergy)
        .def("getInstantaneousReservoirTranslational", &BussiReservoirThermostat::getInstantaneousReservoirTranslational)
        .def("getInstantaneousReservoirRotational", &BussiReservoirThermostat::getInstantaneousReservoirRotational)
        .def("getInstantaneousReservoirTotal", &BussiReservoirThermostat::getInstantaneousReservoirTotal)
        .def("resetReservoirEnergy", &BussiReservoirThermostat::resetReservoirEnergy);
    }
    } // namespace hoomd::md
