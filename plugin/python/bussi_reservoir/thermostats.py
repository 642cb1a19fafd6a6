"""hoomd.bussi_reservoir.BussiReservoir for a real HOOMD install, B200 build: same constructor,
loggables and reset method as the reference (reference src/bussi_reservoir/thermostats.py:14-158),
plus the opt-in `fused_rescale` switch documented in plugin/src/BussiReservoirThermostat.h.
(Untested in the development image, which has no HOOMD.)"""
import hoomd
from hoomd.data.parameterdicts import ParameterDict
from hoomd.md.methods import thermostats


class BussiReservoir(thermostats.Thermostat):
    def __init__(self, kT, tau=0.0, fused_rescale=False):
        super().__init__(kT)
        param_dict = ParameterDict(tau=float, fused_rescale=bool)
        param_dict["tau"] = tau
        param_dict["fused_rescale"] = fused_rescale
        self._param_dict.update(param_dict)

    def _attach_hook(self):
        from . import _bussi_reservoir
        group = self._simulation.state._get_group(self._filter)
        self._cpp_obj = _bussi_reservoir.BussiReservoirThermostat(self.kT, group, self._thermo,
                                                                  self._simulation.state._cpp_sys_def, self.tau)
        self._cpp_obj.fused_rescale = self.fused_rescale
        self._simulation._warn_if_seed_unset()

    def _get(self, name):
        return getattr(self._cpp_obj, name)() if self._attached else 0.0

    @hoomd.logging.log()
    def reservoir_energy_translational(self):
        return self._get("getReservoirEnergyTranslational")

    @hoomd.logging.log()
    def reservoir_energy_rotational(self):
        return self._get("getReservoirEnergyRotational")

    @hoomd.logging.log()
    def total_reservoir_energy(self):
        return self._get("getTotalReservoirEnergy")

    @hoomd.logging.log()
    def instantaneous_reservoir_translational(self):
        return self._get("getInstantaneousReservoirTranslational")

    @hoomd.logging.log()
    def instantaneous_reservoir_rotational(self):
        return self._get("getInstantaneousReservoirRotational")

    @hoomd.logging.log()
    def instantaneous_reservoir_total(self):
        return self._get("getInstantaneousReservoirTotal")

    def reset_reservoir_energy(self):
        if self._attached:
            self._cpp_obj.resetReservoirEnergy()
