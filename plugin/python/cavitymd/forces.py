"""hoomd.cavitymd.CavityForce for a real HOOMD install, B200 build.

Same constructor and properties as the reference wrapper (reference src/cavitymd/forces.py:21-233);
the only implementation is _cavitymd.CavityForceComputeGPU (libcavb200).  No C++/CPU or Python
fallback: attaching on a CPU device, or asking for force_python=True, raises.
(Untested in the development image, which has no HOOMD; the pybind class underneath is tested
against hoomd_shim in tests/test_plugin_gpu.py.)"""
import hoomd
import numpy as np
from hoomd.logging import log

from . import _cavitymd


class CavityForce(hoomd.md.force.Force):
    def __init__(self, kvector, couplstr, omegac, phmass=1.0, force_python=False):
        super().__init__()
        if force_python:
            raise NotImplementedError("the B200 build has no Python fallback")
        param_dict = hoomd.data.parameterdicts.ParameterDict(
            kvector=hoomd.data.typeconverter.to_type_converter([float, float, float]),
            couplstr=float, omegac=float, phmass=float, force_python=bool)
        param_dict.update(dict(kvector=list(kvector), couplstr=couplstr, omegac=omegac, phmass=phmass,
                               force_python=False))
        self._param_dict.update(param_dict)
        self.kvector = np.array(kvector)
        self.couplstr, self.omegac, self.phmass = couplstr, omegac, phmass
        self._implementation = "cuda"

    def _attach_hook(self):
        if not isinstance(self._simulation.device, hoomd.device.GPU):
            raise RuntimeError("hoomd.cavitymd (B200 build) needs hoomd.device.GPU: there is no CPU fallback")
        # note the pybind argument order: omegac before couplstr (reference CavityForceCompute.cc:215-218)
        self._cpp_obj = _cavitymd.CavityForceComputeGPU(self._simulation.state._cpp_sys_def, self.omegac,
                                                        self.couplstr, self.phmass)
        super()._attach_hook()

    @property
    def implementation(self):
        return self._implementation

    @log(requires_run=True)
    def harmonic_energy(self):
        return self._cpp_obj.getHarmonicEnergy()

    @log(requires_run=True)
    def coupling_energy(self):
        return self._cpp_obj.getCouplingEnergy()

    @log(requires_run=True)
    def dipole_self_energy(self):
        return self._cpp_obj.getDipoleSelfEnergy()

    @log(requires_run=True)
    def total_cavity_energy(self):
        return self.harmonic_energy + self.coupling_energy + self.dipole_self_energy

    @property
    def energy(self):
        return self.total_cavity_energy
