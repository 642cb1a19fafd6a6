// _bussi_reservoir: same module and class name as the reference (reference src/bussi_reservoir/module.cc:20-23).
#include "BussiReservoirThermostat.h"
#ifdef HOOMD_SHIM_CORE_H
#include <pybind11/stl.h>
#endif

PYBIND11_MODULE(_bussi_reservoir, m)
    {
    hoomd::md::export_BussiReservoirThermostat(m);
#ifdef HOOMD_SHIM_CORE_H
    // shim test build only: the shim's RandomGenerator hands out injected draws, and each extension
    // module has its own copy of that queue (extension modules are loaded RTLD_LOCAL)
    m.def("_inject_draws",
          [](std::vector<double> v)
          {
              auto& q = hoomd::RandomGenerator::injected();
              q.clear();
              for (double x : v)
                  q.push_back(x);
          });
#endif
    }
