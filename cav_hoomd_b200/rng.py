"""Counter-based random draws for the Bussi thermostat on the host.

The reference draws from HOOMD's RandomGenerator(Seed(RNGIdentifier::BussiThermostat, timestep,
seed), instance) with HOOMD's Normal (Box-Muller) and Gamma (Marsaglia-Tsang) samplers
(reference src/BussiReservoirThermostat.h:63-67,192-199).  Those are HOOMD upstream code that is
not in the reference tree, so bit parity of the random STREAM cannot be pinned here:
in a HOOMD build plugin/ uses HOOMD's generator; this stand-alone host layer uses a Philox4x64
stream keyed the same way (identifier, timestep, seed, instance), one normal then one gamma, in the
reference's draw order."""
from __future__ import annotations

import numpy as np

BUSSI_THERMOSTAT_ID = 0x6A1  # any fixed identifier distinct from other consumers


def bussi_draws(timestep: int, seed: int, instance: int, dof: float):
    """-> (r_normal, gamma_draw): normal(0,1) first, then Gamma((dof-1)/2, 1) only if dof > 1."""
    key = (int(seed) & 0xFFFF) | (BUSSI_THERMOSTAT_ID << 16) | ((int(instance) & 0xFFFFFFFF) << 32)
    bitgen = np.random.Philox(key=key, counter=[int(timestep) & 0xFFFFFFFFFFFFFFFF, 0, 0, 0])
    g = np.random.Generator(bitgen)
    r_normal = float(g.standard_normal())
    gamma_draw = float(g.gamma((dof - 1.0) / 2.0)) if dof > 1.0 else 0.0
    return r_normal, gamma_draw
