"""hoomd.bussi_reservoir.BussiReservoir, B200 build -- host-side mirror of the reference wrapper
(reference src/bussi_reservoir/thermostats.py:14-158 and src/BussiReservoirThermostat.h).

KE reduce, alpha, reservoir bookkeeping and v <- alpha v run in ONE launch (cavb200_bussi); the two
random draws are made here BEFORE the launch because they do not depend on the kinetic energy."""
from __future__ import annotations

import numpy as np

from . import capi, rng
from .state import DeviceState


class BussiReservoir:
    """BussiReservoir(kT, tau=0.0) (reference thermostats.py:81-85)."""

    def __init__(self, kT, tau=0.0):
        self._kT = kT
        self.tau = float(tau)
        self._state = None
        self._handle = None
        self._group = None
        self._attached = False

    @property
    def kT(self):
        return self._kT

    @kT.setter
    def kT(self, v):
        self._kT = v

    def _set_T(self, timestep):
        return float(self._kT(timestep)) if callable(self._kT) else float(self._kT)

    # -- attach (reference thermostats.py:87-93: group from the filter, ComputeThermo, sysdef) ------
    def _attach(self, state: DeviceState, group: np.ndarray | None = None, dof: float | None = None,
                handle: capi.Handle | None = None):
        """group: particle indices (None = every particle that is not of type 'L', the reference's
        filter.Type(['O','N'])).  dof: translational degrees of freedom (default 3 n - 3)."""
        self._state = state
        self._handle = handle or capi.Handle(state.device)
        self._owns_handle = handle is None
        if group is None:
            from .synth import w_to_typeid
            tid = w_to_typeid(state.pos.numpy()[:, 3])
            L = state.types.index("L") if "L" in state.types else -1
            group = np.nonzero(tid != L)[0].astype(np.uint32)
        group = np.ascontiguousarray(group, dtype=np.uint32)
        self._n = len(group)
        self._instance = int(group[0]) if self._n else 0  # tag of group member 0 (reference .h:63-65)
        contiguous = self._n == 0 or np.array_equal(group, np.arange(group[0], group[0] + self._n, dtype=np.uint32))
        self._first = int(group[0]) if (contiguous and self._n) else 0
        self._d_group = None if contiguous else capi.DeviceArray.from_numpy(group)
        self.dof = float(dof) if dof is not None else max(3.0 * self._n - 3.0, 0.0)
        self._handle.bussi_reset()
        self._attached = True
        return self

    def _detach(self):
        if self._handle is not None and getattr(self, "_owns_handle", False):
            self._handle.close()
        self._handle = None
        self._attached = False

    # -- getRescalingFactorsOne + the rescale HOOMD's step one applies ---------------------------------
    def rescale(self, timestep: int, deltaT: float, stream=None, draws=None):
        """One thermostat step.  `draws` = (r_normal, gamma_draw) overrides the host RNG (tests)."""
        if not self._attached:
            raise RuntimeError("BussiReservoir is not attached")
        if deltaT == 0.0:  # reference .h:45-48
            return
        r_normal, gamma_draw = draws if draws is not None else rng.bussi_draws(timestep, self._state.seed,
                                                                               self._instance, self.dof)
        args = capi.BussiArgs(self._set_T(timestep), self.tau, float(deltaT), self.dof, r_normal, gamma_draw)
        self._handle.bussi(self._state.vel, self._d_group, self._first, self._n, args, stream)

    def _read(self):
        out = self._handle.bussi_read()
        if out["err"] == 1.0:
            raise RuntimeError("Bussi thermostat requires non-zero initial momenta.")  # reference .h:57-61
        return out

    # -- the six loggable quantities + reset (reference thermostats.py:95-158) --------------------------
    @property
    def reservoir_energy_translational(self):
        return 0.0 if not self._attached else self._read()["cumulative"]

    @property
    def reservoir_energy_rotational(self):
        return 0.0  # point particles: rotational dof = 0 -> factor 1, delta 0 (reference .h:77-87)

    @property
    def total_reservoir_energy(self):
        return self.reservoir_energy_translational + self.reservoir_energy_rotational

    @property
    def instantaneous_reservoir_translational(self):
        return 0.0 if not self._attached else self._read()["instantaneous"]

    @property
    def instantaneous_reservoir_rotational(self):
        return 0.0

    @property
    def instantaneous_reservoir_total(self):
        return self.instantaneous_reservoir_translational + self.instantaneous_reservoir_rotational

    @property
    def last_alpha(self):
        return self._read()["alpha"]

    @property
    def last_kinetic_energy(self):
        return self._read()["ke"]

    def reset_reservoir_energy(self):
        """No-op when not attached (reference thermostats.py:137-158)."""
        if self._attached:
            self._handle.bussi_reset()
