"""hoomd.cavitymd.CavityForce, B200 build -- host-side mirror of the reference wrapper.

Same constructor, same properties, same meaning as reference src/cavitymd/forces.py:21-233; the
implementation behind it is libcavb200 (one cooperative sm_100a kernel per call), never C++/CPU or
Python: `force_python=True` raises instead of selecting a fallback (BASELINE.json north_star).
In a HOOMD install the class of the same name in plugin/python/cavitymd/forces.py derives from
hoomd.md.force.Force; here (no HOOMD in the image) it attaches to a DeviceState."""
from __future__ import annotations

import numpy as np

from . import capi
from .state import DeviceState


class CavityForce:
    """H = 1/2 K q^2 + g q.d + (g^2 / 2K) d^2,  K = phmass * omegac^2  (reference forces.py:21-43).

    Parameters (reference forces.py:45): kvector (stored, unused -- forces.py:33-34), couplstr g,
    omegac, phmass=1.0, force_python=False."""

    def __init__(self, kvector, couplstr, omegac, phmass=1.0, force_python=False):
        if force_python:
            raise NotImplementedError("the B200 build has no Python/CPU fallback (force_python=True is refused)")
        self.kvector = np.array(kvector, dtype=np.float64)
        if self.kvector.shape != (3,):
            raise ValueError("kvector must have three components")
        self.couplstr = float(couplstr)
        self.omegac = float(omegac)
        self.phmass = float(phmass)
        self._implementation = "cuda"
        self._state = None
        self._handle = None
        self._params = None
        self._L_typeid = 0xFFFFFFFF
        self._ran = False

    # -- attach / detach (reference _attach_hook, forces.py:97-173) -------------------------------
    def _attach(self, state: DeviceState, handle: capi.Handle | None = None):
        self._state = state
        self._handle = handle or capi.Handle(state.device)
        self._owns_handle = handle is None
        # pybind ctor order is (sysdef, omegac, couplstr, phmass) -- reference CavityForceCompute.cc:215-218
        self._params = capi.Params.make(self.omegac, self.couplstr, self.phmass)
        try:
            self._L_typeid = state.type_id("L")
        except RuntimeError:
            # the reference GPU class swallows the missing type and reports zero energies
            # (src/CavityForceComputeGPU.cc:114-123)
            self._L_typeid = 0xFFFFFFFF
        return self

    def _detach(self):
        if self._handle is not None and getattr(self, "_owns_handle", False):
            self._handle.close()
        self._handle = None
        self._state = None

    @property
    def implementation(self):
        """'cuda' -- the only implementation of this build (reference forces.py:175-178)."""
        return self._implementation

    # -- ForceCompute::compute(timestep) -----------------------------------------------------------
    def compute(self, timestep: int = 0, stream=None):
        if self._state is None:
            raise RuntimeError("CavityForce is not attached to a state")
        s = self._state
        self._handle.force(s.pos, s.charge, s.image, s.force, s.N, s.box, self._L_typeid, self._params, stream)
        self._ran = True

    def _energies(self):
        if not self._ran:
            raise RuntimeError("requires_run: no force computation has happened yet")
        en, _, _ = self._handle.force_read()
        return en

    # lazy getters: one small D2H copy when asked, nothing per step
    @property
    def harmonic_energy(self):
        """(1/2) K q^2 (reference forces.py:180-186)."""
        return float(self._energies()[0])

    @property
    def coupling_energy(self):
        """g (q . d) (reference forces.py:188-194)."""
        return float(self._energies()[1])

    @property
    def dipole_self_energy(self):
        """(g^2 / 2K) d^2 (reference forces.py:196-202)."""
        return float(self._energies()[2])

    @property
    def total_cavity_energy(self):
        """Sum of the three components, in that order (reference forces.py:204-207)."""
        e = self._energies()
        return float(e[0] + e[1] + e[2])

    @property
    def energy(self):
        """Overrides Force.energy with the component sum (reference forces.py:209-212)."""
        return self.total_cavity_energy

    @property
    def dipole(self):
        """Total molecular dipole d (device scalar block; not in the reference API)."""
        if not self._ran:
            raise RuntimeError("requires_run")
        return self._handle.force_read()[1]

    @property
    def forces(self):
        """Per-particle forces float64[N,3].  (The reference returns None for the C++ implementation,
        forces.py:214-221, which makes AdaptiveTimestepUpdater skip the cavity force silently.)"""
        if not self._ran:
            raise RuntimeError("requires_run")
        return self._state.force.numpy()[:, :3]
