"""Device-resident particle state in HOOMD's array layouts.

Stand-in for the part of hoomd.State / ParticleData the hot path touches (HOOMD itself is not in
this image).  A real deployment does not use this module: plugin/ binds the same C ABI to HOOMD's
own GlobalArrays."""
from __future__ import annotations

import numpy as np

from . import capi, synth


class DeviceState:
    """pos / vel / charge / image / force on one GPU (Scalar4 / Scalar / int3 layouts)."""

    def __init__(self, system: synth.System, device: int = 0):
        self.device = device
        self.N = system.N
        self.box = tuple(system.box)
        self.types = tuple(system.types)
        self.pos = capi.DeviceArray.from_numpy(system.pos)
        self.vel = capi.DeviceArray.from_numpy(system.vel)
        self.charge = capi.DeviceArray.from_numpy(system.charge)
        self.image = capi.DeviceArray.from_numpy(system.image)
        self.force = capi.DeviceArray((self.N, 4), np.float64)
        self.force.fill_bytes(0)
        self.timestep = 0
        self.seed = 0

    def type_id(self, name: str) -> int:
        """getTypeByName: raises like HOOMD for an unknown name (reference CavityForceCompute.cc:79)."""
        try:
            return self.types.index(name)
        except ValueError:
            raise RuntimeError(f"Type {name} not found!") from None

    def snapshot(self) -> synth.System:
        return synth.System(self.pos.numpy(), self.vel.numpy(), self.charge.numpy(), self.image.numpy(), self.box,
                            self.types.index("L") if "L" in self.types else -1, self.types)
