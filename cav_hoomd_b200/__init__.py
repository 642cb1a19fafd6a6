"""cav_hoomd_b200 -- the B200 (sm_100a) hot path of cav-hoomd: cavity force, Bussi reservoir
thermostat and F(k,t) density field as hand-written CUDA behind a C ABI (include/cavb200.h).

Only what the path needs lives here: csrc/ (kernels + C ABI), capi.py (ctypes binding), and the
host-side mirrors of the reference's Python interface (forces.py, thermostats.py, analysis.py)."""
from .forces import CavityForce  # noqa: F401
from .thermostats import BussiReservoir  # noqa: F401
from .analysis import DensityField, FieldAutocorrelationTracker, generate_fibonacci_sphere  # noqa: F401
from .state import DeviceState  # noqa: F401
from .replicas import parse_replicas  # noqa: F401

__all__ = ["CavityForce", "BussiReservoir", "DensityField", "FieldAutocorrelationTracker",
           "generate_fibonacci_sphere", "DeviceState", "parse_replicas"]
