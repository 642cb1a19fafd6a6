"""hoomd.cavitymd, B200 build: CavityForce backed by libcavb200 (see forces.py)."""
from .forces import CavityForce

__all__ = ["CavityForce"]
