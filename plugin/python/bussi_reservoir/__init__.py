"""hoomd.bussi_reservoir, B200 build."""
from .thermostats import BussiReservoir

__all__ = ["BussiReservoir"]
