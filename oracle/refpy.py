"""oracle/refpy.py -- TEST INFRASTRUCTURE: import the reference's own Python modules by path.

The reference's Python side (src/cavitymd/analysis.py, utils.py, cavity_force_python.py) imports
`hoomd` at module top, and HOOMD-blue is not in this image.  The functions this repo restates
(compute_density_field, generate_fibonacci_sphere, compute_field_autocorr,
compute_total_dipole_moment, unwrap_positions, CavityForcePython.set_forces) only touch NumPy, so a
stub `hoomd` module with the handful of names those files reference at import time (custom.Action,
logging.log, error.DataAccessError, md.force.Custom / Force) is enough to run the reference's code
UNMODIFIED, from where it lies under /root/reference.  Nothing is copied.

Used by tests/test_reference_python.py and tests/golden/make_golden.py (the fixtures are minted from
the imported reference).  /root/reference does not exist on the GPU box: `available()` is False
there and the tests that need it skip; the committed fixtures carry the pin.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CAVB_REFERENCE_ROOT", "/root/reference")
_PKG = "_cavref_cavitymd"
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "cavitymd", "analysis.py"))


def _stub_hoomd() -> types.ModuleType:
    """The names the reference's files look up on `hoomd` while being imported, and nothing more."""
    hoomd = types.ModuleType("hoomd")

    def log(*args, **kwargs):  # used both as @log and as @log(requires_run=True)
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return property(args[0])
        return lambda fn: property(fn)

    class Action:  # hoomd.custom.Action
        def attach(self, simulation):
            self._state = simulation.state

        def detach(self):
            pass

    class DataAccessError(RuntimeError):
        pass

    class _ForceArrays:
        def __init__(self, owner):
            self.owner = owner

        def __enter__(self):
            return self.owner._arrays

        def __exit__(self, *a):
            return False

    class Custom:  # hoomd.md.force.Custom: the two members CavityForcePython uses
        def __init__(self, aniso=False):
            self._state = None
            self._arrays = None

        @property
        def cpu_local_force_arrays(self):
            return _ForceArrays(self)

    class Force:
        def __init__(self):
            self._param_dict = {}

    hoomd.custom = types.ModuleType("hoomd.custom")
    hoomd.custom.Action = Action
    hoomd.logging = types.ModuleType("hoomd.logging")
    hoomd.logging.log = log
    hoomd.error = types.ModuleType("hoomd.error")
    hoomd.error.DataAccessError = DataAccessError
    hoomd.md = types.ModuleType("hoomd.md")
    hoomd.md.force = types.ModuleType("hoomd.md.force")
    hoomd.md.force.Custom = Custom
    hoomd.md.force.Force = Force
    return hoomd


def load():
    """-> namespace with .analysis, .utils, .cavity_force_python = the reference's modules, executed from
    /root/reference/src/cavitymd/*.py under the stub.  Raises FileNotFoundError without the tree."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"{REF_ROOT}/src/cavitymd not present (GPU box?): use the committed fixtures")
    src = os.path.join(REF_ROOT, "src", "cavitymd")
    saved = {k: sys.modules.get(k) for k in ("hoomd", "hoomd.custom", "hoomd.logging", "hoomd.error", "hoomd.md",
                                             "hoomd.md.force")}
    stub = _stub_hoomd()
    sys.modules.update({"hoomd": stub, "hoomd.custom": stub.custom, "hoomd.logging": stub.logging,
                        "hoomd.error": stub.error, "hoomd.md": stub.md, "hoomd.md.force": stub.md.force})
    try:
        # a package shell with the reference directory as its path, so that the files' relative imports
        # (`from .utils import ...`) resolve WITHOUT running cavitymd/__init__.py (which pulls in the simulation driver)
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [src]
        sys.modules[_PKG] = pkg
        mods = {}
        for name in ("utils", "analysis", "cavity_force_python"):
            spec = importlib.util.spec_from_file_location(f"{_PKG}.{name}", os.path.join(src, f"{name}.py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = m
            spec.loader.exec_module(m)
            mods[name] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache = types.SimpleNamespace(**mods, hoomd_stub=stub)
    return _cache


class Snapshot:
    """What the reference's observables read from `sim.state.cpu_local_snapshot`: particles.{position, image,
    charge, typeid, velocity, mass} and global_box.L."""

    def __init__(self, position, image=None, charge=None, typeid=None, velocity=None, mass=None, box=None):
        self.particles = types.SimpleNamespace(position=position, image=image, charge=charge, typeid=typeid,
                                               velocity=velocity, mass=mass)
        self.global_box = types.SimpleNamespace(L=box)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def python_cavity_force(position, image, charge, typeid, box, omegac, couplstr, phmass=1.0):
    """Run the reference's CavityForcePython.set_forces (src/cavitymd/cavity_force_python.py:65-149) on one frame.
    NB: that class takes typeid == 1 as the cavity particle (:75), unlike the C++ class (type name 'L')."""
    import contextlib
    import io

    import numpy as np
    ref = load()
    with contextlib.redirect_stdout(io.StringIO()):
        f = ref.cavity_force_python.CavityForcePython([0, 0, 1], couplstr, omegac, phmass)
        n = len(position)
        f._arrays = types.SimpleNamespace(force=np.zeros((n, 3)), potential_energy=np.zeros(n))
        f._state = types.SimpleNamespace(cpu_local_snapshot=Snapshot(position, image, charge, typeid, box=box))
        f.set_forces(0)
    return dict(force=f._arrays.force, energies=np.array([f.harmonic_energy, f.coupling_energy, f.dipole_self_energy]))
