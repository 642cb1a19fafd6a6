#!/usr/bin/env python
"""tests/golden/make_wrapper_contract.py -- which names the REFERENCE'S Python wrappers look up on the compiled modules.

A B200 deployment ships the reference's src/cavitymd/forces.py and src/bussi_reservoir/thermostats.py unmodified
(plugin/CMakeLists.txt), so the pybind classes of plugin/src must carry every name those files use: the classes they
construct (with how many positional arguments) and the methods they call on the constructed object.  This script reads
the two files with `ast` from /root/reference and writes tests/golden/wrapper_contract.json;
tests/test_wrapper_contract.py checks the B200 modules against it (and, where /root/reference exists, that the
fixture is current)."""
import ast
import json
import os

REF = os.environ.get("CAVB_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {"cavitymd": ("src/cavitymd/forces.py", "_cavitymd", ("_force_impl", "_cpp_obj")),
         "bussi_reservoir": ("src/bussi_reservoir/thermostats.py", "_bussi_reservoir", ("_cpp_obj",))}


def extract(path, module_name, holders):
    tree = ast.parse(open(path).read())
    constructed, methods = {}, set()
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call) or not isinstance(node.func, ast.Attribute):
            continue
        f = node.func
        # _cavitymd.CavityForceComputeGPU(...)  /  _bussi_reservoir.BussiReservoirThermostat(...)
        if isinstance(f.value, ast.Name) and f.value.id == module_name:
            constructed[f.attr] = max(constructed.get(f.attr, 0), len(node.args))
        # self._force_impl.getHarmonicEnergy()  /  self._cpp_obj.resetReservoirEnergy()
        if isinstance(f.value, ast.Attribute) and f.value.attr in holders and isinstance(f.value.value, ast.Name) \
                and f.value.value.id == "self" and not f.attr.startswith("_"):
            methods.add(f.attr)
    return {"constructs": constructed, "calls": sorted(methods)}


def build():
    return {k: dict(file=rel, **extract(os.path.join(REF, rel), mod, holders)) for k, (rel, mod, holders) in FILES.items()}


if __name__ == "__main__":
    out = build()
    with open(os.path.join(HERE, "wrapper_contract.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
