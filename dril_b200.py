"""Import shim: the package directory is literally named `dril.jl_b200/` (not a valid dotted
module name), so it is loaded from its path and registered as `dril_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dril.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "dril_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dril_b200"] = _mod
_spec.loader.exec_module(_mod)
