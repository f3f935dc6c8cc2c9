"""Import shim: `import mgbx` loads the package directory `multigridbarrier.jl_b200/` (whose name
is not a valid Python identifier) under the module name `mgbx`."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multigridbarrier.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "mgbx", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_m = importlib.util.module_from_spec(_spec)
sys.modules["mgbx"] = _m
_spec.loader.exec_module(_m)
