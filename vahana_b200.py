"""Import shim: the package directory is named `vahana.jl_b200` (not a valid Python identifier), so it
is loaded by path and exposed as the module `vahana_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_p = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vahana.jl_b200", "__init__.py")
_spec = _u.spec_from_file_location("vahana_b200", _p, submodule_search_locations=[_os.path.dirname(_p)])
_m = _u.module_from_spec(_spec)
_sys.modules["vahana_b200"] = _m
_spec.loader.exec_module(_m)
