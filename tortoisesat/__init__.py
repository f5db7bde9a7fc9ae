"""Import shim: makes ``import tortoisesat.jl_b200`` resolve to the package that
lives in the directory ``tortoisesat.jl_b200/`` at the repo root (a dotted
directory name cannot be imported directly)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tortoisesat.jl_b200")
_spec = _ilu.spec_from_file_location(
    "tortoisesat.jl_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
jl_b200 = _ilu.module_from_spec(_spec)
_sys.modules["tortoisesat.jl_b200"] = jl_b200
_spec.loader.exec_module(jl_b200)
