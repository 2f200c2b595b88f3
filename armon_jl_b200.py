"""Import shim: `import armon_jl_b200 as armon` loads the package directory `armon.jl_b200/`.

The package directory is named after the reference (`armon.jl_b200`), which is not a valid Python
identifier; this module loads it under the importable name `armon_jl_b200`.
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "armon.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
