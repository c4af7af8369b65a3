"""Import shim: the package directory is named `integrals.jl_b200/` (after the reference,
Integrals.jl), which is not a valid Python identifier; this module loads it under the importable
name `integrals_b200`.  `import integrals_b200 as ib` is the public entry point for Python callers.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "integrals.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "integrals_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["integrals_b200"] = _mod
_spec.loader.exec_module(_mod)
