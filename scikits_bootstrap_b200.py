"""Import shim: the package directory is named `scikits-bootstrap_b200/` (not a
valid Python identifier), so this module loads it under the importable name
`scikits_bootstrap_b200` and replaces itself with it in sys.modules."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "scikits-bootstrap_b200")
_spec = importlib.util.spec_from_file_location(
    "scikits_bootstrap_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["scikits_bootstrap_b200"] = _mod
_spec.loader.exec_module(_mod)
