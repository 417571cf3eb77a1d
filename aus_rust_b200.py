"""Import shim: the package directory is ``aus-rust_b200/`` (not a valid Python identifier), so this module
replaces itself in ``sys.modules`` with that directory loaded as the package ``aus_rust_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "aus-rust_b200")
_spec = _ilu.spec_from_file_location("aus_rust_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["aus_rust_b200"] = _mod
_spec.loader.exec_module(_mod)
