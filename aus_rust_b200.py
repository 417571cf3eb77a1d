"""Import shim: the package directory is ``aus-rust_b200/`` (not a valid Python identifier), so this module
turns itself into that package.  ``import aus_rust_b200`` then behaves like importing the directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "aus-rust_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
