"""Import shim: exposes the package in `raytracing-rust_b200/` (not a valid Python identifier) as `ptb200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raytracing-rust_b200")
_spec = importlib.util.spec_from_file_location("ptb200", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ptb200"] = _mod
_spec.loader.exec_module(_mod)
