"""Importable alias for the package directory `joint-tensorf_b200/`.

The package directory carries the project's hyphenated name, which is not a
valid Python identifier; this shim loads it under the module name
`joint_tensorf_b200` (sub-modules resolve inside the hyphenated directory).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "joint-tensorf_b200")
_spec = importlib.util.spec_from_file_location(
    "joint_tensorf_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["joint_tensorf_b200"] = _mod
_spec.loader.exec_module(_mod)
