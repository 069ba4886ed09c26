"""Import alias: `import mavlm_b200` == the package in `memory-augmented-vlm_b200/` (whose directory
name, fixed by the project layout, is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("memory-augmented-vlm_b200")
sys.modules[__name__] = _pkg
for _name in ("_lib", "ops", "modules", "pipeline", "synthetic", "dist", "autograd", "splice", "patch", "preprocess"):
    importlib.import_module(f"memory-augmented-vlm_b200.{_name}")
for _name in ("_lib", "ops", "modules", "pipeline", "synthetic", "dist", "autograd", "splice", "patch", "preprocess"):
    _m = sys.modules.get(f"memory-augmented-vlm_b200.{_name}")
    if _m is not None:
        sys.modules[f"{__name__}.{_name}"] = _m
