"""Importable alias of the package directory `3d-shape-generation_b200/` (whose name is not a
valid Python identifier): `import pcd_b200` returns that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("3d-shape-generation_b200")
sys.modules[__name__] = _pkg
