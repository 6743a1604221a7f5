"""Import alias: ``import mmf`` == the package in ``dss-ml-at-scale_b200/`` (whose directory
name is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("dss-ml-at-scale_b200")
sys.modules[__name__] = _pkg
