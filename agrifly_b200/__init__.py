"""Importable alias of the `agri-fly_b200` package (its directory name carries the reference's
hyphen, which the `import` statement cannot spell)."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("agri-fly_b200")
_sys.modules[__name__] = _pkg
