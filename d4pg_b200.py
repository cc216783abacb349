"""Importable alias for the `d4pg-pytorch_b200/` package (its directory name has a hyphen)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("d4pg-pytorch_b200")
sys.modules[__name__] = _pkg
