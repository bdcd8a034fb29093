"""Importable alias of the package directory `datafusion-bio-functions_b200/` (a hyphen is not a valid identifier):
    import bio_ranges_b200 as brg
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("datafusion-bio-functions_b200")
sys.modules[__name__] = _pkg
