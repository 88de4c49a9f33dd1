"""Importable alias of the ``nemo-fmi-devel_b200`` package (its prescribed directory name has a hyphen)."""
import importlib
import sys

_pkg = importlib.import_module("nemo-fmi-devel_b200")
sys.modules[__name__] = _pkg
