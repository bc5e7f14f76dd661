"""Import shim: the package directory `0g-ec-gpu_b200` is not a Python identifier."""
import importlib
import sys

_pkg = importlib.import_module("0g-ec-gpu_b200")
sys.modules[__name__] = _pkg
