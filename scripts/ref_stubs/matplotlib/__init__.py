"""Stand-in for `matplotlib` (not installed here): `matplotlib.pyplot` and friends are no-ops."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _noop import Noop  # noqa: E402

for _sub in ("pyplot", "cm", "colors", "patches", "gridspec"):
    _m = types.ModuleType(f"matplotlib.{_sub}")
    _m.__getattr__ = lambda name: Noop()
    sys.modules[f"matplotlib.{_sub}"] = _m
    globals()[_sub] = _m


def use(*a, **k):
    pass


def __getattr__(name):
    return Noop()
