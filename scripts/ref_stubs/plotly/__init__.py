"""Stand-in for `plotly` (not installed here): every attribute is a no-op."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _noop import Noop  # noqa: E402

for _sub in ("graph_objects", "express", "io", "subplots", "offline"):
    _m = types.ModuleType(f"plotly.{_sub}")
    _m.__getattr__ = lambda name: Noop()
    sys.modules[f"plotly.{_sub}"] = _m
    globals()[_sub] = _m


def __getattr__(name):
    return Noop()
