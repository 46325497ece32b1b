"""Stand-in for `mlflow` (not installed here; the reference logs to a remote tracking server): every call is a no-op."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _noop import Noop  # noqa: E402


def __getattr__(name):
    return Noop()
