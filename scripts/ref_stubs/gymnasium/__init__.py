"""Stand-in for `gymnasium` (not installed here): the three names the reference touches (test_environment.py:11-12,241-252)."""
import sys, types


class Env:
    pass


spaces = types.ModuleType("gymnasium.spaces")


class _Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class _Dict(dict):
    def __init__(self, d):
        super().__init__(d)


spaces.Box, spaces.Dict = _Box, _Dict
sys.modules["gymnasium.spaces"] = spaces
