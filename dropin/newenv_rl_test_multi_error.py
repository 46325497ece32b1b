"""Drop-in for the reference module of the same name (reference newenv_rl_test_multi_error.py).

Put this directory in front of the reference on PYTHONPATH and ``from newenv_rl_test_multi_error
import HelioField`` (test_environment.py:3, README.md:63) resolves to the sm_100a implementation.
"""
from doodle_b200.field import HelioField  # noqa: F401
