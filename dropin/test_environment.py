"""Drop-in for the reference module of the same name (reference test_environment.py).

``from test_environment import HelioEnv`` (train_with_env.py:21, env_sanity_check.py:7) resolves to
the sm_100a implementation when this directory precedes the reference on PYTHONPATH.
"""
from doodle_b200.env import (HelioEnv, azimuth_elevation_to_primary_direction, make_distance_maps,  # noqa: F401
                             sample_cone_directions)
from doodle_b200.field import HelioField  # noqa: F401
