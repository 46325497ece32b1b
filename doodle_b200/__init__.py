"""doodle_b200 -- B200-native (sm_100a) implementation of DOODLE's differentiable flux renderer.

Scope: the one hot path of l3th4l/DOODLE -- ``HelioField.render`` forward + backward and the
``HelioEnv.step`` loss block -- behind the reference's own Python API.  See DESIGN.md.

Importing the package does not touch CUDA; constructing a field / env does, and raises if the
extension (doodle_b200/libhelio_sm100.so) or a cc-10.x GPU is missing.  There is no CPU fallback.
"""
from ._lib import HelioLibError, SPLAT_AUTO, SPLAT_SIMT, SPLAT_TC  # noqa: F401
from .env import (HelioEnv, angles_to_normals, azimuth_elevation_to_primary_direction, make_distance_maps,  # noqa: F401
                  sample_cone_directions)
from .field import HelioField  # noqa: F401
from .graphs import GraphedStep  # noqa: F401
from .layers import CenterOfMass2D  # noqa: F401

__all__ = ["HelioField", "HelioEnv", "GraphedStep", "CenterOfMass2D", "HelioLibError", "SPLAT_AUTO", "SPLAT_SIMT", "SPLAT_TC",
           "azimuth_elevation_to_primary_direction", "sample_cone_directions", "make_distance_maps", "angles_to_normals"]
