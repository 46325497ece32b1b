"""Drop-in for the reference's layers/center_of_mass.py (imported as `from layers.center_of_mass import
CenterOfMass2D`, train_with_env_com_trunc_advantage_ttt.py:22): re-exports the sm_100a implementation."""
from doodle_b200.layers import CenterOfMass2D  # noqa: F401
