"""Drop-in for the reference's project/models/TwoTower/Tower.py."""
from recommendsystemproject_b200.modules import MLP_Tower  # noqa: F401
