"""Drop-in for the reference's project/models/TwoTower/GenericTower.py."""
from recommendsystemproject_b200.modules import GenericTower  # noqa: F401
