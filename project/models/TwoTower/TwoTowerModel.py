"""Drop-in for the reference's project/models/TwoTower/TwoTowerModel.py."""
from recommendsystemproject_b200.modules import TwoTowerModel  # noqa: F401
