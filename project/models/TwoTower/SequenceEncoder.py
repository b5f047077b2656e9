"""Drop-in for the reference's project/models/TwoTower/SequenceEncoder.py."""
from recommendsystemproject_b200.modules import SequenceEncoder  # noqa: F401
