"""Drop-in for the reference's project/utils/SequenceFeatureProcessor.py."""
from recommendsystemproject_b200.modules import SequenceFeatureProcessor  # noqa: F401
