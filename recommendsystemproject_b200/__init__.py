"""B200-native two-tower DSSM hot path (sm_100a CUDA kernels behind the
reference's YAML-driven model classes).  No CPU fallback."""
from . import _lib  # noqa: F401
from .modules import GenericTower, MLP_Tower, SequenceEncoder, SequenceFeatureProcessor, TwoTowerModel
from .optim import FusedTwoTowerOptimizer, GraphedTrainStep
from .batching import GpuBatchBuilder
from . import torch_ops  # noqa: F401  (registers torch.ops.tt_b200.*)

__all__ = ["GenericTower", "MLP_Tower", "SequenceEncoder", "SequenceFeatureProcessor", "TwoTowerModel",
           "FusedTwoTowerOptimizer", "GraphedTrainStep", "GpuBatchBuilder"]
