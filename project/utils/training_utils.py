"""Drop-in for the reference's project/utils/training_utils.py."""
from recommendsystemproject_b200.training import (  # noqa: F401
    _log_embedding_stats, build_user_history, extract_item_id, to_device, train_one_epoch, validate)
