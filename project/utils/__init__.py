"""Import-path shim: the reference's `project.utils.*` module paths.

training_utils / config_utils / SequenceFeatureProcessor are served by recommendsystemproject_b200 (files in this
directory).  The host-side pandas loaders (DataLoader.py, CombineTwoTower.py: SURVEY.md 2.1 rows 8-9, out of scope to
rewrite, "reuse as-is") are NOT copied here: when a reference checkout is reachable its own project/utils directory is
appended to this package's search path, so `from project.utils.DataLoader import create_loader` in the unmodified
train_twotower.py resolves to the reference's file while everything on the hot path resolves to this repo.
Search order: $TT_REFERENCE_ROOT, <repo>/baseline/_ref, /root/reference."""
import os as _os

_repo = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
for _root in (_os.environ.get("TT_REFERENCE_ROOT"), _os.path.join(_repo, "baseline", "_ref"), "/root/reference"):
    if _root and _os.path.isdir(_os.path.join(_root, "project", "utils")):
        __path__.append(_os.path.join(_root, "project", "utils"))
        break
