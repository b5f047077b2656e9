"""Shared helpers for the parity tests."""
import os

import torch

from golden_io import load_case, unflatten

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    npz, cfg = load_case(os.path.join(GOLDEN, name + ".npz"))
    return npz, cfg


def get_maps(npz):
    m = unflatten(npz, "maps")
    out = {}
    for side in ("user", "item"):
        s = m.get(side, {}) if m else {}
        out[side] = {"sparse": {k: int(v) for k, v in (s.get("sparse") or {}).items()},
                     "dense": {k: int(v) for k, v in (s.get("dense") or {}).items()},
                     "sequence": {}}
    return out["user"], out["item"]


def clone_state(state):
    return {k: v.clone() for k, v in state.items()}


def to_device(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, list):
        return [to_device(v, device) for v in obj]
    return obj
