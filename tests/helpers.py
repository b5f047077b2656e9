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


# ---------------------------------------------------------------- full-size goldens (tests/golden/make_golden_full.py)
def full_case(name):
    """(npz, cfg, maps, [batch0, batch1], corpus, seed) of a full-size golden case; batches are regenerated from
    synth.py with the seeds the generator used."""
    from recommendsystemproject_b200 import synth
    npz, cfg = load_golden(name)
    seed = int(npz["seed"])
    if name == "c2_full":
        return (npz, cfg, synth.MAPS_C2, [synth.make_batch_c2(512, 20, 10, seed=21), synth.make_batch_c2(512, 20, 10, seed=22)],
                synth.make_corpus_c2(3416, seed=7), seed)
    if name == "c1_full":
        return (npz, cfg, synth.MAPS_C1, [synth.make_batch_c1(1024, 50, seed=41), synth.make_batch_c1(1024, 50, seed=42)],
                {"sparse": torch.arange(1, 3707).unsqueeze(1)}, seed)
    raise KeyError(name)


def check_digest(t, d, what, rtol=2e-4, atol_head=2e-5, atol_sum=None, atol_elem=0.0):
    """A tensor against the (sum, sum|.|, first 16 values[, full]) digest stored in a full-size golden.
    atol_elem: absolute slack per element (tensors that are analytically zero hold fp32 rounding noise only)."""
    t = t.detach().double().cpu().reshape(-1)
    assert t.numel() == int(d["numel"]), what
    scale = float(d["abs"]) + 1e-12
    floor = atol_elem * t.numel() + 1e-9
    assert abs(float(t.abs().sum()) - float(d["abs"])) <= rtol * scale + floor, (what, float(t.abs().sum()), float(d["abs"]))
    tol_sum = atol_sum if atol_sum is not None else rtol * scale + floor
    assert abs(float(t.sum()) - float(d["sum"])) <= tol_sum, (what, float(t.sum()), float(d["sum"]))
    assert torch.allclose(t[:16], d["head"].double(), atol=atol_head, rtol=1e-3), (what, t[:16], d["head"])
