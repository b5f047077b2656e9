"""Flatten / unflatten nested batch dicts + state dicts into one .npz file
(used by tests/golden/make_golden.py and the parity tests)."""
import json
import numpy as np
import torch


def flatten(obj, prefix, out):
    if isinstance(obj, torch.Tensor):
        out[prefix] = obj.detach().cpu().numpy()
    elif isinstance(obj, np.ndarray):
        out[prefix] = obj
    elif isinstance(obj, dict):
        for k, v in obj.items():
            flatten(v, f"{prefix}/{k}", out)
    elif isinstance(obj, (list, tuple)):
        for n, v in enumerate(obj):
            flatten(v, f"{prefix}/#{n}", out)
    elif obj is None:
        pass
    else:
        out[prefix] = np.asarray(obj)


def unflatten(npz, prefix):
    """Rebuild the nested object stored under ``prefix``."""
    root = {}
    plen = len(prefix) + 1
    found = False
    for key in npz.files:
        if key == prefix:
            return torch.from_numpy(np.array(npz[key]))
        if not key.startswith(prefix + "/"):
            continue
        found = True
        parts = key[plen:].split("/")
        node = root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = torch.from_numpy(np.array(npz[key]))
    if not found:
        return None

    def listify(node):
        if not isinstance(node, dict):
            return node
        node = {k: listify(v) for k, v in node.items()}
        if node and all(k.startswith("#") for k in node):
            return [node[f"#{n}"] for n in range(len(node))]
        return node

    return listify(root)


def save_case(path, cfg, **sections):
    out = {}
    for name, obj in sections.items():
        flatten(obj, name, out)
    out["__cfg__"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)


def load_case(path):
    npz = np.load(path)
    cfg = json.loads(bytes(npz["__cfg__"]).decode())
    return npz, cfg
