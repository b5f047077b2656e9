"""Golden fixture for the GPU batch builder: batches produced by the UNMODIFIED reference loaders
(CombinedTwoTowerDataLoader -> RecommendationDataset.__getitem__ -> collate_fn, CombineTwoTower.py:62-92 and
DataLoader.py:226-288) on a small ML-1M-shaped DataFrame, together with the DataFrame's columns as tensors.

    python tests/golden/make_golden_collate.py      (authoring container only: imports /root/reference)
"""
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from golden_io import save_case  # noqa: E402
from make_golden import REF  # noqa: E402


def main():
    import pandas as pd
    rng = np.random.default_rng(5)
    n, n_items = 43, 60                                   # 43 rows, batches of 8: the last batch is partial
    users = rng.integers(1, 200, size=n)
    items = rng.integers(1, n_items + 1, size=n)
    hist = np.zeros((n, 20), dtype=np.int64)
    for r in range(n):
        ln = rng.integers(0, 21)                           # includes an empty history
        hist[r, :ln] = rng.integers(1, n_items + 1, size=ln)
    genres_of = rng.integers(0, 19, size=(n_items + 1, 3))
    genres_of[0] = 0
    hg = genres_of[hist]
    year_of = rng.integers(1, 100, size=n_items + 1)
    act = (rng.random(n) * 5).astype(np.float32)
    df = pd.DataFrame({"user_id_enc": users, "user_activity_log": act, "hist_movie_ids": [x.tolist() for x in hist],
                       "hist_genre_ids": [x.tolist() for x in hg], "movie_id_enc": items,
                       "genre_ids": [x.tolist() for x in genres_of[items]], "release_year_enc": year_of[items]})
    # the reference, with its `project` namespace package first on the path and this repo's shim out of the way
    root = os.path.dirname(os.path.dirname(HERE))
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != root]
    for m in [m for m in sys.modules if m == "project" or m.startswith("project.")]:
        del sys.modules[m]
    from project.utils.CombineTwoTower import CombinedTwoTowerDataLoader
    with tempfile.TemporaryDirectory() as tmp:
        pkl = os.path.join(tmp, "train.pkl")
        df.to_pickle(pkl)
        cfg_path = os.path.join(REF, "config.yaml")        # the shipped YAML decides which columns go where
        loader = CombinedTwoTowerDataLoader(config_path=cfg_path, pickle_path=pkl, batch_size=8, shuffle=False, num_workers=0)
        batches = [b for b in loader]
        maps = loader.get_feature_mappings()
    assert "CombineTwoTower" in sys.modules["project.utils.CombineTwoTower"].__file__ and REF in sys.modules["project.utils.CombineTwoTower"].__file__
    cols = {"user": {"sparse": torch.from_numpy(users).long().unsqueeze(1), "dense": torch.from_numpy(act).unsqueeze(1),
                     "sequence": {"hist_movie_ids": torch.from_numpy(hist), "hist_genre_ids": torch.from_numpy(hg)}},
            "item": {"sparse": torch.stack([torch.from_numpy(items).long(), torch.from_numpy(year_of[items]).long()], dim=1),
                     "sequence": {"genre_ids": torch.from_numpy(genres_of[items])}}}
    enc = lambda m: {k: {kk: (torch.tensor(vv) if isinstance(vv, int) else torch.tensor(-1)) for kk, vv in v.items()} for k, v in m.items()}
    save_case(os.path.join(HERE, "collate_small.npz"), {"batch_size": 8, "n": n}, columns=cols, batches=batches,
              maps={"user": enc(maps["user"]), "item": enc(maps["item"])})
    print("batches", len(batches), {k: tuple(v.shape) for k, v in batches[0]["user_tower"]["sequence"].items()}, maps)


if __name__ == "__main__":
    main()
