"""Synthetic configs + seeded batch generators shaped like BASELINE.json
configs C1..C5.  Batches follow the reference's
collate contract (DataLoader.py:250-288, CombineTwoTower.py:62-92):
  {'user_tower': {'sparse': int64[B,n], 'dense': f32[B,n], 'sequence': {name: int64[B,L] | [B,L,Tags]}},
   'item_tower': {...}, 'hard_negatives': [item dict x N]}
Generated on the CPU (host side of the pipeline); callers move them to the GPU.
"""
from __future__ import annotations

import copy

import torch


def config_c1(dropout: float = 0.0):
    """C1: MovieLens-1M-shaped, mean-pooled 50-item history, dim 64, B=1024."""
    return {
        "two_tower": {
            "user_tower": {
                "mlp_hidden_dim": [256, 128], "output_dims": 64, "dropout": dropout, "embedding_dim": 64,
                "sparse_features": [
                    {"name": "user_id_enc", "vocab_size": 6041, "embedding_dim": 64},
                    {"name": "hist_movie_ids", "vocab_size": 3707, "embedding_dim": 64, "padding_idx": 0,
                     "pooling": "mean"},
                ],
            },
            "item_tower": {
                "mlp_hidden_dim": [256, 128], "output_dims": 64, "dropout": dropout, "embedding_dim": 64,
                "sparse_features": [{"name": "movie_id_enc", "vocab_size": 3707, "embedding_dim": 64}],
            },
        },
        "train": {"batch_size": 1024, "learning_rate": 5e-4, "temperature": 0.15},
    }


def config_c2(dropout_scale: float = 1.0):
    """C2: the reference's shipped config.yaml (Transformer 2L/4H/d64/FFN256, L=20, 3 genre tags)."""
    d = dropout_scale
    return {
        "two_tower": {
            "user_tower": {
                "mlp_hidden_dim": [256, 128], "output_dims": 128, "dropout": 0.3 * d, "embedding_dim": 64,
                "max_seq_len": 20,
                "transformer_parameters": {"max_seq_len": 20, "n_head": 4, "n_layers": 2, "FFN_dim": 256,
                                           "dropout": 0.15 * d},
                "sparse_features": [{"name": "user_id_enc", "vocab_size": 6060, "embedding_dim": 64}],
                "dense_features": [{"name": "user_activity_log", "dim": 1, "embedding_dim": 8}],
                "sequence_features": [
                    {"name": "hist_movie_ids", "vocab_size": 3500, "embedding_dim": 32, "padding_idx": 0},
                    {"name": "hist_genre_ids", "vocab_size": 30, "embedding_dim": 8, "padding_idx": 0,
                     "pooling": "mean"},
                ],
            },
            "item_tower": {
                "mlp_hidden_dim": [256, 128], "output_dims": 128, "dropout": 0.1 * d, "embedding_dim": 64,
                "transformer_parameters": {"max_seq_len": 20, "n_head": 4, "n_layers": 2, "FFN_dim": 256,
                                           "dropout": 0.0},
                "sparse_features": [
                    {"name": "movie_id_enc", "vocab_size": 3500, "embedding_dim": 32},
                    {"name": "genre_ids", "vocab_size": 30, "embedding_dim": 8, "padding_idx": 0, "pooling": "mean"},
                    {"name": "release_year_enc", "vocab_size": 152, "embedding_dim": 8},
                ],
            },
        },
        "train": {"batch_size": 512, "learning_rate": 5e-4, "temperature": 0.15},
    }


def config_c3(v_user=100_000_001, v_item=10_000_001, dim=128, dropout=0.1, shard=True, world=None, rank=0,
              capacity_factor=1.25):
    """C3: 8 sparse features -- user {user_id 100M+1, u_cat1 1e5, u_cat2 1e3, u_cat3 32} + pooled
    hist_item_ids (10M+1, L=200 ragged, mean); item {item_id 10M+1, i_cat 1e4, i_year 152}; all D_f = 128;
    MLP [256, 128] -> 128.  Tables of 1 MB and more are row-sharded (`row_sharded: true`), the rest replicated."""
    def feat(name, vocab, **kw):
        f = {"name": name, "vocab_size": vocab, "embedding_dim": dim}
        f.update(kw)
        if shard and vocab * dim * 4 >= (1 << 20):
            f["row_sharded"] = True
        return f
    cfg = {
        "two_tower": {
            "user_tower": {
                "mlp_hidden_dim": [256, 128], "output_dims": 128, "dropout": dropout, "embedding_dim": dim,
                "sparse_features": [feat("user_id", v_user), feat("u_cat1", 100_000), feat("u_cat2", 1000), feat("u_cat3", 32),
                                    feat("hist_item_ids", v_item, padding_idx=0, pooling="mean")],
            },
            "item_tower": {
                "mlp_hidden_dim": [256, 128], "output_dims": 128, "dropout": dropout, "embedding_dim": dim,
                "sparse_features": [feat("item_id", v_item), feat("i_cat", 10_000), feat("i_year", 152)],
            },
        },
        "train": {"batch_size": 65536, "learning_rate": 5e-4, "temperature": 0.05},
    }
    if world is not None:
        cfg["two_tower"]["sharding"] = {"rank": rank, "world": world, "capacity_factor": capacity_factor}
    return cfg


MAPS_C3 = ({"sparse": {"user_id": 0, "u_cat1": 1, "u_cat2": 2, "u_cat3": 3}, "dense": {}, "sequence": {}},
           {"sparse": {"item_id": 0, "i_cat": 1, "i_year": 2}, "dense": {}, "sequence": {}})


def zipf_approx(gen, shape, vocab, s=1.05, device=None):
    """ids in [1, vocab) with P(k) ~ k^-s for LARGE vocabularies (inverse CDF of the continuous law, then a
    multiplicative hash so that popular ids are spread over the table instead of sitting in rows 1, 2, 3, ...)."""
    u = torch.rand(shape, generator=gen, dtype=torch.float64, device=device)
    n = float(vocab - 1)
    x = ((n ** (1.0 - s) - 1.0) * u + 1.0) ** (1.0 / (1.0 - s))
    k = x.floor().clamp_(1, vocab - 1).long() - 1
    return (k * 2654435761 % (vocab - 1)) + 1


def make_batch_c3(B=65536, L=200, v_user=100_000_001, v_item=10_000_001, seed=3, zipf=False, unique_items=False):
    """One C3 batch (collate contract).  zipf=False: uniform ids (every looked-up row distinct: the HBM worst case);
    zipf=True: Zipf(1.05) item ids as SURVEY 8d words it.  unique_items: distinct positive item ids (no in-batch
    false negatives)."""
    gen = torch.Generator().manual_seed(seed)
    draw = (lambda shape: zipf_approx(gen, shape, v_item)) if zipf else (lambda shape: torch.randint(1, v_item, shape, generator=gen))
    hist = draw((B, L))
    lens = torch.randint(1, L + 1, (B,), generator=gen)
    hist[torch.arange(L)[None, :] >= lens[:, None]] = 0
    user = {"sparse": torch.stack([torch.randint(0, v_user, (B,), generator=gen), torch.randint(0, 100_000, (B,), generator=gen),
                                   torch.randint(0, 1000, (B,), generator=gen), torch.randint(0, 32, (B,), generator=gen)], dim=1),
            "sequence": {"hist_item_ids": hist}}
    item_ids = (torch.randperm(v_item - 1, generator=gen)[:B] + 1) if unique_items else draw((B,))
    item = {"sparse": torch.stack([item_ids, torch.randint(0, 10_000, (B,), generator=gen),
                                   torch.randint(0, 152, (B,), generator=gen)], dim=1)}
    return {"user_tower": user, "item_tower": item}


MAPS_C1 = ({"sparse": {"user_id_enc": 0}, "dense": {}, "sequence": {}},
           {"sparse": {"movie_id_enc": 0}, "dense": {}, "sequence": {}})
MAPS_C2 = ({"sparse": {"user_id_enc": 0}, "dense": {"user_activity_log": 0},
            "sequence": {"hist_movie_ids": "hist_movie_ids", "hist_genre_ids": "hist_genre_ids"}},
           {"sparse": {"movie_id_enc": 0, "release_year_enc": 1}, "dense": {}, "sequence": {"genre_ids": "genre_ids"}})


def right_padded(gen, B, L, vocab, min_len=1):
    ids = torch.randint(1, vocab, (B, L), generator=gen)
    lens = torch.randint(min_len, L + 1, (B,), generator=gen)
    ids[torch.arange(L)[None, :] >= lens[:, None]] = 0
    return ids


def zipf_ids(gen, n, vocab, s=1.0):
    """ids in [1, vocab) with P(k) ~ k^-s."""
    w = torch.arange(1, vocab, dtype=torch.float64).pow(-s)
    return torch.multinomial(w, n, replacement=True, generator=gen) + 1


def make_batch_c1(B=1024, L=50, seed=1):
    gen = torch.Generator().manual_seed(seed)
    user = {"sparse": torch.randint(1, 6041, (B, 1), generator=gen),
            "sequence": {"hist_movie_ids": right_padded(gen, B, L, 3707)}}
    item = {"sparse": zipf_ids(gen, B, 3707).unsqueeze(1)}
    return {"user_tower": user, "item_tower": item}


def _item_slab_c2(gen, B):
    return {"sparse": torch.stack([torch.randint(1, 3500, (B,), generator=gen),
                                   torch.randint(1, 152, (B,), generator=gen)], dim=1),
            "sequence": {"genre_ids": right_padded(gen, B, 3, 30)}}


def make_batch_c2(B=512, L=20, n_neg=10, seed=2):
    gen = torch.Generator().manual_seed(seed)
    hist = right_padded(gen, B, L, 3500)
    genres = torch.randint(1, 30, (B, L, 3), generator=gen)
    genres[hist == 0] = 0
    user = {"sparse": torch.randint(1, 6060, (B, 1), generator=gen),
            "dense": torch.rand(B, 1, generator=gen) * 6.0,
            "sequence": {"hist_movie_ids": hist, "hist_genre_ids": genres}}
    batch = {"user_tower": user, "item_tower": _item_slab_c2(gen, B)}
    if n_neg > 0:
        batch["hard_negatives"] = [_item_slab_c2(gen, B) for _ in range(n_neg)]
    return batch


def make_corpus_c2(n_items=3416, seed=7):
    gen = torch.Generator().manual_seed(seed)
    slab = _item_slab_c2(gen, n_items)
    slab["sparse"][:, 0] = torch.arange(1, n_items + 1)
    return slab


def normalized(gen, n, d):
    return torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=1)


def clone_cfg(cfg):
    return copy.deepcopy(cfg)
