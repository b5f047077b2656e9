"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (imported from /root/reference -- present only in the authoring
container, never on the GPU box) on seeded synthetic inputs.

    python tests/golden/make_golden.py

Writes tests/golden/{pool_small,seq_small}.npz and kat.npz.  The parity tests
read only those files.  Nothing from the reference is copied: only inputs and
numeric outputs are stored.
"""
import copy
import os
import sys

import numpy as np
import torch

REF = os.environ.get("TT_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from golden_io import save_case  # noqa: E402


def import_reference():
    # make sure `project.*` resolves to the reference, not to this repo's shim
    for name in [m for m in sys.modules if m == "project" or m.startswith("project.")]:
        del sys.modules[name]
    sys.path.insert(0, REF)
    from project.models.TwoTower.GenericTower import GenericTower
    from project.models.TwoTower.TwoTowerModel import TwoTowerModel
    sys.path.pop(0)
    return GenericTower, TwoTowerModel


CFG_POOL = {  # C1-shaped (BASELINE.json configs[0]), shrunk; mean + sum + max pooling
    "two_tower": {
        "user_tower": {
            "mlp_hidden_dim": [32, 16], "output_dims": 16, "dropout": 0.0, "embedding_dim": 16,
            "sparse_features": [
                {"name": "user_id_enc", "vocab_size": 50, "embedding_dim": 16},
                {"name": "hist_movie_ids", "vocab_size": 40, "embedding_dim": 16, "padding_idx": 0,
                 "pooling": "mean"},
                {"name": "hist_sum", "vocab_size": 30, "embedding_dim": 8, "padding_idx": 0, "pooling": "sum"},
                {"name": "hist_max", "vocab_size": 30, "embedding_dim": 4, "padding_idx": 0, "pooling": "max"},
            ],
        },
        "item_tower": {
            "mlp_hidden_dim": [32, 16], "output_dims": 16, "dropout": 0.0, "embedding_dim": 16,
            "sparse_features": [{"name": "movie_id_enc", "vocab_size": 40, "embedding_dim": 16}],
        },
    },
    "train": {"batch_size": 24, "learning_rate": 5e-3, "temperature": 0.15},
}

CFG_SEQ = {  # shipped config.yaml shape (BASELINE.json configs[1]), shrunk vocab/dims
    "two_tower": {
        "user_tower": {
            "mlp_hidden_dim": [32, 16], "output_dims": 16, "dropout": 0.0, "embedding_dim": 16,
            "transformer_parameters": {"max_seq_len": 6, "n_head": 4, "n_layers": 2, "FFN_dim": 32,
                                       "dropout": 0.0},
            "sparse_features": [{"name": "user_id_enc", "vocab_size": 60, "embedding_dim": 16}],
            "dense_features": [{"name": "user_activity_log", "dim": 1, "embedding_dim": 8}],
            "sequence_features": [
                {"name": "hist_movie_ids", "vocab_size": 50, "embedding_dim": 8, "padding_idx": 0},
                {"name": "hist_genre_ids", "vocab_size": 12, "embedding_dim": 4, "padding_idx": 0,
                 "pooling": "mean"},
            ],
        },
        "item_tower": {
            "mlp_hidden_dim": [32, 16], "output_dims": 16, "dropout": 0.0, "embedding_dim": 16,
            "sparse_features": [
                {"name": "movie_id_enc", "vocab_size": 50, "embedding_dim": 8},
                {"name": "genre_ids", "vocab_size": 12, "embedding_dim": 4, "padding_idx": 0,
                 "pooling": "mean"},
                {"name": "release_year_enc", "vocab_size": 20, "embedding_dim": 4},
            ],
        },
    },
    "train": {"batch_size": 16, "learning_rate": 5e-3, "temperature": 0.15},
}


def right_padded(gen, B, L, vocab, min_len=0):
    ids = torch.randint(1, vocab, (B, L), generator=gen)
    lens = torch.randint(min_len, L + 1, (B,), generator=gen)
    ids[torch.arange(L)[None, :] >= lens[:, None]] = 0
    return ids


def batch_pool(gen, B, L=7):
    user = {"sparse": torch.randint(1, 50, (B, 1), generator=gen),
            "sequence": {"hist_movie_ids": right_padded(gen, B, L, 40, 1),
                         "hist_sum": right_padded(gen, B, 5, 30, 0),
                         "hist_max": right_padded(gen, B, 4, 30, 1)}}
    item = {"sparse": torch.randint(1, 12, (B, 1), generator=gen)}  # small range -> id collisions
    maps = ({"sparse": {"user_id_enc": 0}, "dense": {}, "sequence": {}},
            {"sparse": {"movie_id_enc": 0}, "dense": {}, "sequence": {}})
    return {"user_tower": user, "item_tower": item}, maps


def item_slab(gen, B, vocab=50):
    return {"sparse": torch.stack([torch.randint(1, vocab, (B,), generator=gen),
                                   torch.randint(1, 20, (B,), generator=gen)], dim=1),
            "sequence": {"genre_ids": right_padded(gen, B, 3, 12, 1)}}


def batch_seq(gen, B, L=6, n_neg=3):
    hist = right_padded(gen, B, L, 50, 1)
    hist[0] = 0  # an all-padding row (SequenceEncoder.py:43-46)
    hist[1] = torch.randint(1, 50, (L,), generator=gen)  # a full row
    genres = torch.randint(1, 12, (B, L, 3), generator=gen)
    genres[hist == 0] = 0
    genres[:, :, 2][torch.rand(B, L, generator=gen) < 0.5] = 0
    user = {"sparse": torch.randint(1, 60, (B, 1), generator=gen),
            "dense": torch.rand(B, 1, generator=gen) * 3.0,
            "sequence": {"hist_movie_ids": hist, "hist_genre_ids": genres}}
    item = item_slab(gen, B, vocab=10)  # collisions among positives
    maps = ({"sparse": {"user_id_enc": 0}, "dense": {"user_activity_log": 0},
             "sequence": {"hist_movie_ids": "hist_movie_ids", "hist_genre_ids": "hist_genre_ids"}},
            {"sparse": {"movie_id_enc": 0, "release_year_enc": 1}, "dense": {},
             "sequence": {"genre_ids": "genre_ids"}})
    batch = {"user_tower": user, "item_tower": item,
             "hard_negatives": [item_slab(gen, B) for _ in range(n_neg)]}
    return batch, maps


def corpus_pool(gen):
    return {"sparse": torch.arange(1, 40).unsqueeze(1)}  # unique ids (training_utils.py:173-175)


def corpus_seq(gen):
    slab = item_slab(gen, 49)
    slab["sparse"][:, 0] = torch.arange(1, 50)
    return slab


def run_case(name, cfg, make_batch, make_corpus, seed):
    GenericTower, TwoTowerModel = import_reference()
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1000)
    B = cfg["train"]["batch_size"]
    T = cfg["train"]["temperature"]
    lr = cfg["train"]["learning_rate"]
    batch, (umap, imap) = make_batch(gen, B)
    batch2, _ = make_batch(gen, B)
    model = TwoTowerModel(GenericTower(cfg, "user_tower"), GenericTower(cfg, "item_tower"), umap, imap)
    state_init = copy.deepcopy(model.state_dict())  # what torch.manual_seed(seed) + construction gives
    # give BN affine params / biases non-trivial values so they are exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") or ("bn" in n or ".mlp.1." in n or ".mlp.5." in n) and n.endswith("weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    state0 = copy.deepcopy(model.state_dict())

    # eval-mode forward (running stats) -- retrieval encodes the corpus this way
    model.eval()
    with torch.no_grad():
        ue, ie, hne = model(batch)
    model.train()

    opt = torch.optim.Adam(model.parameters(), lr=lr)
    steps = []
    for b in (batch, batch2):
        opt.zero_grad()
        u, i, hn = model(b)
        ids = b["item_tower"]["sparse"][:, 0]
        loss = model.compute_loss(u, i, hard_neg_emb=hn, item_ids=ids, temperature=T)
        loss.backward()
        grads = {n: (torch.zeros_like(p) if p.grad is None else p.grad.clone())
                 for n, p in model.named_parameters()}
        total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        steps.append({"u": u, "i": i, "hn": hn, "loss": loss, "grads": grads, "total_norm": total_norm,
                      "state_after": copy.deepcopy(model.state_dict())})

    # retrieval over a small catalog with unique item ids
    model.eval()
    with torch.no_grad():
        corpus_in = make_corpus(gen)
        corpus = model.get_item_embeddings(corpus_in)
        uq, _, _ = model(batch)
        scores = uq @ corpus.t()
        topv, topi = torch.topk(scores, k=5, dim=1)
    save_case(os.path.join(HERE, name + ".npz"), cfg,
              state_init=state_init, seed=torch.tensor(seed), state0=state0, batch=batch, batch2=batch2,
              maps={"user": _enc_map(umap), "item": _enc_map(imap)},
              eval0={"u": ue, "i": ie, "hn": hne},
              step0=steps[0], step1=steps[1],
              retrieval={"corpus_in": corpus_in, "corpus": corpus, "queries": uq, "scores": scores, "topk_vals": topv, "topk_idx": topi})
    print(name, "loss0", float(steps[0]["loss"]), "loss1", float(steps[1]["loss"]),
          "norm0", float(steps[0]["total_norm"]))


def _enc_map(m):
    # sparse/dense column maps only (sequence map is name->name)
    return {"sparse": {k: torch.tensor(v) for k, v in m["sparse"].items()},
            "dense": {k: torch.tensor(v) for k, v in m["dense"].items()}}


def run_kats():
    """KATs 1-6 of SURVEY.md section 4, evaluated on the reference."""
    GenericTower, TwoTowerModel = import_reference()
    m = TwoTowerModel(None, None)
    eye = torch.eye(2)
    out = {
        "kat1": m.compute_loss(eye, eye, temperature=0.5),
        "kat2": m.compute_loss(eye, eye, item_ids=torch.tensor([5, 5]), temperature=0.5),
        "kat3": m.compute_loss(eye, eye, hard_neg_emb=eye.unsqueeze(1), temperature=0.5),
    }
    # KAT4: autograd grads of the loss at B=8, N=3, D=16
    g = torch.Generator().manual_seed(4)
    u = torch.nn.functional.normalize(torch.randn(8, 16, generator=g), dim=1).requires_grad_(True)
    i = torch.nn.functional.normalize(torch.randn(8, 16, generator=g), dim=1).requires_grad_(True)
    hn = torch.nn.functional.normalize(torch.randn(8, 3, 16, generator=g), dim=2).requires_grad_(True)
    ids = torch.tensor([3, 1, 3, 2, 2, 7, 3, 9])
    loss = m.compute_loss(u, i, item_ids=ids, hard_neg_emb=hn, temperature=0.15)
    loss.backward()
    out["kat4"] = {"u": u.detach(), "i": i.detach(), "hn": hn.detach(), "ids": ids, "loss": loss.detach(),
                   "du": u.grad, "di": i.grad, "dhn": hn.grad}
    # larger loss-only case with a shared pool expressed as per-row negatives
    g = torch.Generator().manual_seed(5)
    B, H, D = 96, 40, 32
    u = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).requires_grad_(True)
    i = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).requires_grad_(True)
    pool = torch.nn.functional.normalize(torch.randn(H, D, generator=g), dim=1).requires_grad_(True)
    ids = torch.randint(1, 30, (B,), generator=g)
    loss = m.compute_loss(u, i, item_ids=ids, hard_neg_emb=pool.unsqueeze(0).expand(B, H, D), temperature=0.05)
    loss.backward()
    out["pool"] = {"u": u.detach(), "i": i.detach(), "pool": pool.detach(), "ids": ids, "loss": loss.detach(),
                   "du": u.grad, "di": i.grad, "dpool": pool.grad}
    # KAT6: mean pooling includes pad positions and W[0] != 0
    cfg = copy.deepcopy(CFG_POOL)
    torch.manual_seed(6)
    t = GenericTower(cfg, "user_tower")
    w = t.embeddings["hist_movie_ids"].weight.detach()
    idsq = torch.tensor([[3, 5, 0, 0]])
    pooled = torch.mean(t.embeddings["hist_movie_ids"](idsq), dim=1)
    out["kat6"] = {"w": w, "ids": idsq, "pooled": pooled.detach()}
    save_case(os.path.join(HERE, "kat.npz"), {}, **out)
    print("kat1", float(out["kat1"]), "kat2", float(out["kat2"]), "kat3", float(out["kat3"]))


if __name__ == "__main__":
    torch.set_num_threads(1)
    run_case("pool_small", CFG_POOL, batch_pool, corpus_pool, seed=11)
    run_case("seq_small", CFG_SEQ, batch_seq, corpus_seq, seed=12)
    run_kats()
