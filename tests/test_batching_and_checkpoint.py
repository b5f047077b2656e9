"""GPU batch builder and checkpoint resume (N4).  The builder is pure tensor indexing and is checked on
the CPU against a per-sample restatement of the reference's collate (DataLoader.py:250-288: np.stack of the sparse /
dense rows and of every sequence feature; CombineTwoTower.py:62-92: the user/item combination) plus the hard-negative
slabs the reference left as a TODO."""
import numpy as np
import pytest
import torch

from recommendsystemproject_b200 import synth
from recommendsystemproject_b200.batching import GpuBatchBuilder


def _columns(n, n_items, seed):
    gen = torch.Generator().manual_seed(seed)
    user = {"sparse": torch.randint(1, 6060, (n, 1), generator=gen), "dense": torch.rand(n, 1, generator=gen),
            "sequence": {"hist_movie_ids": synth.right_padded(gen, n, 20, 3500),
                         "hist_genre_ids": torch.randint(0, 30, (n, 20, 3), generator=gen)}}
    catalog = synth.make_corpus_c2(n_items, seed=seed + 1)
    catalog_ids = catalog["sparse"][:, 0].clone()
    pos_rows = torch.randint(0, n_items, (n,), generator=gen)
    item = {"sparse": catalog["sparse"][pos_rows], "sequence": {"genre_ids": catalog["sequence"]["genre_ids"][pos_rows]}}
    neg = catalog_ids[torch.randint(0, n_items, (n, 4), generator=gen)]
    neg[torch.rand(n, 4, generator=gen) < 0.2] = 0            # "no negative" slots
    return user, item, catalog, catalog_ids, neg


def _collate_reference(user, item, catalog, catalog_ids, neg, indices):
    """per-sample python loop, like the reference's __getitem__ + collate_fn"""
    def one(group, i):
        s = {}
        if "sparse" in group:
            s["sparse"] = group["sparse"][i].numpy()
        if "dense" in group:
            s["dense"] = group["dense"][i].numpy()
        if "sequence" in group:
            s["sequence"] = {k: v[i].numpy() for k, v in group["sequence"].items()}
        return s

    def collate(samples):
        out = {}
        if "sparse" in samples[0]:
            out["sparse"] = torch.from_numpy(np.stack([s["sparse"] for s in samples])).long()
        if "dense" in samples[0]:
            out["dense"] = torch.from_numpy(np.stack([s["dense"] for s in samples])).float()
        if "sequence" in samples[0]:
            out["sequence"] = {k: torch.from_numpy(np.stack([s["sequence"][k] for s in samples])).long()
                               for k in samples[0]["sequence"]}
        return out
    row_of = {int(cid): r for r, cid in enumerate(catalog_ids.tolist())}
    batch = {"user_tower": collate([one(user, i) for i in indices]), "item_tower": collate([one(item, i) for i in indices])}
    slabs = []
    for n in range(neg.shape[1]):
        samples = []
        for i in indices:
            nid = int(neg[i, n])
            if nid == 0:
                samples.append({"sparse": np.zeros_like(catalog["sparse"][0].numpy()),
                                "sequence": {k: np.zeros_like(v[0].numpy()) for k, v in catalog["sequence"].items()}})
            else:
                samples.append(one(catalog, row_of[nid]))
        slabs.append(collate(samples))
    batch["hard_negatives"] = slabs
    return batch


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return a.dtype == b.dtype and torch.equal(a.cpu(), b.cpu())
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))


def test_gpu_batch_builder_matches_per_sample_collate():
    user, item, catalog, catalog_ids, neg = _columns(200, 57, seed=3)
    builder = GpuBatchBuilder(user, item, catalog, catalog_ids, neg, device="cpu")
    idx = [5, 199, 0, 42, 42, 17]
    got = builder.batch(torch.tensor(idx))
    ref = _collate_reference(user, item, catalog, catalog_ids, neg, idx)
    assert _same(got, ref)
    batches = list(builder.epoch(64, shuffle=True, generator=torch.Generator().manual_seed(1)))
    assert sum(b["user_tower"]["sparse"].shape[0] for b in batches) == 200
    assert len(batches[0]["hard_negatives"]) == 4


@pytest.mark.gpu
def test_builder_feeds_the_model_and_checkpoint_resumes_across_optimizers(tmp_path):
    """A batch from the builder drives a training step; a checkpoint written in the reference's format by the fused
    optimizer resumes into torch.optim.Adam (and back) with identical next-step parameters."""
    import recommendsystemproject_b200 as tt
    from recommendsystemproject_b200 import training
    dev = "cuda"
    user, item, catalog, catalog_ids, neg = _columns(300, 80, seed=9)
    builder = GpuBatchBuilder(user, item, catalog, catalog_ids, neg, device=dev)
    cfg = synth.config_c2(dropout_scale=0.0)

    def fresh():
        torch.manual_seed(0)
        return tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), *synth.MAPS_C2).to(dev).train()

    def step(model, opt, batch, fused):
        opt.zero_grad()
        u, i, hn = model(batch)
        loss = model.compute_loss(u, i, item_ids=batch["item_tower"]["sparse"][:, 0], hard_neg_emb=hn, temperature=0.15)
        loss.backward()
        if not fused:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return float(loss)

    b1 = builder.batch(torch.arange(0, 128))
    b2 = builder.batch(torch.arange(128, 256))
    m_a = fresh()
    o_a = tt.FusedTwoTowerOptimizer(m_a, lr=1e-3, max_grad_norm=1.0, table_mode="dense")
    step(m_a, o_a, b1, True)
    path = str(tmp_path / "ckpt" / "best_model_epoch_0.pt")
    training.save_checkpoint(path, 0, m_a, o_a, 1.0, 2.0, {10: 0.5}, synth.MAPS_C2[0], synth.MAPS_C2[1], cfg)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "train_loss", "val_loss", "metrics",
                         "user_mapping", "item_mapping", "config"}
    # resume into the reference's optimizer ...
    m_b = fresh()
    o_b = torch.optim.Adam(m_b.parameters(), lr=1e-3)
    meta = training.load_checkpoint(path, m_b, o_b)
    assert meta["epoch"] == 0 and meta["metrics"] == {10: 0.5}
    # ... and into a fresh fused optimizer
    m_c = fresh()
    o_c = tt.FusedTwoTowerOptimizer(m_c, lr=1e-3, max_grad_norm=1.0, table_mode="dense")
    training.load_checkpoint(path, m_c, o_c)
    step(m_a, o_a, b2, True)
    step(m_b, o_b, b2, False)
    step(m_c, o_c, b2, True)
    sa, sb, sc = m_a.state_dict(), m_b.state_dict(), m_c.state_dict()
    for k in sa:
        if sa[k].is_floating_point():
            assert torch.allclose(sa[k], sc[k], atol=1e-7, rtol=0), k          # same optimizer kind: same numbers
            assert torch.allclose(sa[k], sb[k], atol=2e-5, rtol=1e-4), k       # torch Adam after resume: fp32 noise only


def test_gpu_batch_builder_matches_the_reference_collate_fixture():
    """Pinned against the reference itself: tests/golden/collate_small.npz holds the batches the UNMODIFIED
    CombinedTwoTowerDataLoader (CombineTwoTower.py:62-92 -> DataLoader.py:226-288) yields for a 43-row DataFrame with the
    shipped config.yaml (batch 8, no shuffle: five full batches and a partial one), plus the DataFrame's columns."""
    from golden_io import unflatten
    from helpers import load_golden
    npz, meta = load_golden("collate_small")
    cols = unflatten(npz, "columns")
    ref_batches = unflatten(npz, "batches")
    builder = GpuBatchBuilder(cols["user"], cols["item"], device="cpu")
    assert len(builder) == meta["n"]
    bs = meta["batch_size"]
    assert len(ref_batches) == (meta["n"] + bs - 1) // bs
    for k, ref in enumerate(ref_batches):
        idx = torch.arange(k * bs, min(meta["n"], (k + 1) * bs))
        got = builder.batch(idx)
        assert set(got.keys()) == set(ref.keys())
        for tower in ("user_tower", "item_tower"):
            assert _same(got[tower], ref[tower]), (k, tower)
    # epoch(): same batches through the iterator interface
    for got, ref in zip(builder.epoch(bs, shuffle=False), ref_batches):
        assert _same(got["user_tower"], ref["user_tower"]) and _same(got["item_tower"], ref["item_tower"])
    # the reference's feature -> column mappings are what synth.MAPS_C2 hard-codes for the shipped config
    maps = unflatten(npz, "maps")
    assert {k: int(v) for k, v in maps["user"]["sparse"].items()} == synth.MAPS_C2[0]["sparse"]
    assert {k: int(v) for k, v in maps["item"]["sparse"].items()} == synth.MAPS_C2[1]["sparse"]
    assert {k: int(v) for k, v in maps["user"]["dense"].items()} == synth.MAPS_C2[0]["dense"]


def test_hard_negative_id_outside_the_catalog_range_maps_to_the_zero_row():
    """ids above the catalog's largest id used to be clamped onto the LAST catalog item."""
    user, item, catalog, catalog_ids, neg = _columns(50, 20, seed=9)
    neg[0, 0] = int(catalog_ids.max()) + 1000
    neg[1, 1] = -5
    b = GpuBatchBuilder(user, item, catalog, catalog_ids, neg, device="cpu").batch(torch.tensor([0, 1]))
    assert int(b["hard_negatives"][0]["sparse"][0].abs().sum()) == 0
    assert int(b["hard_negatives"][1]["sparse"][1].abs().sum()) == 0
