"""Full-size golden fixtures: the UNMODIFIED reference (imported from /root/reference, authoring container only) run
at the sizes BASELINE.json quotes --

  c2_full : /root/reference/config.yaml VERBATIM (Transformer 2L/4H/d64/FFN256, L=20, 3 genre tags), B=512, 10
            hard-negative item slabs, dropout set to 0 (parity needs a deterministic forward)
  c1_full : configs[0]: ML-1M-shaped, mean-pooled L=50 history, D=64, B=1024

two training steps (zero_grad -> model -> compute_loss -> backward -> clip_grad_norm_(1.0) -> Adam.step,
training_utils.py:28-60) and one retrieval pass (training_utils.py:153-258).  The batches come from
recommendsystemproject_b200/synth.py (seeded, CPU); weights from torch.manual_seed(seed) + the reference's own
constructors.  To keep the fixtures small only digests are stored: losses, norms, the first rows of every embedding
output, per-tensor (sum, sum|.|, first 16 values) of gradients and of the state after step 2, full tensors below 4096
elements, and the top-K ids of the retrieval pass.

    python tests/golden/make_golden_full.py
"""
import copy
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from golden_io import save_case  # noqa: E402
from make_golden import import_reference, REF  # noqa: E402
from recommendsystemproject_b200 import synth  # noqa: E402


def digest(t: torch.Tensor):
    t = t.detach().double().reshape(-1)
    d = {"sum": t.sum(), "abs": t.abs().sum(), "head": t[:16].clone(), "numel": torch.tensor(t.numel())}
    if t.numel() <= 4096:
        d["full"] = t.clone()
    return d


def reference_classes():
    """The reference's `project` is a namespace package (no __init__.py); this repo's import-path shim `project/` is a
    regular package and would win whatever the path order: keep the repo root off sys.path while importing."""
    saved = list(sys.path)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    try:
        return import_reference()
    finally:
        sys.path[:] = saved


def run(name, cfg, maps, batches, corpus, seed, k=50):
    GenericTower, TwoTowerModel = reference_classes()
    assert "reference" in sys.modules[GenericTower.__module__].__file__, "not the reference's GenericTower"
    torch.manual_seed(seed)
    model = TwoTowerModel(GenericTower(cfg, "user_tower"), GenericTower(cfg, "item_tower"), *maps)
    T, lr = cfg["train"]["temperature"], cfg["train"]["learning_rate"]
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    model.train()
    steps = []
    for b in batches:
        opt.zero_grad()
        u, i, hn = model(b)
        ids = b["item_tower"]["sparse"][:, 0]
        loss = model.compute_loss(u, i, hard_neg_emb=hn, item_ids=ids, temperature=T)
        loss.backward()
        grads = {n: digest(torch.zeros_like(p) if p.grad is None else p.grad) for n, p in model.named_parameters()}
        tn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        steps.append({"loss": loss.detach(), "total_norm": tn, "grads": grads, "u_head": u[:32].detach(),
                      "i_head": i[:32].detach(), "hn_head": None if hn is None else hn[:8].detach(),
                      "u": digest(u), "i": digest(i)})
    state_after = {n: digest(v) for n, v in model.state_dict().items() if v.dtype.is_floating_point}
    model.eval()
    with torch.no_grad():
        emb = model.get_item_embeddings(corpus)
        uq, _, _ = model({k2: v for k2, v in batches[0].items() if k2 != "hard_negatives"})
        scores = uq @ emb.t()
        topv, topi = torch.topk(scores, k=k, dim=1)
        # margin between the k-th and (k+1)-th score: rows whose margin is below fp32 noise may legitimately differ
        top2 = torch.topk(scores, k=k + 1, dim=1).values
        margin = top2[:, k - 1] - top2[:, k]
    save_case(os.path.join(HERE, name + ".npz"), cfg, seed=torch.tensor(seed), step0=steps[0], step1=steps[1],
              state_after=state_after,
              retrieval={"topk_idx": topi.to(torch.int32), "topk_vals": topv, "margin": margin, "corpus": digest(emb),
                         "queries_head": uq[:16]})
    print(name, "loss", [float(s["loss"]) for s in steps], "norm", [float(s["total_norm"]) for s in steps])


def main():
    torch.set_num_threads(8)
    # ---- C2: the shipped YAML, verbatim, dropout 0
    with open(os.path.join(REF, "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    for tower in ("user_tower", "item_tower"):
        cfg["two_tower"][tower]["dropout"] = 0.0
        cfg["two_tower"][tower]["transformer_parameters"]["dropout"] = 0.0
    run("c2_full", cfg, synth.MAPS_C2, [synth.make_batch_c2(512, 20, 10, seed=21), synth.make_batch_c2(512, 20, 10, seed=22)],
        synth.make_corpus_c2(3416, seed=7), seed=31)
    # ---- C1
    cfg1 = synth.config_c1(dropout=0.0)
    gen = torch.Generator().manual_seed(8)
    corpus1 = {"sparse": torch.arange(1, 3707).unsqueeze(1)}
    run("c1_full", cfg1, synth.MAPS_C1, [synth.make_batch_c1(1024, 50, seed=41), synth.make_batch_c1(1024, 50, seed=42)],
        corpus1, seed=32)


if __name__ == "__main__":
    main()
