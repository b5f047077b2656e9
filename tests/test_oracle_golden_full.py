"""Pin the CPU oracle at FULL size: the unmodified reference's outputs for the verbatim shipped config.yaml (B=512, 10
hard-negative slabs) and for C1 (B=1024, L=50, D=64), stored as digests in tests/golden/c{1,2}_full.npz by
tests/golden/make_golden_full.py.  Weights come from the same torch.manual_seed + this repo's constructors (their
equality with the reference's is test_cabi_and_host.py's same-seed test; here it is re-checked through the step-0
outputs)."""
import numpy as np
import pytest
import torch

from golden_io import unflatten
from helpers import check_digest, full_case
from oracle import twotower_oracle as O
import recommendsystemproject_b200 as tt


@pytest.mark.parametrize("name", ["c1_full", "c2_full"])
def test_oracle_two_steps_and_retrieval_at_full_size(name):
    npz, cfg, maps, batches, corpus, seed = full_case(name)
    torch.manual_seed(seed)
    model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), *maps)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    opt = {"step": 0, "m": {}, "v": {}}
    T, lr = cfg["train"]["temperature"], cfg["train"]["learning_rate"]
    for step, batch in enumerate(batches):
        gold = unflatten(npz, f"step{step}")
        u, i, hn = O.two_tower_forward(batch, dict(state), cfg, *maps, training=True)
        assert torch.allclose(u[:32], gold["u_head"], atol=3e-6), step
        assert torch.allclose(i[:32], gold["i_head"], atol=3e-6), step
        if hn is not None:
            assert torch.allclose(hn[:8], gold["hn_head"], atol=3e-6), step
        loss, grads = O.train_step(batch, state, opt, cfg, *maps, temperature=T, lr=lr)
        assert abs(float(loss) - float(gold["loss"])) < 1e-5, (step, float(loss), float(gold["loss"]))
        _, total = O.clip_coef(list(grads.values()))
        assert abs(total - float(gold["total_norm"])) < 1e-4 * total
        for k, d in gold["grads"].items():
            # step 1 starts from weights that carry Adam's +-lr amplification of step 0's rounding-noise gradients
            check_digest(grads[k], d, (step, k), rtol=1e-3 if step == 0 else 1e-2, atol_head=5e-6 if step == 0 else 3e-4,
                         atol_elem=3e-7 if step == 0 else 3e-6)
    after = unflatten(npz, "state_after")
    g0 = unflatten(npz, "step0")["grads"]
    for k, d in after.items():
        # Adam turns a gradient that is pure rounding noise (a BatchNorm bias feeding another BatchNorm: analytically 0)
        # into +-lr steps (see test_oracle_golden.py): such tensors are not comparable between two fp32 implementations
        if k in g0 and float(g0[k]["abs"]) / int(g0[k]["numel"]) < 1e-6:
            continue
        slack = 4 * lr if k.endswith("running_mean") else 0.0      # those walks move downstream BN means by O(lr)
        check_digest(state[k], d, k, rtol=2e-3, atol_head=2.5 * lr, atol_elem=slack,
                     atol_sum=2e-3 * float(d["abs"]) + 4 * lr * (int(d["numel"]) ** 0.5) + slack * int(d["numel"]) + 1e-6)
    # retrieval: eval-mode corpus encode + exact top-K (training_utils.py:153-258)
    r = unflatten(npz, "retrieval")
    emb = O.tower_forward(corpus, maps[1], state, "item_tower", cfg, training=False)
    check_digest(emb, r["corpus"], "corpus embeddings", rtol=5e-4, atol_head=5e-4)
    q0 = {k: v for k, v in batches[0].items() if k != "hard_negatives"}
    uq, _, _ = O.two_tower_forward(q0, dict(state), cfg, *maps, training=False)
    k = r["topk_idx"].shape[1]
    _, idx = O.score_topk(uq.numpy(), emb.numpy(), k)
    # two Adam steps leave +-lr noise on rounding-noise gradient elements, so scores move by up to ~1e-3 and ids near
    # the K-th boundary may swap: compare the top-K VALUES, and the id sets through their overlap
    vals, idx = O.score_topk(uq.numpy(), emb.numpy(), k)
    assert np.allclose(vals, r["topk_vals"].numpy(), atol=3e-3)
    ref = r["topk_idx"].numpy().astype(np.int64)
    overlap = np.mean([len(set(a) & set(b)) / k for a, b in zip(idx, ref)])
    assert overlap > 0.97, overlap
