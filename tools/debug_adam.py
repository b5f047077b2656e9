import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_io import unflatten
from helpers import get_maps, load_golden, to_device
import recommendsystemproject_b200 as tt
DEV = "cuda"
npz, cfg = load_golden("seq_small")
key = 'user_tower.seq_encoder.transformer_backbone.layers.1.linear1.weight'
res = {}
for grouped in (False, True):
    umap, imap = get_maps(npz)
    model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), umap, imap)
    model.load_state_dict(unflatten(npz, "state0")); model = model.to(DEV).train()
    model.group_hard_negatives = grouped
    opt = tt.FusedTwoTowerOptimizer(model, lr=cfg["train"]["learning_rate"], max_grad_norm=1.0, table_mode="dense")
    T = cfg["train"]["temperature"]
    out = []
    for step, b in enumerate(("batch", "batch2")):
        batch = to_device(unflatten(npz, b), DEV)
        opt.zero_grad()
        u, i, hn = model(batch)
        loss = model.compute_loss(u, i, item_ids=batch["item_tower"]["sparse"][:, 0], hard_neg_emb=hn, temperature=T)
        loss.backward()
        g = dict(model.named_parameters())[key].grad.clone().cpu()
        opt.step()
        out.append((g, dict(model.named_parameters())[key].detach().clone().cpu(), float(opt.total_norm)))
    res[grouped] = out
gold1 = unflatten(npz, "step1")
v = gold1["state_after"][key]
d = (res[True][1][1] - v).abs()
idx = torch.nonzero(d > 1e-5)
print("n bad", len(idx), "of", d.numel())
for r, c in idx[:8].tolist():
    print((r, c), "got", float(res[True][1][1][r, c]), "gold", float(v[r, c]), "perslab", float(res[False][1][1][r, c]))
    for st in (0, 1):
        gg = unflatten(npz, f"step{st}")["grads"][key][r, c]
        print("   step", st, "g gold", float(gg), "g grouped", float(res[True][st][0][r, c]), "g perslab", float(res[False][st][0][r, c]),
              "tn", res[True][st][2], float(unflatten(npz, f"step{st}")["total_norm"]))
