#!/usr/bin/env python
"""Round 2: one pass over the kernels that are new or changed this round, at the shapes of the integrated C3 step on one
GPU (global batch 65536), for a single `ncu --set full` capture (profiles/r2_hot_kernels_ncu.md):
  row-sharded lookup through the group (route, owner-side gather + pool over int32 CSR lists, combine), its backward
  (gradient pack, list-form segment reduction) and the row-wise Adam; TF32 tcgen05 Linear fwd / dgrad / wgrad at the
  user tower's first layer; fused BatchNorm + ReLU + Dropout fwd / bwd; the fused CE at the headline shape; top-K shard.
Each op runs twice; read the second launch of each kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystemproject_b200 import ops, sharded

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(3)
B, L, D, V = 65536, 200, 128, 10_000_001
grp = sharded.ShardedTableGroup(0, 1, dev)
w = torch.empty(V, D, device=dev).uniform_(-0.01, 0.01)
grp.add_table("hist", V, D, ops.POOL_MEAN, 0, w, w[0].clone())
grp.init_state()
ids = torch.randint(1, V, (B, L), device=dev, generator=gen)
lens = torch.randint(1, L + 1, (B,), device=dev, generator=gen)
ids[torch.arange(L, device=dev)[None, :] >= lens[:, None]] = 0
up = torch.randn(B, D, device=dev)
step = torch.ones(1, dtype=torch.int64, device=dev)
coef = torch.ones(1, device=dev)
for _ in range(int(os.environ.get("TT_PROF_REPS", "2"))):
    grp.zero_grad()
    out = grp.lookup({"hist": ids})["hist"]
    (out * up).sum().backward()
    grp.step(coef, 5e-4, step)
torch.cuda.synchronize()
del grp, w, out
torch.cuda.empty_cache()
# tower layer 1 of the user tower: [65536, 640] -> 256, TF32
torch.backends.cuda.matmul.allow_tf32 = True
x = torch.randn(B, 640, device=dev, requires_grad=True)
wt = (torch.randn(256, 640, device=dev) * 0.05).requires_grad_(True)
bs = torch.zeros(256, device=dev, requires_grad=True)
gamma, beta = torch.ones(256, device=dev, requires_grad=True), torch.zeros(256, device=dev, requires_grad=True)
rm, rv, nb = torch.zeros(256, device=dev), torch.ones(256, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
seed = torch.tensor([7], dtype=torch.int64, device=dev)
for _ in range(int(os.environ.get("TT_PROF_REPS", "2"))):
    y = ops.linear(x, wt, bs)
    z, _, _ = ops.batch_norm_act(y, gamma, beta, rm, rv, nb, 0.1, 1e-5, 256, relu=True, dropout_p=0.1, seed_dev=seed, call_id=1)
    z.sum().backward()
torch.cuda.synchronize()
del x, y, z
torch.cuda.empty_cache()
# the fused CE at the headline shape (N = 1: 65536 x 65536) and at one rank's slab of the 8-GPU run (8192 x 65536)
u = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
it = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
iid = torch.randint(1, 10_000_001, (B,), device=dev)
for _ in range(int(os.environ.get("TT_PROF_REPS", "2"))):
    ops.fused_inbatch_ce(u, it, iid, None, None, 0.05, precision="bf16")[0].backward()
us = u[:8192].detach().requires_grad_(True)
for _ in range(int(os.environ.get("TT_PROF_REPS", "2"))):
    ops.fused_inbatch_ce(us, it, iid, None, None, 0.05, precision="bf16", item_offset=8192)[0].backward()
torch.cuda.synchronize()
q = torch.nn.functional.normalize(torch.randn(16384, D, device=dev), dim=1)
e = torch.nn.functional.normalize(torch.randn(1_250_000, D, device=dev), dim=1)
prep = ops.PreparedCorpus(e)
for _ in range(int(os.environ.get("TT_PROF_REPS", "2"))):
    ops.score_topk(q, e, 100, precision="bf16", prepared=prep)
torch.cuda.synchronize()
print("prof_hot_r2 done")
