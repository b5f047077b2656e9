#!/usr/bin/env python
"""One pass over every hot-path kernel at a BASELINE-scale shape, for a single `ncu --set full` capture
(profiles/): gather+pool (C3), segment grad + row-wise Adam (C3), fused CE fwd + bwd (tcgen05), top-K (tcgen05).
Each op runs twice; profile the second launch of each kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystemproject_b200 import ops

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(3)
B, L, D, V = 65536, 200, 128, 10_000_001
table = torch.empty(V, D, device=dev).uniform_(-0.01, 0.01)
ids = torch.randint(1, V, (B, L), device=dev, generator=gen)
lens = torch.randint(1, L + 1, (B,), device=dev, generator=gen)
ids[torch.arange(L, device=dev)[None, :] >= lens[:, None]] = 0
out = torch.empty(B, D, device=dev)
oob = torch.zeros(1, dtype=torch.int32, device=dev)
g = torch.randn(B, D, device=dev)
sq = torch.zeros(1, device=dev)
m, v = torch.zeros_like(table), torch.zeros_like(table)
step = torch.ones(1, dtype=torch.int64, device=dev)
coef = torch.ones(1, device=dev)
for _ in range(2):
    ops.gather_pool_into(table, ids, ops.POOL_MEAN, 0, out, None, oob)
    rows, rg, nu = ops.segment_grad(ids, ops.POOL_MEAN, 0, V, g, None, D, sq)
    ops.rowwise_adam_(table, m, v, rows, rg, nu, coef, 5e-4, 0.9, 0.999, 1e-8, step)
torch.cuda.synchronize()
del table, m, v, rows, rg
torch.cuda.empty_cache()
Bc, Hc = 16384, 2048
u = torch.nn.functional.normalize(torch.randn(Bc, D, device=dev), dim=1).requires_grad_(True)
it = torch.nn.functional.normalize(torch.randn(Bc, D, device=dev), dim=1).requires_grad_(True)
pool = torch.nn.functional.normalize(torch.randn(Hc, D, device=dev), dim=1).requires_grad_(True)
iid = torch.randint(1, Bc * 50, (Bc,), device=dev)
for _ in range(2):
    loss = ops.fused_inbatch_ce(u, it, iid, None, pool, 0.05, precision="bf16")[0]
    loss.backward()
torch.cuda.synchronize()
q = torch.nn.functional.normalize(torch.randn(16384, D, device=dev), dim=1)
e = torch.nn.functional.normalize(torch.randn(1_250_000, D, device=dev), dim=1)
prep = ops.PreparedCorpus(e)
for _ in range(2):
    ops.score_topk(q, e, 100, precision="bf16", prepared=prep)
torch.cuda.synchronize()
print("prof_hot done")
