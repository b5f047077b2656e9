#!/usr/bin/env python
"""Round 2, late kernels, for ncu (one GPU):  TT_PROF_PART=ce  the single-pass CE (ce_tc_kernel<128, 3> = forward + dU,
ce_tc_kernel<128, 2> = dI) at the headline shape 65536 x 65536, D = 128;  TT_PROF_PART=seg  the deferred segment gradient
(norm-only seg_reduce_rows_wide, then seg_adam_rows_wide) on the 10M x 128 table, B = 65536, L = 200.
The op runs twice; read the second launch of each kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from recommendsystemproject_b200 import ops, sharded  # noqa: E402

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(3)
B, L, D, V = 65536, 200, 128, 10_000_001
part = os.environ.get("TT_PROF_PART", "ce")
if part == "ce":
    u = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
    it = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
    iid = torch.randint(1, V, (B,), device=dev)
    for _ in range(2):
        ops.fused_inbatch_ce(u, it, iid, None, None, 0.05, precision="bf16", id_bits=24, single_pass=True)[0].backward()
else:
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
    for _ in range(2):
        grp.zero_grad()
        (grp.lookup({"hist": ids})["hist"] * up).sum().backward()
        grp.step(coef, 5e-4, step)
torch.cuda.synchronize()
print("prof_r2b done", part)
