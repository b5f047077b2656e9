#!/usr/bin/env python
"""A/B of the owner-side table backward + update at the C3 shape (10M x 128 table, B = 65536, L = 200 ragged):
segment reduction -> row_grad -> row-wise Adam (two kernels, 9 HBM streams of U x D x 4 bytes) against the deferred form
(norm-only reduction, then segment sum + Adam in one kernel: 7 streams).  CUDA events around backward + step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from recommendsystemproject_b200 import ops, sharded  # noqa: E402

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(3)
B, L, D, V = 65536, 200, 128, 10_000_001
ids = torch.randint(1, V, (B, L), device=dev, generator=gen)
lens = torch.randint(1, L + 1, (B,), device=dev, generator=gen)
ids[torch.arange(L, device=dev)[None, :] >= lens[:, None]] = 0
up = torch.randn(B, D, device=dev)
out_json = {"B": B, "L": L, "D": D, "V": V}
ref = None
for name, fused in (("two_kernel", False), ("deferred", True)):
    grp = sharded.ShardedTableGroup(0, 1, dev)
    grp.fused_adam = fused
    w = torch.empty(V, D, device=dev).uniform_(-0.01, 0.01, generator=gen) if ref is None else ref[0].clone()
    if ref is None:
        ref = (w.clone(),)
    grp.add_table("hist", V, D, ops.POOL_MEAN, 0, w, w[0].clone())
    grp.init_state()
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    coef = torch.ones(1, device=dev)
    t_b, t_s = [], []
    for it in range(5):
        grp.zero_grad()
        out = grp.lookup({"hist": ids})["hist"]
        loss = (out * up).sum()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        loss.backward()
        e[1].record()
        step += 1
        grp.step(coef, 5e-4, step)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            t_b.append(e[0].elapsed_time(e[1]))
            t_s.append(e[1].elapsed_time(e[2]))
    out_json[name] = {"backward_ms": sum(t_b) / len(t_b), "step_ms": sum(t_s) / len(t_s),
                      "total_ms": (sum(t_b) + sum(t_s)) / len(t_b), "checksum": float(w.double().sum()),
                      "m_checksum": float(grp.tables["hist"].exp_avg.double().abs().sum())}
    del grp, w, out, loss
    torch.cuda.empty_cache()
out_json["bitwise_equal_checksums"] = (out_json["two_kernel"]["checksum"] == out_json["deferred"]["checksum"] and
                                       out_json["two_kernel"]["m_checksum"] == out_json["deferred"]["m_checksum"])
print(json.dumps(out_json))
