#!/usr/bin/env python
"""Timeline of CTA 0 of the tcgen05 CE kernels (developer tool): per tile, when the TMA was issued, when the
MMA thread issued S / Out, when the softmax group saw S, finished loading, finished the exponentials, got the G
buffer and published G.  Usage: python tools/ce_trace.py [B H D] [mode]   (mode: fwd | bwd)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystemproject_b200 import _lib, ops

B, H, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16384, 2048, 128)
mode = sys.argv[4] if len(sys.argv) > 4 else "bwd"
dev = "cuda"
u = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
it = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
pool = torch.nn.functional.normalize(torch.randn(H, D, device=dev), dim=1).requires_grad_(True)
ids = torch.randint(1, B * 50, (B,), device=dev)
lib = _lib.load()
for _ in range(2):
    loss = ops.fused_inbatch_ce(u, it, ids, None, pool, 0.05, precision="bf16")[0]
    loss.backward()
torch.cuda.synchronize()
dbg = torch.zeros(11, 256, dtype=torch.int64, device=dev)
loss = ops.fused_inbatch_ce(u, it, ids, None, pool, 0.05, precision="bf16")[0]
torch.cuda.synchronize()
if mode == "fwd":
    lib.tt_ce_tc_debug_trace(ctypes.c_void_p(dbg.data_ptr()))
    loss = ops.fused_inbatch_ce(u, it, ids, None, pool, 0.05, precision="bf16")[0]
    torch.cuda.synchronize()
    lib.tt_ce_tc_debug_trace(None)
else:
    # only the first backward pass (dU) writes meaningful stamps: the second pass overwrites them -> capture after
    lib.tt_ce_tc_debug_trace(ctypes.c_void_p(dbg.data_ptr()))
    loss.backward()
    torch.cuda.synchronize()
    lib.tt_ce_tc_debug_trace(None)
t = dbg.cpu().numpy()
names = ["tma_issue", "S_wait_full", "S_issue", "O_wait_gfull", "O_issue", "sm_wait_sfull", "sm_got_S", "sm_loaded",
         "sm_computed", "sm_got_G", "sm_published"]
t0 = t[t > 0].min()
print("tile " + " ".join(f"{n:>13s}" for n in names))
for i in range(40, 72):
    print(f"{i:4d} " + " ".join(f"{(t[e, i] - t0) if t[e, i] else -1:13d}" for e in range(11)))
import numpy as np
d = np.diff(t[2, 40:120].astype(np.int64))
print("mean cycles between S issues (tiles 40..120):", d.mean())
