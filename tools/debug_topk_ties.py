"""Debug helper: where does score_topk differ from the oracle on a corpus with duplicated rows?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from recommendsystemproject_b200 import ops
from oracle import twotower_oracle as O

gen = torch.Generator().manual_seed(9)
base = torch.nn.functional.normalize(torch.randn(40, 32, generator=gen), dim=1)
e = base[torch.randint(0, 40, (500,), generator=gen)]
q = torch.nn.functional.normalize(torch.randn(17, 32, generator=gen), dim=1)
vals_ref, idx_ref = O.score_topk(q.numpy(), e.numpy(), 60)
vals, idx = ops.score_topk(q.cuda(), e.cuda(), 60)
vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
bad = np.argwhere(idx != idx_ref)
print("mismatches:", len(bad))
for r, c in bad[:12]:
    print(f"q{r} pos{c}: got row {idx[r,c]} score {vals[r,c]!r}; ref row {idx_ref[r,c]} score {vals_ref[r,c]!r}; "
          f"same set: {set(idx[r]) == set(idx_ref[r])}")
    print("   neighbours got", idx[r, max(0,c-2):c+3], vals[r, max(0,c-2):c+3])
    print("   neighbours ref", idx_ref[r, max(0,c-2):c+3], vals_ref[r, max(0,c-2):c+3])
