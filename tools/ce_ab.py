#!/usr/bin/env python
"""A/B of the tensor-core in-batch CE: single-pass training form (forward + dU in one walk, then dI) against the
three-pass kernels, forward + backward, CUDA events.  python tools/ce_ab.py [B_loc B_glob H D]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import load_peaks, time_op  # noqa: E402
from recommendsystemproject_b200 import ops  # noqa: E402


def main():
    dev = "cuda"
    peaks = load_peaks()
    cases = [tuple(int(x) for x in sys.argv[1:5])] if len(sys.argv) >= 5 else [(65536, 65536, 0, 128), (65536, 65536, 4096, 128),
                                                                                (8192, 65536, 0, 128), (16384, 16384, 0, 64)]
    for B, Bg, H, D in cases:
        u = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
        it = torch.nn.functional.normalize(torch.randn(Bg, D, device=dev), dim=1).requires_grad_(True)
        pool = torch.nn.functional.normalize(torch.randn(H, D, device=dev), dim=1).requires_grad_(True) if H else None
        ids = torch.randint(1, 10_000_000, (Bg,), device=dev)
        out = {"B_loc": B, "B_glob": Bg, "H": H, "D": D}
        grads = {}
        for name, sp in (("three_pass", False), ("single_pass", True)):
            res = {}

            def f():
                res["l"], _, res["f"] = ops.fused_inbatch_ce(u, it, ids, None, pool, 0.05, precision="bf16", item_offset=0,
                                                             id_bits=24, single_pass=sp)
                res["l"].backward()
            ms, best = time_op(f, 6, lambda: None)
            u.grad = it.grad = None
            f()
            grads[name] = (u.grad.clone(), it.grad.clone())
            flops = 6.0 * B * (Bg + H) * D
            out[name] = {"ms": ms, "best_ms": best, "tflops": flops / ms / 1e9, "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"],
                         "loss": float(res["l"]), "flags": int(res["f"])}
        out["dU_rel_diff"] = float((grads["single_pass"][0] - grads["three_pass"][0]).norm() / grads["three_pass"][0].norm())
        out["dI_rel_diff"] = float((grads["single_pass"][1] - grads["three_pass"][1]).norm() / grads["three_pass"][1].norm())
        print(json.dumps(out), flush=True)
        del u, it, pool, grads
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
