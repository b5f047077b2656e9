#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, L2 flushed between reps).  Usage:
    python tools/kbench.py ce_tc [B H D]      fused CE forward, tcgen05 path vs fp32 SIMT path
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import load_peaks, time_op  # noqa: E402
from recommendsystemproject_b200 import ops  # noqa: E402


def main():
    what = sys.argv[1]
    dev = "cuda"
    peaks = load_peaks()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush = lambda: flush_buf.fill_(1)  # noqa: E731
    if what == "ce_tc":
        B, H, D = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (65536, 4096, 128)
        u = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
        it = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1).requires_grad_(True)
        pool = torch.nn.functional.normalize(torch.randn(H, D, device=dev), dim=1).requires_grad_(True) if H else None
        ids = torch.randint(1, B * 50, (B,), device=dev)
        out = {}
        for prec in ("bf16", "fp32") if B <= 32768 else ("bf16",):
            res = {}

            def f():
                res["l"] = ops.fused_inbatch_ce(u, it, ids, None, pool, 0.05, precision=prec)[0]
            ms, best = time_op(f, 5, flush)
            flops = 2.0 * B * (B + H) * D
            out[prec] = {"ms": ms, "best_ms": best, "tflops": flops / ms / 1e9,
                         "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"], "loss": float(res["l"])}
            if prec == "bf16" or B <= 16384:
                ms_fb, best_fb = time_op(lambda: (f(), res["l"].backward()), 5, flush)
                out[prec].update({"fwd_bwd_ms": ms_fb, "fwd_bwd_best_ms": best_fb, "fwd_bwd_tflops": 3 * flops / ms_fb / 1e9,
                                  "fwd_bwd_frac_of_bf16_peak": 3 * flops / ms_fb / 1e9 / peaks["bf16_tflops"]})
        print(json.dumps({"kernel": "ce_fwd", "B": B, "H": H, "D": D, **out}))


def bench_topk():
    dev = "cuda"
    peaks = load_peaks()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush = lambda: flush_buf.fill_(1)  # noqa: E731
    Q, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (16384, 1_250_000, 100)
    q = torch.nn.functional.normalize(torch.randn(Q, 128, device=dev), dim=1)
    e = torch.nn.functional.normalize(torch.randn(N, 128, device=dev), dim=1)
    prep = ops.PreparedCorpus(e)
    out = {}
    ms, best = time_op(lambda: ops.score_topk(q, e, K, precision="bf16", prepared=prep), 3, flush)
    flops = 2.0 * Q * N * 128
    out["bf16"] = {"ms": ms, "best_ms": best, "tflops": flops / best / 1e9, "frac_of_bf16_peak": flops / best / 1e9 / peaks["bf16_tflops"],
                   "queries_per_s": Q / best * 1e3, "unverified": ops.topk_stats["unverified"]}
    if Q * N <= 4096 * 1_000_000:
        ms, best = time_op(lambda: ops.score_topk(q, e, K), 2, flush)
        out["fp32"] = {"ms": ms, "best_ms": best, "tflops": flops / best / 1e9, "queries_per_s": Q / best * 1e3}
    print(json.dumps({"kernel": "score_topk", "Q": Q, "N": N, "K": K, **out}))


def bench_seg():
    dev = "cuda"
    peaks = load_peaks()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush = lambda: flush_buf.fill_(1)  # noqa: E731
    B, L, D, V = 65536, 200, 128, 10_000_001
    gen = torch.Generator(device=dev).manual_seed(3)
    ids = torch.randint(1, V, (B, L), device=dev, generator=gen)
    lens = torch.randint(1, L + 1, (B,), device=dev, generator=gen)
    ids[torch.arange(L, device=dev)[None, :] >= lens[:, None]] = 0
    n_valid = int((ids != 0).sum().item())
    g = torch.randn(B, D, device=dev)
    sq = torch.zeros(1, device=dev)
    res = {}

    def seg():
        res["r"] = ops.segment_grad(ids, ops.POOL_MEAN, 0, V, g, None, D, sq)
    ms, best = time_op(seg, 3, flush)
    U = int(res["r"][2].item())
    alg = B * L * 8 + n_valid * D * 4 + U * (D * 4 + 8)
    print(json.dumps({"kernel": "emb_segment_grad", "ms": ms, "best_ms": best, "U": U, "n_valid": n_valid,
                      "GBps": alg / best / 1e6, "frac": alg / best / 1e6 / peaks["hbm_gbs"]}))


if __name__ == "__main__":
    if sys.argv[1] == "seg":
        bench_seg()
        sys.exit(0)
    if sys.argv[1] == "topk":
        bench_topk()
        sys.exit(0)
    main()
