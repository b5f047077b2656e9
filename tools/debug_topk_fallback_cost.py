"""Debug helper: cost of the two repair paths of the tensor-core top-K for a handful of queries on a 10M corpus."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystemproject_b200 import ops

N = int(os.environ.get("N", 10_000_000)); D = 128; K = 100
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
corpus = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=gen), dim=1)
prep = ops.PreparedCorpus(corpus)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


for Q in (1, 4, 64, 128, 1024):
    q = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=gen), dim=1)
    a = timed(lambda: ops.score_topk(q, corpus, K, precision="fp32"))
    b = timed(lambda: ops._score_topk_tc(q, corpus, K, 0, None, None, prep, ops.TOPK_WIDE))
    c = timed(lambda: ops._score_topk_tc(q, corpus, K, 0, None, None, prep, ops.TOPK_SAMPLING))
    print(f"Q={Q}: fp32 path {a:.2f} ms | tc no-sampling {b:.2f} ms | tc sampling {c:.2f} ms  stats {ops.topk_stats}", flush=True)
