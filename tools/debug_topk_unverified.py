"""Debug helper: which queries of the C5 workload fail the tensor-core top-K proof obligation, and by how much?
Replays bench.py's c5 shapes (10M x 128 corpus, Q=16384), finds the queries whose `unverified` flag is set, and for
each prints the exact K-th score, the K'-th best bf16-approximate score (upper bound of tau_max) and the margin."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystemproject_b200 import ops

N = int(os.environ.get("N", 10_000_000)); Q = int(os.environ.get("Q", 16384)); D = 128; K = 100
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
corpus = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=gen), dim=1)
qgen = torch.Generator().manual_seed(6)
query = torch.nn.functional.normalize(torch.randn(Q, D, generator=qgen), dim=1).to(dev)
prep = ops.PreparedCorpus(corpus)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


for sampling in (ops.TOPK_SAMPLING, ops.TOPK_WIDE):
    s, i = ops._score_topk_tc(query, corpus, K, 0, None, None, prep, sampling)
    torch.cuda.synchronize()
    print(f"sampling={sampling}: stats {ops.topk_stats}  {timed(lambda: ops._score_topk_tc(query, corpus, K, 0, None, None, prep, sampling)):.2f} ms", flush=True)

# flag straight from the C entry point, without the python-side repair
lib = ops._lib.load()
import ctypes
for sampling in (1, 2):
    scores = torch.empty(Q, K, dtype=torch.float64, device=dev); idx = torch.empty(Q, K, dtype=torch.int64, device=dev)
    bad = torch.empty(Q, dtype=torch.int32, device=dev)
    nb = ctypes.c_size_t(0)
    ops.check(lib.tt_score_topk_tc_workspace(Q, N, D, K, 0, ctypes.byref(nb)), "ws")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ops.check(lib.tt_score_topk_tc(ops._p(query), Q, ops._p(corpus), ops._p(prep.bf16), ops._p(prep.bounds), N, D, K, 0,
                                   None, None, ops._p(scores), ops._p(idx), ops._p(bad), sampling, ops._p(ws), ws.numel(),
                                   ops._stream()), "topk")
    redo = torch.nonzero(bad).reshape(-1)
    print(f"C entry sampling={sampling}: unverified rows {redo.tolist()[:20]} (n={redo.numel()})")
    for r in redo.tolist()[:4]:
        q = query[r]
        exact = (corpus.double() @ q.double()) if N <= 2_000_000 else torch.cat([(corpus[a:a + 1_000_000].double() @ q.double()) for a in range(0, N, 1_000_000)])
        approx = (prep.bf16.float() @ q.bfloat16().float())
        ev, _ = exact.topk(160); av, _ = approx.topk(200)
        margin = (1.02 * 2 ** -8 + 1.6e-5) * float(q.norm()) * float(prep.bounds)
        print(f"  row {r}: exact[K]={ev[K-1]:.6f} exact[K+1]={ev[K]:.6f} approx[150]={av[149]:.6f} approx[151]={av[150]:.6f} margin={margin:.6f} "
              f"=> provable with K'=150: {float(ev[K-1]) > float(av[150]) + margin};  #approx >= exact[K]-margin: {int((approx >= float(ev[K-1]) - margin).sum())}")
