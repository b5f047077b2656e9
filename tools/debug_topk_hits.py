"""Debug helper: what do candidate hits cost the tensor-core top-K main pass?
A: normalised random corpus (every query sees ~384 scores above its sampled threshold).
B: the same corpus with every NON-sampled 256-row tile scaled by 0.5 -- the sampled thresholds stay where they
   were, nothing outside the sampled tiles beats them, so the main pass runs on its fast path only (the results
   then fail the proof obligation; only the time is of interest).
"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystemproject_b200 import ops

N = int(os.environ.get("N", 1_250_000)); Q = int(os.environ.get("Q", 16384)); D = 128; K = 100
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
corpus = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=gen), dim=1)
query = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=gen), dim=1)
lib = ops._lib.load()


def run(c, label):
    prep = ops.PreparedCorpus(c)
    scores = torch.empty(Q, K, dtype=torch.float64, device=dev); idx = torch.empty(Q, K, dtype=torch.int64, device=dev)
    bad = torch.empty(Q, dtype=torch.int32, device=dev)
    nb = ctypes.c_size_t(0)
    ops.check(lib.tt_score_topk_tc_workspace(Q, N, D, K, 0, ctypes.byref(nb)), "ws")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    def call():
        ops.check(lib.tt_score_topk_tc(ops._p(query), Q, ops._p(c), ops._p(prep.bf16), ops._p(prep.bounds), N, D, K, 0,
                                       None, None, ops._p(scores), ops._p(idx), ops._p(bad), 1, ops._p(ws), ws.numel(),
                                       ops._stream()), "topk")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        call()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{label}: {ms:.3f} ms per call  ({2*Q*N*D/ms/1e9:.0f} TF/s)  unverified {int(bad.sum())}", flush=True)


run(corpus, "A normal")
tile = torch.arange(N, device=dev) // 256
scaled = corpus * torch.where(tile % 16 == 0, 1.0, 0.5)[:, None]
run(scaled, "B hits only in sampled tiles")
