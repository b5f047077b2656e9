#!/usr/bin/env python
"""bench.py -- headline benchmark of the two-tower DSSM hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1]

Prints ONE JSON line (rank 0).  A "step" is one full training step of the
reference loop (training_utils.py:28-60: zero_grad -> model(batch) ->
compute_loss -> backward -> clip_grad_norm_(1.0) -> Adam.step) on one batch of
synthetic input shaped like BASELINE.json configs[1] (shipped config.yaml:
Transformer 2L/4H/d64 user tower, B=512, 10 hard-negative slabs per step).

  value : samples/s with the batch already resident in HBM (CUDA-graph replay,
          CUDA events, L2 flushed between steps outside the timed region)
  e2e   : same metric through the public API with HOST (pinned) batches: H2D
          copy of the batch + step + D2H read of the loss inside the timed region
  roofline / kernels : the dominant kernel of the step and the hot-path kernels
          at their BASELINE-scale shapes, algorithmic bytes (flops) / CUDA-event
          time against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle port of the same step (oracle/twotower_oracle.py)
          timed on this box's host cores, bounded sample

--impl reference runs ONLY the CPU port (no GPU work) on the same config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def tree_to(obj, device, pin=False):
    if isinstance(obj, torch.Tensor):
        if pin:
            return obj.pin_memory()
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: tree_to(v, device, pin) for k, v in obj.items()}
    if isinstance(obj, list):
        return [tree_to(v, device, pin) for v in obj]
    return obj


def tree_bytes(obj):
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, dict):
        return sum(tree_bytes(v) for v in obj.values())
    if isinstance(obj, list):
        return sum(tree_bytes(v) for v in obj)
    return 0


def workload(name):
    from recommendsystemproject_b200 import synth
    if name == "c2":
        return dict(cfg=synth.config_c2(), maps=synth.MAPS_C2, batch_fn=lambda seed: synth.make_batch_c2(512, 20, 10, seed),
                    B=512, T=0.15, lr=5e-4,
                    desc="BASELINE configs[1]: shipped config.yaml (Transformer 2L/4H/d64/FFN256, L=20, 3 tags), "
                         "B=512, 10 hard-negative slabs, shipped dropout, fwd+bwd+clip+Adam")
    if name == "c1":
        cfg = synth.config_c1(dropout=0.1)
        return dict(cfg=cfg, maps=synth.MAPS_C1, batch_fn=lambda seed: synth.make_batch_c1(1024, 50, seed),
                    B=1024, T=0.15, lr=5e-4,
                    desc="BASELINE configs[0]: ML-1M-shaped, mean-pooled L=50 history, dim 64, B=1024, dropout 0.1")
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------- CPU port (reference arm)
def cpu_port_steps(wl, steps, warmup, seed=2):
    """Time the CPU oracle port of the training step with all host threads."""
    from oracle import twotower_oracle as O
    import recommendsystemproject_b200 as tt
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = wl["cfg"]
    # parity runs use dropout 0; the port has no dropout, which only makes the CPU arm faster
    torch.manual_seed(0)
    model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), *wl["maps"])
    state = {k: v.clone() for k, v in model.state_dict().items()}
    opt = {"step": 0, "m": {}, "v": {}}
    batch = wl["batch_fn"](seed)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(batch, state, opt, cfg, *wl["maps"], temperature=wl["T"], lr=wl["lr"])
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    steps = max(1, min(args.steps, 20))
    times = cpu_port_steps(wl, steps, max(1, min(args.warmup, 3)))
    total = sum(times)
    value = wl["B"] * len(times) / total
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd+clip+Adam)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "arm": "CPU port of the reference step (oracle/twotower_oracle.py), torch fp32, "
                                                  "dropout omitted"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full steps of B={wl['B']}"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- kernel rooflines
def time_op(fn, reps, flush):
    """Mean CUDA-event duration (ms) of fn() on the current stream, L2 flushed before each rep."""
    for _ in range(3):  # untimed warm-up: first-launch costs (function attributes, lazy module load) and the caching
        fn()            # allocator's first cudaMalloc of each output / workspace block stay out of the mean
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts), min(ts)


def kernel_rooflines(peaks, flush, quick=False):
    """Hot-path kernels at BASELINE-scale shapes (per-GPU slices of C3/C4/C5)."""
    from recommendsystemproject_b200 import ops
    dev = "cuda"
    out = []
    gen = torch.Generator(device=dev).manual_seed(3)
    # --- C3-shaped gather + pool: B=65536 samples, L=200 ragged history, D=128, 10M-row fp32 table (5.1 GB)
    B, L, D, V = (16384, 200, 128, 2_000_000) if quick else (65536, 200, 128, 10_000_001)
    table = torch.empty(V, D, device=dev).uniform_(-0.01, 0.01)
    ids = torch.randint(1, V, (B, L), device=dev, generator=gen)
    lens = torch.randint(1, L + 1, (B,), device=dev, generator=gen)
    ids[torch.arange(L, device=dev)[None, :] >= lens[:, None]] = 0
    n_valid = int((ids != 0).sum().item())
    outbuf = torch.empty(B, D, device=dev)
    oob = torch.zeros(1, dtype=torch.int32, device=dev)
    ms, best = time_op(lambda: ops.gather_pool_into(table, ids, ops.POOL_MEAN, 0, outbuf, None, oob), 5, flush)
    alg = n_valid * D * 4 + B * L * 8 + B * D * 4  # rows actually read (pads are counted, not read) + ids + out
    out.append({"kernel": "gather_pool_kernel", "workload": f"C3 slice: B={B} L={L} ragged mean-pool D={D} fp32, V={V}",
                "bound": "hbm", "ms": ms, "achieved": alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": alg / ms / 1e6 / peaks["hbm_gbs"], "alg_bytes": alg,
                "traffic": None if quick else 3.495e9, "traffic_source": "profiles/r1_hot_kernels_ncu.md (dram read+write, same shape)"})
    # --- sorted-segment gradient + row-wise Adam on the same ids
    g = torch.randn(B, D, device=dev)
    sq = torch.zeros(1, device=dev)
    res = {}

    def seg():
        res["r"] = ops.segment_grad(ids, ops.POOL_MEAN, 0, V, g, None, D, sq)
    ms, best = time_op(seg, 5, flush)
    rows, row_grad, nu = res["r"]
    U = int(nu.item())
    alg = B * L * 8 + n_valid * D * 4 + U * (D * 4 + 8)
    out.append({"kernel": "emb_segment_grad (keys + cub radix sort + scans + seg_reduce_rows_wide + norm)",
                "workload": f"same ids, U={U} unique rows", "bound": "hbm", "ms": ms, "best_ms": best, "achieved": alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": alg / ms / 1e6 / peaks["hbm_gbs"], "alg_bytes": alg})
    m = torch.zeros_like(table)
    v = torch.zeros_like(table)
    step = torch.ones(1, dtype=torch.int64, device=dev)
    coef = torch.ones(1, device=dev)
    ms, best = time_op(lambda: ops.rowwise_adam_(table, m, v, rows, row_grad, nu, coef, 5e-4, 0.9, 0.999, 1e-8, step), 5, flush)
    alg = U * (D * 4 * 7 + 8)  # grad read + 3 state reads + 3 state writes
    out.append({"kernel": "rowwise_adam_kernel", "workload": f"U={U} rows x D={D} fp32 state", "bound": "hbm", "ms": ms,
                "achieved": alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": alg / ms / 1e6 / peaks["hbm_gbs"], "alg_bytes": alg,
                "traffic": None if quick else 17.27e9, "traffic_source": "profiles/r1_hot_kernels_ncu.md (dram read+write, same shape)"})
    del table, m, v, ids, g, rows, row_grad, res
    torch.cuda.empty_cache()
    # --- C4 fused CE, tcgen05/TMA bf16 path, forward + backward: the B x (B+H) logits live only in TMEM
    Bc, Hc, Dc = (8192, 1024, 128) if quick else (65536, 4096, 128)
    u = torch.nn.functional.normalize(torch.randn(Bc, Dc, device=dev), dim=1).requires_grad_(True)
    it = torch.nn.functional.normalize(torch.randn(Bc, Dc, device=dev), dim=1).requires_grad_(True)
    pool = torch.nn.functional.normalize(torch.randn(Hc, Dc, device=dev), dim=1).requires_grad_(True)
    item_ids = torch.randint(1, Bc * 50, (Bc,), device=dev)
    res = {}

    def ce_f():
        res["l"] = ops.fused_inbatch_ce(u, it, item_ids, None, pool, 0.05, precision="bf16")[0]
    ms_f, best_f = time_op(ce_f, 5, flush)
    ms_b, best_b = time_op(lambda: (ce_f(), res["l"].backward()), 5, flush)
    flops = 6.0 * Bc * (Bc + Hc) * Dc
    out.append({"kernel": "ce_tc_kernel<fwd> + ce_tc_kernel<bwd dU> + ce_tc_kernel<bwd dI,dPool> (tcgen05, bf16 in / fp32 acc)",
                "workload": f"C4: B={Bc} H={Hc} D={Dc} T=0.05, fwd+bwd incl. id sort + bf16 conversion", "bound": "tensor",
                "ms": ms_b, "best_ms": best_b, "ms_fwd": ms_f, "achieved": flops / ms_b / 1e9, "peak": peaks["bf16_tflops"],
                "peak_sustained": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": flops / ms_b / 1e9 / peaks["bf16_tflops"], "frac_of_sustained": flops / ms_b / 1e9 / peaks["bf16_tflops_sustained"],
                "alg_flops": flops, "loss": float(res["l"])})
    if not quick:   # the exact fp32 SIMT path at a quarter of the rows, for comparison
        Bs = 16384
        u2, it2 = u[:Bs].detach().requires_grad_(True), it[:Bs].detach().requires_grad_(True)
        res2 = {}

        def ce32():
            res2["l"] = ops.fused_inbatch_ce(u2, it2, item_ids[:Bs], None, pool.detach(), 0.05)[0]
        ms32, _ = time_op(lambda: (ce32(), res2["l"].backward()), 2, flush)
        f32 = 6.0 * Bs * (Bs + Hc) * Dc
        out.append({"kernel": "ce_fwd_tiles + ce_bwd_pass (exact fp32 SIMT path)", "workload": f"B={Bs} H={Hc} D={Dc} fwd+bwd",
                    "bound": "tensor", "ms": ms32, "achieved": f32 / ms32 / 1e9, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": f32 / ms32 / 1e9 / peaks["bf16_tflops"], "alg_flops": f32})
        del u2, it2
    del u, it, pool
    torch.cuda.empty_cache()
    # --- C5 scoring + top-100, tcgen05 filter + exact re-rank: one GPU's shard of the 10M-item corpus
    Q, N, K = (2048, 200_000, 100) if quick else (32768, 1_250_000, 100)
    q = torch.nn.functional.normalize(torch.randn(Q, 128, device=dev), dim=1)
    e = torch.nn.functional.normalize(torch.randn(N, 128, device=dev), dim=1)
    prep = ops.PreparedCorpus(e)
    ms, best = time_op(lambda: ops.score_topk(q, e, K, precision="bf16", prepared=prep), 3, flush)
    flops = 2.0 * Q * N * 128
    out.append({"kernel": "topk_tc_kernel + topk_tc_stage2 (tcgen05 bf16 filter, fp64 exact re-rank, bit-exact rows)",
                "workload": f"C5 shard: Q={Q} N={N} D=128 K={K}", "bound": "tensor", "ms": ms, "best_ms": best,
                "achieved": flops / ms / 1e9, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": flops / ms / 1e9 / peaks["bf16_tflops"], "queries_per_s": Q / ms * 1e3, "alg_flops": flops,
                "fp32_fallback_queries": ops.topk_stats["unverified"]})
    if not quick:
        Qs = 2048
        ms32, _ = time_op(lambda: ops.score_topk(q[:Qs], e, K), 1, flush)
        out.append({"kernel": "topk_stage1 + topk_stage2 (exact fp32 SIMT path)", "workload": f"Q={Qs} N={N} D=128 K={K}",
                    "bound": "tensor", "ms": ms32, "achieved": 2.0 * Qs * N * 128 / ms32 / 1e9, "peak": peaks["bf16_tflops"],
                    "unit": "TFLOP/s", "frac": 2.0 * Qs * N * 128 / ms32 / 1e9 / peaks["bf16_tflops"],
                    "queries_per_s": Qs / ms32 * 1e3})
    return out


# ----------------------------------------------------------------------------- extra workloads: C3, C5
def _clock_wrap(rank, local_rank):
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    return sampler


def build_c3(rank, world, dev, B_global, dropout=0.1, zipf=False, users_per_gpu=12_500_000, v_item=10_000_001, L=200,
             tf32=True):
    """BASELINE configs[2] as ONE model: 8 sparse features (D=128), row-sharded big tables, MLP [256,128]->128 towers,
    in-batch softmax over the global batch.  The user table has `users_per_gpu` rows per rank (100M at 8 GPUs: the
    whole table + fp32 Adam moments is 154 GB and does not fit one GPU); everything else is the config's size at any N."""
    import recommendsystemproject_b200 as tt
    from recommendsystemproject_b200 import synth
    from recommendsystemproject_b200.dist import ShardedTrainStep
    v_user = users_per_gpu * world + 1
    cfg = synth.config_c3(v_user=v_user, v_item=v_item, dropout=dropout, shard=True, world=world, rank=rank)
    torch.manual_seed(0)
    with torch.device(dev):       # shards are born on the GPU (a 51 GB table never exists on the host)
        model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), *synth.MAPS_C3)
    model = model.to(dev).train()
    for m in model.modules():
        if hasattr(m, "gather_on_save"):
            m.gather_on_save = False
    opt = tt.FusedTwoTowerOptimizer(model, lr=5e-4, max_grad_norm=1.0, table_mode="sparse")
    B = B_global // world
    host = [synth.make_batch_c3(B=B, L=L, v_user=v_user, v_item=v_item, seed=300 + 17 * rank + s, zipf=zipf) for s in range(2)]
    host = [tree_to(h, None, pin=True) for h in host]
    dev_batch = tree_to(host[0], dev)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    step = ShardedTrainStep(model, opt, dev_batch, 0.05, restore_tables=False)
    return model, opt, step, host, dev_batch, cfg, v_user


def run_c3(args, rank, local_rank, world, dev, peaks, as_dict=False):
    """BASELINE configs[2], integrated: gather (row-sharded) -> towers -> global in-batch softmax -> backward -> global-norm
    clip -> dense Adam + row-wise Adam, global batch 65536 at every N (strong scaling in the batch)."""
    import torch.distributed as dist
    from recommendsystemproject_b200 import ops
    B_global = args.c3_batch
    model, opt, step, host, dev_batch, cfg, v_user = build_c3(rank, world, dev, B_global, zipf=args.zipf)
    B = B_global // world
    L, D = 200, 128
    n_valid = int((host[0]["user_tower"]["sequence"]["hist_item_ids"] != 0).sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    step.check_flags()
    sampler = _clock_wrap(rank, local_rank)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    dev_ms = a.elapsed_time(b)
    # e2e: pinned host batch -> H2D (copy stream, one step ahead) -> step -> D2H loss, all inside the timed region
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    staging = [tree_to(host[0], dev), tree_to(host[1], dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(s):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s % 2])
            _copy_into(staging[s % 2], host[s % 2])
            ready[s % 2].record(copy_stream)
    for e in consumed:
        e.record()
    barrier()
    a.record()
    prefetch(0)
    for s in range(args.steps):
        if s + 1 < args.steps:
            prefetch(s + 1)
        torch.cuda.current_stream().wait_event(ready[s % 2])
        step.load_batch(staging[s % 2])          # device-to-device into the graph's static buffers
        consumed[s % 2].record()
        loss_host.copy_(step())
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None
    step.check_flags()
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    # per-kernel device times of one step (CUDA events around eager phases would perturb the graph: use the profiler-free
    # kernel section below instead); dominant kernel + roofline come from kernel_rooflines at this rank's shapes
    if rank != 0:
        return None
    ms = dev_ms / args.steps
    h2d = tree_bytes(host[0])
    grp = model.shard_group
    line = {"metric": "train samples/sec (fwd+bwd+clip+Adam), integrated C3 step", "value": B_global * args.steps / (dev_ms / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 tables + TF32 tower GEMMs + bf16 tcgen05 loss",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: 8 sparse features D=128 (user_id {v_user - 1} rows = 12.5M per GPU, u_cat1 1e5, u_cat2 1e3, "
                                   f"u_cat3 32, hist_item_ids 10M x L=200 ragged mean-pooled; item_id 10M, i_cat 1e4, i_year 152), MLP [256,128]->128, "
                                   f"in-batch softmax over the global batch {B_global}, dropout 0.1; tables >= 1 MB row-sharded owner=row%W "
                                   f"({len(grp.tables)} tables), towers data-parallel with global BatchNorm statistics",
                       "global_batch": B_global, "per_gpu_batch": B, "parallelism": f"row-sharded tables x{world} + dp{world}",
                       "ids": "zipf(1.05)" if args.zipf else "uniform (every looked-up row distinct: HBM worst case)",
                       "l2": "working set (tables + Adam state, > 40 GB per GPU) far exceeds L2", "peaks": peaks["source"],
                       "step": "one CUDA graph per rank, NCCL collectives inside"},
            "e2e": {"value": B_global * args.steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
                    "note": "pinned host batch -> H2D on a copy stream one step ahead -> D2D into the graph's buffers -> step -> loss read-back"},
            "gpu_launches": (step.launches_per_step or 0) * args.steps * 2, "gpu_launches_per_step": step.launches_per_step,
            "nvlink_bytes_per_step_per_gpu": getattr(step, "a2a_bytes_per_step", 0), "hist_valid_positions_per_gpu": n_valid,
            "clocks": clocks, "final_loss": float(loss_host)}
    if as_dict:
        return line
    print(json.dumps(line), flush=True)
    return line


def _copy_into(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k])
    elif isinstance(dst, list):
        for x, y in zip(dst, src):
            _copy_into(x, y)


def ops_count():
    from recommendsystemproject_b200 import ops
    return ops.launch_counter["calls"]


def run_c5(args, rank, local_rank, world, dev, peaks):
    """BASELINE configs[4]: top-100 over a 10M-item corpus sharded over the GPUs (each GPU: tcgen05 filter + exact
    re-rank over its shard), all-gather of the [Q, K] lists, global merge.  Q = 16384 queries per step."""
    import torch.distributed as dist
    from recommendsystemproject_b200 import dist as tdist, ops
    N_total, Q, K, D = 10_000_000, 16384, 100, 128
    bounds = [N_total * r // world for r in range(world + 1)]
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    shard = torch.nn.functional.normalize(torch.randn(bounds[rank + 1] - bounds[rank], D, device=dev, generator=gen), dim=1)
    prep = ops.PreparedCorpus(shard)
    qgen = torch.Generator().manual_seed(6)
    host_q = torch.nn.functional.normalize(torch.randn(Q, D, generator=qgen), dim=1).pin_memory()
    q_dev = host_q.to(dev)

    def step(q):
        return tdist.sharded_topk(q, shard, K, rank, world, bounds[:-1],
                                  topk_fn=lambda a, e, k, off: ops.score_topk(a, e, k, off, precision="bf16", prepared=prep))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step(q_dev)
    sampler = _clock_wrap(rank, local_rank)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step(q_dev)
    b.record()
    barrier()
    dev_ms = a.elapsed_time(b)
    out_host = torch.empty(Q, K, dtype=torch.int64).pin_memory()
    barrier()
    a.record()
    for _ in range(args.steps):
        _, idx = step(host_q.to(dev, non_blocking=True))
        out_host.copy_(idx)
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    flops = 2.0 * Q * N_total * D / world      # per GPU per step
    ms = dev_ms / args.steps
    line = {"metric": "corpus top-K queries/sec (top-100, 10M-item corpus)", "value": Q * args.steps / (dev_ms / 1e3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16 filter / f64 exact re-rank", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: 10M x 128 corpus in {world} shard(s), Q=16384 queries per step, K=100, "
                                   "bit-exact rows (score desc, row asc)", "parallelism": f"corpus-sharded x{world}",
                       "l2": "corpus shard (>= 320 MB bf16) exceeds L2", "peaks": peaks["source"]},
            "e2e": {"value": Q * args.steps / (e2e_ms / 1e3), "unit": "queries/s", "h2d_bytes_per_step": Q * D * 4,
                    "d2h_bytes_per_step": Q * K * 8, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": ops_count(),
            "roofline": {"bound": "tensor", "kernel": "topk_tc_kernel (sampling pass + full pass) + topk_tc_stage2 + topk_merge",
                         "achieved": flops / ms / 1e9, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": flops / ms / 1e9 / peaks["bf16_tflops"], "traffic": None, "alg_flops": flops},
            "cpu_baseline": None, "clocks": clocks, "resampled_queries": ops.topk_stats.get("resampled"),
            "fp32_fallback_queries": ops.topk_stats.get("unverified")}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- main GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline section")
    ap.add_argument("--quick", action="store_true", help="smaller kernel-roofline shapes")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--c3-batch", type=int, default=65536, help="global batch of the integrated C3 step")
    ap.add_argument("--zipf", action="store_true", help="Zipf(1.05) item ids instead of uniform")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist
    import recommendsystemproject_b200 as tt
    from recommendsystemproject_b200 import _lib, ops
    torch.cuda.set_device(local_rank)
    _lib.require_device()
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    if args.workload in ("c3", "c5"):
        (run_c3 if args.workload == "c3" else run_c5)(args, rank, local_rank, world, dev, peaks)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    wl = workload(args.workload)

    torch.manual_seed(0)
    model = tt.TwoTowerModel(tt.GenericTower(wl["cfg"], "user_tower"), tt.GenericTower(wl["cfg"], "item_tower"),
                             *wl["maps"]).to(dev).train()
    opt = tt.FusedTwoTowerOptimizer(model, lr=wl["lr"], max_grad_norm=1.0, table_mode="dense")
    host_batches = [tree_to(wl["batch_fn"](100 + rank * 17 + s), None, pin=True) for s in range(4)]
    dev_batch = tree_to(host_batches[0], dev)
    torch.cuda.synchronize()
    if world > 1:
        from recommendsystemproject_b200.dist import DataParallelStep
        step = DataParallelStep(model, opt, dev_batch, wl["T"])
    else:
        step = tt.GraphedTrainStep(model, opt, dev_batch, wl["T"])

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.fill_(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    calls0 = ops.launch_counter["calls"]
    step.count_launches = True
    for _ in range(args.warmup):
        step()
    launches_per_step = getattr(step, "launches_per_step", None)

    # ---- value: device-resident batch, per-step CUDA events, L2 flushed (untimed) between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    evs = []
    for _ in range(args.steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        evs.append((a, b))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- e2e: pinned host batch -> H2D -> step -> D2H loss, every step, one timed region
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(args.steps):
        loss = step(host_batches[s % len(host_batches)])
        loss_host.copy_(loss, non_blocking=False)
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    final_loss = float(loss_host)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    B = wl["B"]
    value = world * B * args.steps / (dev_ms / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    h2d = tree_bytes(host_batches[0])

    # ---- secondary number: the same step with TF32 tensor-core GEMMs in the towers (BASELINE quotes this config in
    # bf16; TF32 keeps fp32 range and 10 mantissa bits).  The headline `value` above stays fp32, the precision every
    # parity test checks; tolerance of the TF32 step (loss 1e-3, gradient 2.1e-2 relative): tests/test_gpu_model.py::test_tf32_tower_step_within_tolerance.
    tf32 = None
    if world == 1:
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
            torch.manual_seed(0)
            model_t = tt.TwoTowerModel(tt.GenericTower(wl["cfg"], "user_tower"), tt.GenericTower(wl["cfg"], "item_tower"),
                                       *wl["maps"]).to(dev).train()
            opt_t = tt.FusedTwoTowerOptimizer(model_t, lr=wl["lr"], max_grad_norm=1.0, table_mode="dense")
            step_t = tt.GraphedTrainStep(model_t, opt_t, dev_batch, wl["T"])
            for _ in range(args.warmup):
                step_t()
            torch.cuda.synchronize()
            evs_t = []
            for _ in range(args.steps):
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step_t()
                b.record()
                evs_t.append((a, b))
            torch.cuda.synchronize()
            ms_t = sum(a.elapsed_time(b) for a, b in evs_t) / args.steps
            tf32 = {"value": B * 1e3 / ms_t, "unit": "samples/s", "ms_per_step": ms_t,
                    "note": "tower / encoder GEMMs in TF32 (torch.backends.cuda.matmul.allow_tf32), everything else as in `value`"}
            del step_t, opt_t, model_t
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False

    kernels = []
    if not args.no_kernels:
        try:
            kernels = kernel_rooflines(peaks, flush, quick=args.quick)
        except torch.cuda.OutOfMemoryError as ex:  # report, never hide
            kernels = [{"error": f"kernel roofline section skipped: {ex}"}]
    # dominant hand-written kernel of the C2 step (see profiles/r1_c2_step_launches_fused.md); its standalone roofline
    roof = step_roofline(model, wl, dev, peaks, flush, dev_batch)

    cpu = None
    try:
        times = cpu_port_steps(wl, args.cpu_steps, 1)
        cpu = {"value": B * len(times) / sum(times), "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{len(times)} full steps of B={B} (oracle/twotower_oracle.py, torch fp32, all host threads)"}
    except Exception as ex:  # noqa: BLE001
        cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": "train samples/sec (fwd+bwd+clip+Adam)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "tf32_towers": tf32,
        "config": {"workload": wl["desc"], "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "flushed between steps by an untimed 256 MiB write; step = one CUDA-graph replay",
                   "table_mode": "dense Adam on every table row (reference semantics)", "peaks": peaks["source"]},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": (launches_per_step or 0) * args.steps * 2,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks, "final_loss": final_loss,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def step_roofline(model, wl, dev, peaks, flush, batch):
    """Roofline of the dominant hand-written kernel inside the C2 step (profiles/r1_c2_step_launches_fused.md): the
    one-pass Linear weight + bias gradient (linear_wgrad_partial + linear_wgrad_reduce, 16 Linear layers per step).
    Its shapes are recorded from one eager backward of this very model and batch, then every call is timed standalone
    with CUDA events (L2 flushed before each).  HBM-bound by construction (each operand is read once):
    algorithmic bytes = (rows * (n_out + n_in) + n_out * n_in + n_out) * 4 per call."""
    from recommendsystemproject_b200 import ops
    import ctypes
    ops.wgrad_shapes = []
    try:
        model.zero_grad(set_to_none=False)
        u, i, hn = model(batch)
        ids = batch["item_tower"]["sparse"][:, 0]
        model.compute_loss(u, i, item_ids=ids, hard_neg_emb=hn, temperature=wl["T"]).backward()
        torch.cuda.synchronize()
        shapes = list(ops.wgrad_shapes)
    finally:
        ops.wgrad_shapes = None
    lib = ops._lib.load()
    total_ms, total_bytes, total_flops = 0.0, 0, 0.0
    for rows, n_out, n_in in shapes:
        g = torch.randn(rows, n_out, device=dev)
        x = torch.randn(rows, n_in, device=dev)
        gw, gb = torch.empty(n_out, n_in, device=dev), torch.empty(n_out, device=dev)
        nb = ctypes.c_size_t(0)
        ops.check(lib.tt_linear_wgrad_workspace(rows, n_out, n_in, ctypes.byref(nb)), "tt_linear_wgrad_workspace")
        ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)

        def call():
            ops.check(lib.tt_linear_wgrad(ops._p(g), ops._p(x), rows, n_out, n_in, ops._p(gw), ops._p(gb), 0, ops._p(ws),
                                          ws.numel(), ops._stream()), "tt_linear_wgrad")
        ms, _ = time_op(call, 5, flush)
        total_ms += ms
        total_bytes += (rows * (n_out + n_in) + n_out * n_in + n_out) * 4
        total_flops += 2.0 * rows * n_out * n_in
    gbs = total_bytes / total_ms / 1e6
    return {"bound": "hbm", "kernel": f"linear_wgrad_partial + linear_wgrad_reduce (weight + bias gradient of the step's {len(shapes)} Linear layers)",
            "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
            "ms": total_ms, "launches": 2 * len(shapes), "alg_bytes": total_bytes, "alg_flops": total_flops,
            "shapes": sorted(set(shapes)),
            "note": "B=512 step: every kernel of it is launch/latency-bound (6-12 us per launch for ~10 MB of operands); "
                    "see `kernels` for the hot-path kernels at BASELINE-scale shapes"}


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner under NCCL_DEBUG, torchrun helpers)
    # write to fd 1 too: send fd 1 to stderr while the benchmark runs and hand the real stdout back to print().
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w", buffering=1)
    main()
