"""Multi-GPU pieces (one process per GPU, torch.distributed / NCCL over NVLink).

The reference is single-process (no collective of any
kind), so everything here is new:
  * DataParallelStep   -- towers data-parallel: per-rank batch, NCCL all-reduce
                          (mean) of the flat dense-gradient buffer between the
                          backward graph and the optimizer graph; the clip
                          norm is therefore the norm of the AVERAGED gradient,
                          identical on every rank.
  * shard helpers       -- row-sharded embedding routing (owner = row % W) and
                          corpus-sharded top-K + global merge.
Host-side routing logic is plain index arithmetic on tensors of either device
so that it can be exercised with gloo on CPU (tests/test_dist_cpu.py); the
kernels it feeds are CUDA-only.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from .optim import FusedTwoTowerOptimizer, _clone_tree, _copy_tree


class DataParallelStep:
    """forward+backward graph -> all-reduce(flat grads) -> optimizer graph."""

    def __init__(self, model, optimizer: FusedTwoTowerOptimizer, example_batch, temperature, item_id_col=0, warmup=3):
        if optimizer.table_mode != "dense":
            raise ops.TTError("DataParallelStep all-reduces dense gradients; use table_mode='dense' "
                              "(row-sharded sparse tables use ShardedEmbedding instead)")
        self.model, self.opt, self.temperature, self.item_id_col = model, optimizer, temperature, item_id_col
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.static_batch = _clone_tree(example_batch)
        # replicas must start identical
        if self.world > 1:
            dist.broadcast(optimizer.flat_p, src=0)
            for b in model.buffers():
                dist.broadcast(b, src=0)
        snap_model = {k: v.clone() for k, v in model.state_dict().items()}
        snap_opt = (optimizer.flat_m.clone(), optimizer.flat_v.clone(), optimizer.step_dev.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                self._reduce()
                self.opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        model.load_state_dict(snap_model)
        optimizer.flat_m.copy_(snap_opt[0])
        optimizer.flat_v.copy_(snap_opt[1])
        optimizer.step_dev.copy_(snap_opt[2])
        self.g_fb = torch.cuda.CUDAGraph()
        self.g_opt = torch.cuda.CUDAGraph()
        c0 = ops.launch_counter["calls"]
        with torch.cuda.graph(self.g_fb):
            self.static_loss = self._fwd_bwd()
        with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
            self.opt.step()
        self.launches_per_step = ops.launch_counter["calls"] - c0
        torch.cuda.synchronize()

    def _fwd_bwd(self):
        self.opt.zero_grad()
        u, i, hn = self.model(self.static_batch)
        ids = self.static_batch["item_tower"]["sparse"][:, self.item_id_col]
        loss = self.model.compute_loss(u, i, item_ids=ids, hard_neg_emb=hn, temperature=self.temperature)
        loss.backward()
        return loss.detach()

    def _reduce(self):
        if self.world > 1:
            dist.all_reduce(self.opt.flat_g, op=dist.ReduceOp.AVG)

    def load_batch(self, batch, non_blocking=True):
        _copy_tree(self.static_batch, batch, non_blocking)

    def __call__(self, batch=None):
        if batch is not None:
            self.load_batch(batch)
        self.g_fb.replay()
        self._reduce()
        self.g_opt.replay()
        return self.static_loss


# ---------------------------------------------------------------------------
# row-sharded embedding routing (owner = row % W), device-agnostic index logic
# ---------------------------------------------------------------------------
def route_ids(ids: torch.Tensor, world: int):
    """Bucket flat ids by owner rank.  Returns (send_ids sorted by owner holding LOCAL rows = id // W,
    send_counts[W], order) with ids.reshape(-1)[order] being the owner-sorted sequence."""
    flat = ids.reshape(-1)
    owner = flat % world
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    return (flat[order] // world), counts, order


def exchange_counts(send_counts: torch.Tensor) -> torch.Tensor:
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts)
    return recv


def all_to_all_rows(send: torch.Tensor, send_counts: List[int], recv_counts: List[int]) -> torch.Tensor:
    out = send.new_empty((sum(recv_counts),) + tuple(send.shape[1:]))
    dist.all_to_all_single(out, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts)
    return out


def sharded_lookup(local_table_fn, ids: torch.Tensor, world: int) -> torch.Tensor:
    """rows = table[ids] for a table row-sharded as owner = id % W, local row = id // W.
    all-to-all #1 routes ids to owners, `local_table_fn(local_rows) -> [n, D]` gathers on the owner (the CUDA
    gather kernel in production), all-to-all #2 returns the rows, which are un-permuted to id order."""
    send_ids, counts, order = route_ids(ids, world)
    recv_counts = exchange_counts(counts)
    sc, rc = counts.tolist(), recv_counts.tolist()
    local_rows = all_to_all_rows(send_ids, sc, rc)
    vecs = local_table_fn(local_rows)
    back = all_to_all_rows(vecs, rc, sc)
    out = torch.empty_like(back)
    out[order] = back
    return out.reshape(*ids.shape, -1)


def sharded_topk(query: torch.Tensor, local_corpus: torch.Tensor, k: int, rank: int, world: int,
                 shard_offsets: List[int], topk_fn=None, merge_fn=None):
    """Per-GPU top-K over the local corpus shard, all-gather of the [Bq, K] candidate lists, global merge with
    the (score desc, global row asc) tie-break.  Every rank returns the full result."""
    topk_fn = topk_fn or ops.score_topk
    merge_fn = merge_fn or ops.topk_merge
    s, i = topk_fn(query, local_corpus, k, shard_offsets[rank])
    if world == 1:
        return s, i
    ss = [torch.empty_like(s) for _ in range(world)]
    ii = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(ss, s)
    dist.all_gather(ii, i)
    return merge_fn(torch.stack(ss), torch.stack(ii))
