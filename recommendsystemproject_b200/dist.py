"""Multi-GPU pieces (one process per GPU, torch.distributed / NCCL over NVLink + own peer-memory kernels).

The reference is single-process (no collective of any
kind), so everything here is new:
  * ShardedTrainStep    -- THE integrated step: row-sharded tables (one batched exchange per step, sharded.py),
                          data-parallel towers with BatchNorm statistics over the global batch, in-batch softmax over
                          the global batch (rectangular tcgen05 CE, false-negative mask across ranks), SUM all-reduce of
                          the flat gradient buffer, dense Adam + owner-side row-wise Adam; one CUDA graph per rank.
  * global_inbatch_ce   -- the loss of that step: all-gather of item embeddings (+ ids) with a gradient-carrying
                          backward (reduce-scatter).
  * sharded_topk        -- corpus-sharded top-K: per-GPU top-K, ONE all-to-all of the [Q/W, K] candidate slices,
                          per-rank merge.
  * DataParallelStep    -- round 1's data parallelism (per-rank in-batch negatives, gradients averaged); still used for
                          batches with per-row hard-negative slabs.
  * ShardedEmbeddingBag -- round 1's single-table bag (kept for its bitwise exchange="rows" mode); ShardedTableGroup
                          in sharded.py supersedes it (compacted per-owner lists, all tables in one exchange).
Host-side logic is plain index arithmetic on tensors of either device so that it can be exercised with gloo on CPU
(tests/test_dist_cpu.py, tests/test_sharded_cpu.py); the kernels it feeds are CUDA-only.
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
        self.opt.sync_lr()
        self.g_fb.replay()
        self._reduce()
        self.g_opt.replay()
        return self.static_loss


SYNC_BN = True      # diagnosis switch (bench.py TT_C3_LOCAL_BN): False = per-rank BatchNorm statistics
P2P_BN = True       # BatchNorm exchanges over NVLink peer memory (ops.P2PSmallGather); False = NCCL all-gathers


class ShardedTrainStep:
    """The integrated multi-GPU training step, one CUDA graph per rank, collectives included:

        zero_grad
        model(batch)        row-sharded tables: ONE batched exchange for every sharded feature of both towers
                            (sharded.ShardedTableGroup: route -> all-to-all -> owner gather/pool -> all-to-all -> combine);
                            towers data-parallel, BatchNorm statistics over the GLOBAL batch (ops.batch_norm_act: all-gather
                            of per-channel (mean, M2, n) forward, all-reduce of the two gradient sums backward)
        loss                in-batch softmax over the GLOBAL batch: item embeddings (+ item ids) all-gathered, every rank
                            computes its B/W x B slab with the fused CE kernel (never in HBM), gradients of the gathered
                            items flow back to their owners (sum)
        backward            table gradients travel to the owning rank (all-to-all) and are segment-reduced there
        all-reduce (SUM)    of the flat dense-gradient buffer; its trailing float carries the row-sharded tables' local
                            sum of squares, so the global gradient norm costs no extra collective
        clip + Adam         dense parameters (replicated, bit-identical on every rank) + row-wise Adam on the local shards

    Scaling of the loss: rank r back-propagates mean_r / W, so every SUM above yields the gradient of the global-batch
    mean -- what one process running the reference step on all W*B samples computes (TwoTowerModel.py:95-140,
    GenericTower.py:234, training_utils.py:51-56).  Per-rank batches must have equal size.
    Works at W = 1 (no collectives) and without sharded tables (then it is DataParallelStep with global-batch semantics)."""

    def __init__(self, model, optimizer: FusedTwoTowerOptimizer, example_batch, temperature, item_id_col=0, warmup=3,
                 graph=True, loss_precision="auto", global_loss=True, restore_tables=True):
        self.model, self.opt, self.temperature, self.item_id_col = model, optimizer, temperature, item_id_col
        self.restore_tables = restore_tables
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.loss_precision, self.global_loss = loss_precision, global_loss
        if batch_has_hard_negatives(example_batch) and global_loss and self.world > 1:
            raise ops.TTError("ShardedTrainStep: per-row hard negatives with a global in-batch loss are not supported")
        self.static_batch = _clone_tree(example_batch)
        # the in-batch ids are column `item_id_col` of the item tower's sparse block, i.e. indices into that feature's
        # table: the loss kernel sorts only the bits such an index can have (an id outside the table raises the
        # feature's out-of-range flag in the gather AND bit 3 of the loss flags)
        feats = getattr(model.item_tower, "sparse_features", None) or []
        self.id_bits = ops.id_bits_for(feats[item_id_col]["vocab_size"]) if item_id_col < len(feats) else 64
        self.loss_flags = torch.zeros(1, dtype=torch.int32, device=optimizer.flat_p.device)
        # tensor-core loss: forward + dU in one walk over the logit tiles (ops.FusedInBatchCE single_pass) when the
        # temperature keeps L2-normalised embeddings inside that kernel's range; checked on the device, see below
        self.single_pass = ops.single_pass_ok(temperature) and not batch_has_hard_negatives(example_batch)
        if self.world > 1 and optimizer.sparse_tables:
            raise ops.TTError("ShardedTrainStep on several ranks: replicated tables need table_mode='dense' (their gradients "
                              "ride the dense all-reduce); only row-sharded tables are updated touched-rows-only")
        if self.world > 1:
            ops.bn_sync.world, ops.bn_sync.rank, ops.bn_sync.group = self.world, self.rank, None
            if SYNC_BN and P2P_BN and ops.bn_sync.p2p is None and optimizer.flat_p.is_cuda:
                # BatchNorm vectors cross ranks through NVLink peer memory (one kernel, one round trip); NCCL if the
                # ranks cannot map each other's memory (every rank must take the same branch: agree on it)
                ok = torch.ones(1, device=optimizer.flat_p.device)
                try:
                    p2p = ops.P2PSmallGather(self.rank, self.world, optimizer.flat_p.device)
                except Exception as ex:  # noqa: BLE001
                    p2p, ok = None, torch.zeros(1, device=optimizer.flat_p.device)
                    self.p2p_error = repr(ex)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                ops.bn_sync.p2p = p2p if float(ok) > 0 else None
            k = 0
            for m in model.modules():
                if isinstance(m, torch.nn.BatchNorm1d):
                    m._tt_sync = k if SYNC_BN else False      # call-site key of this layer's exchanges
                    k += 1
            # replicas start identical: dense parameters, buffers (BatchNorm statistics, dropout seeds, pad rows)
            dist.broadcast(optimizer.flat_p, src=0)
            for b in model.buffers():
                dist.broadcast(b, src=0)
        self.bn_exchange = ("none" if self.world == 1 or not SYNC_BN else
                            ("nvlink-p2p kernel (symmetric memory)" if ops.bn_sync.p2p is not None else
                             "nccl all-gather" + (f" (p2p unavailable: {getattr(self, 'p2p_error', 'disabled')})")))
        optimizer._norm_staged_by_caller = True
        snap = self._snapshot()
        side = torch.cuda.Stream()
        for attempt in range(2):
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._step_eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore(snap)
            # embeddings outside the single-pass loss kernel's range (towers that do not normalise): every rank falls
            # back to the three-pass kernels together and warms up again
            bad = (self.loss_flags & ops.CE_FLAG_LOGIT_RANGE).clamp(max=1).float()
            if self.world > 1:
                dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if not (self.single_pass and float(bad) > 0):
                break
            if not self.restore_tables and getattr(model, "shard_group", None) is not None and model.shard_group.tables:
                raise ops.TTError("ShardedTrainStep: the towers' embeddings are outside the single-pass loss kernel's range and the "
                                  "warm-up steps already updated the row-sharded tables (restore_tables=False); rebuild the model "
                                  "and pass single_pass-incompatible towers through TT_CE_SINGLE_PASS=0")
            self.single_pass = False
            self.loss_flags.zero_()
        self.graph = None
        self.launches_per_step = None
        if graph:
            grp = getattr(model, "shard_group", None)
            if grp is not None:
                grp.a2a_bytes = 0
            self.graph = torch.cuda.CUDAGraph()
            c0 = ops.launch_counter["calls"]
            with torch.cuda.graph(self.graph):
                self.static_loss = self._step_eager()
            self.launches_per_step = ops.launch_counter["calls"] - c0
            self.a2a_bytes_per_step = grp.a2a_bytes if grp is not None else 0
            torch.cuda.synchronize()

    def _snapshot(self):
        """Warm-up steps are real optimizer steps: snapshot and restore so construction has no side effect.  The shards of
        row-sharded tables are included only when `restore_tables` (default; switch off for tables of tens of GB)."""
        o = self.opt
        grp = getattr(self.model, "shard_group", None)
        shard_w = {id(t.weight) for t in grp.tables.values()} if grp else set()
        named = list(self.model.named_parameters()) + list(self.model.named_buffers())
        state = [(v, v.detach().clone()) for _, v in named if id(v) not in shard_w]
        shard = {}
        if grp and self.restore_tables:
            shard = {n: (t.weight.detach().clone(), t.exp_avg.clone(), t.exp_avg_sq.clone()) for n, t in grp.tables.items()}
        tables = {k: (m.clone(), v.clone()) for k, (m, v) in o.table_state.items()}
        return state, o.flat_m.clone(), o.flat_v.clone(), o.step_dev.clone(), shard, tables

    def _restore(self, snap):
        state, m, v, step, shard, tables = snap
        o = self.opt
        for k, (tm, tv) in tables.items():
            o.table_state[k][0].copy_(tm)
            o.table_state[k][1].copy_(tv)
        with torch.no_grad():
            for dst, src in state:
                dst.copy_(src)
        o.flat_m.copy_(m); o.flat_v.copy_(v); o.step_dev.copy_(step)
        grp = getattr(self.model, "shard_group", None)
        for n, (w, ea, es) in shard.items():
            t = grp.tables[n]
            with torch.no_grad():
                t.weight.copy_(w)
            t.exp_avg.copy_(ea); t.exp_avg_sq.copy_(es)

    def _loss(self, u, i, ids):
        if self.world > 1 and self.global_loss:
            return global_inbatch_ce(u, i, ids, None, self.temperature, precision=self._precision(u),
                                     nan_flags=self.loss_flags, id_bits=self.id_bits, single_pass=self.single_pass)
        return ops.fused_inbatch_ce(u, i, ids, None, None, self.temperature, nan_flags=self.loss_flags,
                                    precision=self._precision(u), id_bits=self.id_bits, single_pass=self.single_pass)[0]

    def _precision(self, u):
        if self.loss_precision != "auto":
            return self.loss_precision
        return "bf16" if (u.shape[1] in (64, 128) and u.shape[0] * self.world >= 4096) else "fp32"

    def _step_eager(self):
        nvtx = torch.cuda.nvtx
        o = self.opt
        o.zero_grad()
        nvtx.range_push("tt.forward (sharded exchange + towers)")
        u, i, hn = self.model(self.static_batch)
        nvtx.range_pop()
        ids = self.static_batch["item_tower"]["sparse"][:, self.item_id_col]
        nvtx.range_push("tt.loss (global in-batch softmax CE)")
        if hn is not None:
            loss = self.model.compute_loss(u, i, item_ids=ids, hard_neg_emb=hn, temperature=self.temperature)
        else:
            loss = self._loss(u, i, ids)
        nvtx.range_pop()
        nvtx.range_push("tt.backward (CE, towers, gradient exchange, segment reduce)")
        (loss / self.world if self.world > 1 else loss).backward()
        nvtx.range_pop()
        nvtx.range_push("tt.optimizer (all-reduce, clip, Adam, row-wise Adam)")
        o.stage_sharded_norm()
        if self.world > 1:
            dist.all_reduce(o.flat_g_ext)
        o.step()
        nvtx.range_pop()
        out = loss.detach().clone()
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.AVG)      # the global-batch mean, for reporting
        return out

    def load_batch(self, batch, non_blocking=True):
        _copy_tree(self.static_batch, batch, non_blocking)

    def __call__(self, batch=None):
        if batch is not None:
            self.load_batch(batch)
        self.opt.sync_lr()
        if self.graph is None:
            return self._step_eager()
        self.graph.replay()
        return self.static_loss

    def check_flags(self):
        """One host read of the step's device flags (call it between steps, never inside the capture): NaN embeddings
        raise the reference's RuntimeError texts (TwoTowerModel.py:84-92), an item id outside its table IndexError."""
        flags = int(self.loss_flags.item()) if self.loss_flags.is_cuda else 0
        if flags:
            self.loss_flags.zero_()
        if flags & ops.CE_FLAG_LOGIT_RANGE:
            raise RuntimeError("in-batch loss: |logit| exceeds the single-pass tensor-core kernel's range; "
                               "build ShardedTrainStep and set .single_pass = False before capture")
        if flags & 8:
            raise IndexError(f"item id outside [0, 2**{self.id_bits}) in the in-batch loss (ids must index the item table)")
        if flags & 1:
            raise RuntimeError("Found NaN in User Embedding")
        if flags & 2:
            raise RuntimeError("Found NaN in Item Embedding")
        grp = getattr(self.model, "shard_group", None)
        if grp is not None:
            grp.check_flags()


def batch_has_hard_negatives(batch) -> bool:
    return bool(batch.get("hard_negatives"))


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
                 shard_offsets: List[int], topk_fn=None, merge_fn=None, gather_result: bool = True):
    """Corpus-sharded top-K (SURVEY 8e "Retrieval"): every GPU scores ALL queries against its corpus shard, then the
    [Q, K] candidate lists are exchanged with ONE all-to-all so that rank r receives the W lists of ITS Q/W queries and
    merges only those ((score desc, global row asc) tie-break, rows already carry the shard offset).
    gather_result=True: the merged slices are all-gathered and every rank returns the full [Q, K] result;
    False: returns (scores, rows) of this rank's query slice [ceil(Q/W), K] (its rows [r*ceil(Q/W), ...) of the batch;
    slices past Q are padding)."""
    topk_fn = topk_fn or ops.score_topk
    merge_fn = merge_fn or ops.topk_merge
    s, i = topk_fn(query, local_corpus, k, shard_offsets[rank])
    if world == 1:
        return s, i
    Q = s.shape[0]
    per = (Q + world - 1) // world
    if per * world != Q:                                    # pad the query dimension to a multiple of W
        s = torch.cat([s, s.new_full((per * world - Q, k), float("-inf"))])
        i = torch.cat([i, i.new_full((per * world - Q, k), -1)])
    rs, ri = torch.empty_like(s), torch.empty_like(i)
    dist.all_to_all_single(rs, s.contiguous())              # block w of rs = shard w's list for my query slice
    dist.all_to_all_single(ri, i.contiguous())
    ms, mi = merge_fn(rs.view(world, per, k), ri.view(world, per, k))
    if not gather_result:
        return ms, mi
    fs = torch.empty(world * per, k, dtype=ms.dtype, device=ms.device)
    fi = torch.empty(world * per, k, dtype=mi.dtype, device=mi.device)
    dist.all_gather_into_tensor(fs, ms.contiguous())
    dist.all_gather_into_tensor(fi, mi.contiguous())
    return fs[:Q], fi[:Q]


# ---------------------------------------------------------------------------
# row-sharded embedding table with pooled lookup, backward and fused row-wise Adam
# ---------------------------------------------------------------------------
class _ShardedOps:
    """The three device ops ShardedEmbeddingBag needs.  Production = the CUDA kernels of libtt_b200; the CPU gloo
    tests inject oracle implementations (tests may use oracle/, the product may not)."""

    @staticmethod
    def gather(table, rows):                      # [n] -> [n, D]
        if rows.numel() == 0:
            return table.new_empty(0, table.shape[1], dtype=torch.float32)
        return ops.gather_rows(table, rows.reshape(-1), None, None)

    @staticmethod
    def pool(table, idx, mode):                   # table [n+1, D] (last row = pad row), idx [B, L] -> [B, D]
        out = torch.empty(idx.shape[0], table.shape[1], dtype=torch.float32, device=table.device)
        oob = torch.zeros(1, dtype=torch.int32, device=table.device)
        # the pad slot is declared as padding_idx so the kernel counts pads and adds the pad row once, exactly like
        # the unsharded lookup does (bitwise-equal pooling)
        ops.gather_pool_into(table, idx.contiguous(), mode, table.shape[0] - 1, out, None, oob)
        return out

    @staticmethod
    def segment_grad(rows, vocab, grad, sq_norm):  # rows [n], grad [n, D] -> (unique_rows, row_grad, n_unique)
        return ops.segment_grad(rows.reshape(-1, 1), ops.POOL_NONE, None, vocab, grad, None, grad.shape[1], sq_norm)

    @staticmethod
    def pool_sum(table, idx, null_row):           # table [n+1, D] (row null_row = zeros), idx [R, L] -> [R, D] sums
        out = torch.empty(idx.shape[0], table.shape[1], dtype=torch.float32, device=table.device)
        oob = torch.zeros(1, dtype=torch.int32, device=table.device)
        ops.gather_pool_into(table, idx.contiguous(), ops.POOL_SUM, null_row, out, None, oob)
        return out

    @staticmethod
    def segment_grad_pooled(idx, null_row, grad, sq_norm):   # idx [R, L] (null_row = not mine), grad [R, D] per sample
        return ops.segment_grad(idx.contiguous(), ops.POOL_SUM, null_row, null_row, grad, None, grad.shape[1], sq_norm)

    @staticmethod
    def adam(table, m, v, rows, row_grad, n_unique, coef, lr, b1, b2, eps, step_dev):
        ops.rowwise_adam_(table, m, v, rows, row_grad, n_unique, coef, lr, b1, b2, eps, step_dev)


class ShardedEmbeddingBag:
    """Embedding table row-sharded over the ranks (owner = row % W, local row = row // W) with a pooled lookup
    (mean / sum over the L ids of a sample, pads included like GenericTower.py:153-160 does).

    forward   ids [B, L]  ->  all-to-all #1: valid ids to their owners  ->  owner gathers its rows (CUDA gather kernel)
              ->  all-to-all #2: rows back  ->  local gather+pool kernel over (returned rows, pad row)  ->  [B, D].
              The pooling sums the L rows of a sample in position order on the sample's own rank, exactly like the
              unsharded kernel: the result is BITWISE what one GPU holding the whole table computes.
    backward  d pooled [B, D]  ->  per-position gradient rows in owner order  ->  all-to-all #3 to the owners  ->
              deterministic sorted-segment reduction (tt_emb_segment_grad) on the owner  ->  (rows, row_grad, sum g^2)
              kept for step().  No dense [V, D] gradient and no all-reduce of table gradients.
    step      fused row-wise Adam on the touched rows of the local shard; the clip coefficient comes from the global
              gradient norm (all-reduce of one scalar), like the unsharded optimizer.
    The pad row (GenericTower leaves it non-zero and frozen) is replicated on every rank and never routed.

    exchange="pooled" (default for sum / mean): the owners pool.  Every rank sends each owner the [B, L] id matrix with
    the ids that owner does not hold blanked out (equal-sized all-to-all: no counts to exchange, NO host sync), the owner
    runs the fused gather+pool kernel over its shard -> a partial sum per (sample, owner), one [B, D] block travels
    back per peer and the partials are added in rank order (+ pads x pad row, / L for mean).  Backward: all-gather of
    the [B, D] upstream gradients, tt_emb_segment_grad on the owner expands them over the ids it already holds.
    Per GPU and step this moves B*L*8 + 2*B*D*4 bytes per peer instead of the ~n_valid*D*8 bytes of whole rows
    (C3 shapes: 15 MB instead of 867 MB).  The sum is regrouped by owner, so the result equals the single-GPU one
    to rounding (1e-6), not bitwise; exchange="rows" keeps the bitwise path above.
    """

    def __init__(self, vocab: int, dim: int, rank: int, world: int, mode: str = "mean", padding_idx: Optional[int] = 0,
                 device="cuda", seed: int = 0, dev_ops=None, full_weight: Optional[torch.Tensor] = None,
                 exchange: str = "pooled"):
        self.vocab, self.dim, self.rank, self.world = int(vocab), int(dim), int(rank), int(world)
        self.mode = ops.POOL_MODES[mode]
        if self.mode not in (ops.POOL_SUM, ops.POOL_MEAN, ops.POOL_NONE):
            raise ops.TTError("ShardedEmbeddingBag supports mean / sum pooling (and L = 1 lookups)")
        self.padding_idx = padding_idx
        self.ops = dev_ops or _ShardedOps
        if exchange not in ("pooled", "rows"):
            raise ops.TTError(f"unknown exchange '{exchange}' (use 'pooled' or 'rows')")
        self.exchange = exchange
        self.local_rows = (self.vocab - self.rank + self.world - 1) // self.world
        # shard + one all-zero "null" row behind it (index local_rows): what an id held by another rank points at
        # when this rank pools; never touched by the optimizer
        self._weight_ext = torch.zeros(self.local_rows + 1, dim, dtype=torch.float32, device=device)
        self.weight = self._weight_ext[:self.local_rows]
        if full_weight is not None:                                   # tests: shard a given table
            self.weight.copy_(full_weight[self.rank::self.world])
            pad = full_weight[padding_idx].clone() if padding_idx is not None else torch.zeros(dim)
        else:
            gen = torch.Generator(device=device).manual_seed(seed * 1000003 + self.rank)
            bound = (6.0 / (self.vocab + self.dim)) ** 0.5            # xavier_uniform_ over the whole table
            self.weight.copy_((torch.rand(self.local_rows, dim, device=device, generator=gen) * 2 - 1) * bound)
            pad = (torch.rand(dim, generator=torch.Generator().manual_seed(seed)) * 2 - 1) * bound
        self.pad_row = pad.to(device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(self.weight, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros_like(self.weight, dtype=torch.float32)
        self.sq_norm = torch.zeros(1, dtype=torch.float32, device=device)
        self.pending = []
        self.a2a_bytes = 0
        # null row index of every rank's shard (shards differ in length by at most one row)
        self._null_rows = torch.tensor([(self.vocab - r + self.world - 1) // self.world for r in range(self.world)],
                                       dtype=torch.int64, device=device)
        # autograd anchor: the lookup's inputs are integer ids, so something that requires grad must enter the node
        self._anchor = torch.zeros(1, dtype=torch.float32, device=device, requires_grad=True)

    # -- collectives (degenerate to local copies at W = 1)
    def _a2a(self, send, send_counts, recv_counts):
        self.a2a_bytes += send.numel() * send.element_size()
        if self.world == 1:
            return send
        return all_to_all_rows(send, send_counts, recv_counts)

    def _forward_pooled(self, ids: torch.Tensor) -> torch.Tensor:
        B, L = ids.shape
        W = self.world
        is_pad = (ids == self.padding_idx) if self.padding_idx is not None else torch.zeros_like(ids, dtype=torch.bool)
        owner = ids % W
        local = ids // W
        ranks = torch.arange(W, device=ids.device).view(W, 1, 1)
        # [W, B, L]: block w = what owner w pools for my samples (its local row, or its null row)
        send = torch.where((owner.unsqueeze(0) == ranks) & ~is_pad.unsqueeze(0), local.unsqueeze(0),
                           self._null_rows.view(W, 1, 1)).contiguous()
        recv = self._a2a_equal(send)                                             # all-to-all #1 (equal splits)
        partial = self.ops.pool_sum(self._weight_ext, recv.view(W * B, L), self.local_rows)
        back = self._a2a_equal(partial.view(W, B, self.dim))                     # all-to-all #2: [owner, B, D]
        pooled = back[0]
        for w in range(1, W):                                                    # fixed order
            pooled = pooled + back[w]
        if self.padding_idx is not None:
            pooled = pooled + is_pad.sum(1, keepdim=True).to(torch.float32) * self.pad_row
        if self.mode == ops.POOL_MEAN and L > 1:
            pooled = pooled * (1.0 / L)
        return _ShardedPooledFn.apply(pooled, self._anchor, self, recv, L)

    def _backward_pooled(self, grad_pooled, recv, L):
        W = self.world
        scale = 1.0 / L if (self.mode == ops.POOL_MEAN and L > 1) else 1.0
        g = (grad_pooled * scale).contiguous() if scale != 1.0 else grad_pooled.contiguous()
        self.a2a_bytes += g.numel() * g.element_size() * (W - 1)
        if W > 1:
            g_all = torch.empty((W,) + tuple(g.shape), dtype=g.dtype, device=g.device)
            dist.all_gather(list(g_all.unbind(0)), g)                            # every owner needs every rank's rows
        else:
            g_all = g.unsqueeze(0)
        rows, row_grad, n_unique = self.ops.segment_grad_pooled(recv.view(W * g.shape[0], L), self.local_rows,
                                                                g_all.view(W * g.shape[0], self.dim), self.sq_norm)
        self.pending.append((rows, row_grad, n_unique))

    def _a2a_equal(self, send):
        self.a2a_bytes += send.numel() * send.element_size() * (self.world - 1) // self.world
        if self.world == 1:
            return send
        out = torch.empty_like(send)
        dist.all_to_all_single(out, send)
        return out

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        if self.exchange == "pooled":
            return self._forward_pooled(ids)
        B, L = ids.shape
        flat = ids.reshape(-1)
        if self.padding_idx is None:
            valid_pos = torch.arange(flat.numel(), device=flat.device)
        else:
            valid_pos = torch.nonzero(flat != self.padding_idx, as_tuple=False).reshape(-1)
        vids = flat[valid_pos]
        owner = vids % self.world
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.world)
        if self.world > 1:
            recv_counts = exchange_counts(counts)
            sc, rc = counts.tolist(), recv_counts.tolist()          # the one host sync of the step (NCCL split sizes)
        else:
            sc = rc = [int(vids.numel())]
        send_ids = (vids // self.world)[order].contiguous()
        recv_ids = self._a2a(send_ids, sc, rc)                                   # all-to-all #1
        rows = self.ops.gather(self.weight, recv_ids)                            # owner-side gather
        n_valid = int(vids.numel())
        buf = torch.empty(n_valid + 1, self.dim, dtype=torch.float32, device=flat.device)
        back = self._a2a(rows, rc, sc)                                           # all-to-all #2
        buf[:n_valid] = back
        buf[n_valid] = self.pad_row
        idx = torch.full((B * L,), n_valid, dtype=torch.int64, device=flat.device)
        sorted_pos = valid_pos[order]
        idx[sorted_pos] = torch.arange(n_valid, device=flat.device)
        pooled = self.ops.pool(buf, idx.view(B, L), self.mode if L > 1 else ops.POOL_NONE)
        return _ShardedLookupFn.apply(pooled, self._anchor, self, sorted_pos // L, recv_ids, sc, rc, L)

    __call__ = forward

    def _backward(self, grad_pooled, sample_of_sorted, recv_ids, sc, rc, L):
        scale = 1.0 / L if (self.mode == ops.POOL_MEAN and L > 1) else 1.0
        g = (grad_pooled * scale).contiguous() if scale != 1.0 else grad_pooled.contiguous()
        g_send = self.ops.gather(g, sample_of_sorted)                            # per-position rows, owner order
        g_recv = self._a2a(g_send, sc, rc)                                       # all-to-all #3
        if recv_ids.numel() == 0:
            return
        rows, row_grad, n_unique = self.ops.segment_grad(recv_ids, self.local_rows, g_recv, self.sq_norm)
        self.pending.append((rows, row_grad, n_unique))

    def zero_grad(self):
        self.pending = []
        self.sq_norm.zero_()

    def step(self, clip_coef, lr, step_dev, betas=(0.9, 0.999), eps=1e-8):
        for rows, row_grad, n_unique in self.pending:
            self.ops.adam(self.weight, self.exp_avg, self.exp_avg_sq, rows, row_grad, n_unique, clip_coef, lr, betas[0],
                          betas[1], eps, step_dev)
        self.pending = []


class _ShardedLookupFn(torch.autograd.Function):
    """Identity on the pooled vectors in forward; routes their gradient to the owners in backward."""

    @staticmethod
    def forward(ctx, pooled, anchor, bag, sample_of_sorted, recv_ids, sc, rc, L):
        ctx.bag, ctx.sc, ctx.rc, ctx.L = bag, sc, rc, L
        ctx.save_for_backward(sample_of_sorted, recv_ids)
        return pooled.view_as(pooled)

    @staticmethod
    def backward(ctx, grad):
        sample_of_sorted, recv_ids = ctx.saved_tensors
        ctx.bag._backward(grad, sample_of_sorted, recv_ids, ctx.sc, ctx.rc, ctx.L)
        return (None,) * 8


class _ShardedPooledFn(torch.autograd.Function):
    """Identity on the pooled vectors in forward; sends their gradient to every owner in backward."""

    @staticmethod
    def forward(ctx, pooled, anchor, bag, recv, L):
        ctx.bag, ctx.L = bag, L
        ctx.save_for_backward(recv)
        return pooled.view_as(pooled)

    @staticmethod
    def backward(ctx, grad):
        (recv,) = ctx.saved_tensors
        ctx.bag._backward_pooled(grad, recv, ctx.L)
        return (None,) * 5


class _AllGatherWithGrad(torch.autograd.Function):
    """[B, D] -> [W, B, D] (all ranks' blocks); backward: every rank holds a gradient for every block, the owner of a
    block receives their SUM (all-reduce + own slice: gloo has no reduce-scatter; the tensors are a few hundred KB)."""

    @staticmethod
    def forward(ctx, x):
        world = dist.get_world_size()
        out = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        dist.all_gather(list(out.unbind(0)), x.contiguous())
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if g.is_cuda:       # NCCL: reduce-scatter moves half the bytes of an all-reduce (gloo has none: CPU tests all-reduce)
            out = torch.empty_like(g[0])
            dist.reduce_scatter_tensor(out, g)
            return out
        dist.all_reduce(g)
        return g[dist.get_rank()]


def global_inbatch_ce(user, item, item_ids, pool, temperature, precision: str = "fp32", ce_fn=None, nan_flags=None,
                      id_bits: int = 64, single_pass: bool = False):
    """SURVEY 8e, "towers + loss": the in-batch softmax over the GLOBAL batch of W*B items while every rank keeps only
    its own B user rows.  Item embeddings (and item ids) are all-gathered; their gradients flow back to the owners
    through the all-gather's backward (reduce-scatter / sum); with the loss of rank r scaled by 1/W and the dense
    gradients SUM-reduced (ShardedTrainStep) the result is d(global-batch mean loss).
    Returns this rank's mean loss over its B rows (global loss = mean over ranks).

    precision='bf16' (tensor-core kernel): the RECTANGULAR form of the fused kernel -- B user rows against the W*B
    gathered item rows, the positive of user b being row rank*B + b -- with the false-negative mask
    (TwoTowerModel.py:101-104) applied against the item ids of ALL ranks, exactly what one process computes on the
    global batch.  The [B, W*B + H] logit slab never exists in HBM.
    precision='fp32' (exact SIMT kernel, small batches): the other ranks' items enter as extra shared negatives (the
    kernel's pool argument); the mask then covers the rank's own B x B block only -- an item id repeated on ANOTHER
    rank counts as a negative there.
    `ce_fn(user, item, item_ids, pool, temperature)` overrides the CUDA kernel (CPU gloo tests inject the oracle)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        rank = dist.get_rank()
        blocks = _AllGatherWithGrad.apply(item)                      # [W, B, D]
        if precision == "bf16" and ce_fn is None:
            B = item.shape[0]
            ids_all = None
            if item_ids is not None:
                ids_all = torch.empty(world * B, dtype=torch.int64, device=item.device)
                dist.all_gather_into_tensor(ids_all, item_ids.reshape(-1).contiguous().long())
            return ops.fused_inbatch_ce(user, blocks.reshape(world * B, -1), ids_all, None, pool, temperature,
                                        nan_flags=nan_flags, precision="bf16", item_offset=rank * B, id_bits=id_bits,
                                        single_pass=single_pass)[0]
        others = torch.cat([blocks[r] for r in range(world) if r != rank], dim=0)
        pool = others if pool is None else torch.cat([others, pool], dim=0)
    if ce_fn is not None:
        return ce_fn(user, item, item_ids, pool, temperature)
    return ops.fused_inbatch_ce(user, item, item_ids, None, pool, temperature, nan_flags=nan_flags, precision=precision,
                                id_bits=id_bits, single_pass=single_pass)[0]


def global_clip_coef(sq_terms: List[torch.Tensor], max_norm: float = 1.0) -> torch.Tensor:
    """min(1, max_norm / (||g||_2 + 1e-6)) over the gradients of ALL ranks' shards (+ replicated dense terms already
    averaged): one all-reduce of a scalar (training_utils.py:53-54 semantics for row-sharded tables)."""
    total = torch.stack([t.reshape(()) for t in sq_terms]).sum().reshape(1)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    norm = total.sqrt()
    return torch.clamp(max_norm / (norm + 1e-6), max=1.0)
