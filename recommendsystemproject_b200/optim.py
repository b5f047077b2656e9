"""clip_grad_norm_ + Adam for the two-tower model, fused and sync-free.

Mirrors what the reference loop does per step (training_utils.py:53-56 with
train_twotower.py:111: ``clip_grad_norm_(model.parameters(), 1.0)`` then a
dense ``optim.Adam`` over EVERY parameter, embedding tables included) but
keeps every scalar (global norm, clip coefficient, Adam step count) on the
device so the whole step can be captured in one CUDA graph.

Embedding tables run in one of two modes (SURVEY.md section 7 hard part 3):
  * ``table_mode="dense"``  -- exact reference semantics: the table gradient is
    scattered into a dense buffer and every row gets the Adam update (moments
    decay even for untouched rows).  Right for ML-1M sized tables and for
    multi-step parity tests.
  * ``table_mode="sparse"`` -- the sorted-segment kernel hands (rows, row_grad)
    to ``tt_emb_rowwise_adam``; only touched rows move ("lazy Adam": identical
    to dense Adam the first time a row is touched, diverges afterwards).  The
    dense [V, D] gradient never exists; required for the 100M-row config.
In both modes the global gradient norm includes the table gradients BEFORE
any update is applied (two-phase: reduce + norm, then apply).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .modules import GenericTower, SequenceFeatureProcessor, ShardedEmbedding, TwoTowerModel


def _embedding_tables(model: nn.Module) -> Dict[int, nn.Parameter]:
    """id(param) -> param for every GenericTower nn.Embedding table (the big id / history tables).
    Sequence-encoder tables are small (vocab x 8..32) in every reference config and stay dense."""
    out = {}
    for mod in model.modules():
        if isinstance(mod, GenericTower):
            for sub in mod.embeddings.values():
                if isinstance(sub, nn.Embedding):
                    out[id(sub.weight)] = sub.weight
    return out


class FusedTwoTowerOptimizer:
    def __init__(self, model: TwoTowerModel, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 1.0, table_mode: str = "dense"):
        if table_mode not in ("dense", "sparse"):
            raise ValueError("table_mode must be 'dense' or 'sparse'")
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise ops.TTError("move the model to the CUDA device before building FusedTwoTowerOptimizer")
        dev = params[0].device
        self.model = model
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        self.max_grad_norm = float(max_grad_norm)
        self.table_mode = table_mode
        self.param_groups = [{"lr": self.lr, "params": params}]  # torch.optim-like surface used by the loop

        tables = _embedding_tables(model) if table_mode == "sparse" else {}
        self.sparse_tables: List[nn.Parameter] = [p for p in params if id(p) in tables]
        # row-sharded tables (modules.ShardedEmbedding): always touched-rows-only, updated on their owners by the group
        self.shard_group = getattr(model, "shard_group", None)
        sharded_ids = {id(m.weight) for m in model.modules() if isinstance(m, ShardedEmbedding)}
        if sharded_ids and self.shard_group is None:
            raise ops.TTError("model has row-sharded tables but no shard group (build it through TwoTowerModel)")
        self.sharded_params = [p for p in params if id(p) in sharded_ids]
        flat_params = [p for p in params if id(p) not in tables and id(p) not in sharded_ids]

        # one flat fp32 buffer each for params / grads / exp_avg / exp_avg_sq
        n = sum(p.numel() for p in flat_params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        # + one trailing float: the row-sharded tables' local sum of squares rides the dense-gradient all-reduce
        self.flat_g_ext = torch.zeros(n + 1, dtype=torch.float32, device=dev)
        self.flat_g = self.flat_g_ext[:n]
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self._views = []
        for p in flat_params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            p._tt_grad_direct = True      # ops.LinearFn / AddDropoutLayerNorm add into the flat buffer themselves
            self._views.append((p, off, k))
            off += k
        self.flat_params = flat_params

        # sparse tables: per-table moments + one sink shared by both towers
        self.sink = ops.SparseGradSink()
        self.table_state = {id(p): (torch.zeros_like(p, dtype=torch.float32), torch.zeros_like(p, dtype=torch.float32))
                            for p in self.sparse_tables}
        # [0] dense parameters, [1..] one slot per sparse-gradient entry of the step (never shared: the two towers'
        # backward passes run on different streams)
        self.sq_terms = torch.zeros(2 + ops.SparseGradSink.MAX_ENTRIES, dtype=torch.float32, device=dev)
        self.sink.sq_terms = self.sq_terms
        # learning rate on the device: the Adam kernels read it there, so a scheduler that edits
        # param_groups[0]["lr"] reaches a captured CUDA graph too (refreshed by step() / GraphedTrainStep.__call__)
        self.lr_dev = torch.full((1,), self.lr, dtype=torch.float64, device=dev)
        self.coef = torch.ones(1, dtype=torch.float32, device=dev)
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._ws = torch.empty(4096, dtype=torch.uint8, device=dev)
        if table_mode == "sparse":
            sparse_ids = set(tables)
            for mod in model.modules():
                if isinstance(mod, GenericTower):
                    mod.sparse_sink = self.sink
                    mod.sparse_grad_tables = {name for name, sub in mod.embeddings.items()
                                              if isinstance(sub, nn.Embedding) and id(sub.weight) in sparse_ids}
        self._sparse_param_ids = {id(p) for p in self.sparse_tables}
        self._norm_staged_by_caller = False     # dist.ShardedTrainStep stages + all-reduces the sharded norm itself
        if self.shard_group is not None:
            self.shard_group.init_state()

    # torch.optim-like surface ------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        self.flat_g_ext.zero_()
        self.sq_terms.zero_()
        self.sink.clear()
        if self.shard_group is not None:
            self.shard_group.zero_grad()

    SHARD_SLOT = 1 + ops.SparseGradSink.MAX_ENTRIES     # last slot of sq_terms: row-sharded tables, summed over ranks

    def stage_sharded_norm(self):
        """Put this rank's sum of squared row-sharded gradients behind the dense gradients (flat_g_ext[-1]); after the
        caller's SUM all-reduce of flat_g_ext it holds the global value that step() feeds to the clip coefficient."""
        if self.shard_group is not None and self.shard_group.tables:
            self.flat_g_ext[-1:].copy_(self.shard_group.local_sq_norm())

    def sync_lr(self):
        """Copy param_groups[0]['lr'] (what torch LR schedulers edit) to the device scalar the kernels read.  Call
        outside CUDA-graph capture; step() does it when not capturing, GraphedTrainStep before every replay."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self.lr:
            self.lr = lr
            self.lr_dev.fill_(lr)

    def _merged_entries(self):
        """One (table, rows, row_grad, n_unique) per table.  A table that met several backward calls in this step (the
        item tower run once per hard-negative slab, group_hard_negatives=False) gets its entries merged by one more
        sorted-segment reduction over (rows, row gradients): Adam then sees sum_i g_i once -- what autograd's
        accumulation into .grad gives the reference -- and the norm term is ||sum_i g_i||^2."""
        by_table: Dict[int, list] = {}
        for e in self.sink.entries:
            by_table.setdefault(id(e[0]), []).append(e)
        out = []
        for group in by_table.values():
            if len(group) == 1:
                out.append(group[0][:4])
                continue
            table = group[0][0]
            rows_all, grads_all = [], []
            for _, rows, row_grad, n_unique, slot in group:
                live = torch.arange(rows.numel(), device=rows.device) < n_unique
                rows_all.append(torch.where(live, rows, torch.full_like(rows, -1)))   # id -1: dropped by the kernel
                grads_all.append(row_grad)
                self.sq_terms[slot].zero_()
            slot = group[0][4]
            rows, row_grad, n_unique = ops.segment_grad(torch.cat(rows_all).view(-1, 1), ops.POOL_NONE, None,
                                                        table.shape[0], torch.cat(grads_all), None, table.shape[1],
                                                        self.sq_terms[slot:slot + 1])
            out.append((table, rows, row_grad, n_unique))
        return out

    def step(self):
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        entries = self._merged_entries()
        if self.shard_group is not None and self.shard_group.tables:
            if not self._norm_staged_by_caller:
                self.stage_sharded_norm()
            self.sq_terms[self.SHARD_SLOT:self.SHARD_SLOT + 1].copy_(self.flat_g_ext[-1:])
        ops.sq_norm_accum_(self.flat_g, self.sq_terms[0:1], self._ws)
        coef = None
        if self.max_grad_norm > 0:
            ops.clip_coef_(self.sq_terms, self.max_grad_norm, self.coef, self.total_norm)
            coef = self.coef
        self.step_dev.add_(1)
        ops.adam_flat_(self.flat_p, self.flat_g, self.flat_m, self.flat_v, coef, self.lr, self.beta1, self.beta2,
                       self.eps, self.step_dev, self.lr_dev)
        for table, rows, row_grad, n_unique in entries:
            m, v = self.table_state[id(table)]
            ops.rowwise_adam_(table.data, m, v, rows, row_grad, n_unique, coef, self.lr, self.beta1, self.beta2,
                              self.eps, self.step_dev, self.lr_dev)
        self.sink.clear()
        if self.shard_group is not None and self.shard_group.tables:
            self.shard_group.step(coef, self.lr, self.step_dev, (self.beta1, self.beta2), self.eps, self.lr_dev)

    # checkpoint surface: the SAME dict torch.optim.Adam(model.parameters()) produces (train_twotower.py:184-195 stores
    # optimizer.state_dict()), so checkpoints move freely between the reference's optimizer and this one
    def _param_states(self):
        """(param, exp_avg view, exp_avg_sq view) in model.parameters() order."""
        flat = {id(p): (off, k) for p, off, k in self._views}
        out = []
        for p in self.param_groups[0]["params"]:
            if id(p) in flat:
                off, k = flat[id(p)]
                out.append((p, self.flat_m[off:off + k].view_as(p), self.flat_v[off:off + k].view_as(p)))
            elif id(p) in self.table_state:
                m, v = self.table_state[id(p)]
                out.append((p, m, v))
            else:   # row-sharded table: the moments of the LOCAL shard (checkpoints of sharded runs are per rank)
                t = next(t for t in self.shard_group.tables.values() if t.weight is p)
                out.append((p, t.exp_avg, t.exp_avg_sq))
        return out

    def state_dict(self):
        step = float(self.step_dev.item())
        state = {}
        if step > 0:
            for i, (_, m, v) in enumerate(self._param_states()):
                state[i] = {"step": torch.tensor(step), "exp_avg": m.detach().clone(), "exp_avg_sq": v.detach().clone()}
        group = {"lr": self.lr, "betas": (self.beta1, self.beta2), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.param_groups[0]["params"])))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam.state_dict() (or our own) for the same model.  Lazy-Adam tables (table_mode='sparse')
        share one global step count, like the dense path."""
        states = self._param_states()
        if len(sd["param_groups"][0]["params"]) != len(states):
            raise ValueError("optimizer state_dict does not match this model's parameter list")
        g = sd["param_groups"][0]
        self.lr = float(g.get("lr", self.lr))
        self.beta1, self.beta2 = (float(b) for b in g.get("betas", (self.beta1, self.beta2)))
        self.eps = float(g.get("eps", self.eps))
        self.param_groups[0]["lr"] = self.lr
        self.lr_dev.fill_(self.lr)
        step = 0.0
        with torch.no_grad():
            for i, (_, m, v) in enumerate(states):
                st = sd["state"].get(i)
                if st is None:
                    m.zero_()
                    v.zero_()
                    continue
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                step = max(step, float(st["step"]))
            self.step_dev.fill_(int(step))


class GraphedTrainStep:
    """zero_grad -> model(batch) -> compute_loss -> backward -> clip + Adam as ONE CUDA graph
    (training_utils.py:28-60 of the reference, minus its host syncs).  The batch is copied into
    static device buffers; ``__call__`` returns the (device) loss tensor of the step."""

    def __init__(self, model: TwoTowerModel, optimizer: FusedTwoTowerOptimizer, example_batch: dict,
                 temperature: float, item_id_col: int = 0, warmup: int = 3):
        self.model, self.opt, self.temperature, self.item_id_col = model, optimizer, temperature, item_id_col
        self.static_batch = _clone_tree(example_batch)
        self.graph = torch.cuda.CUDAGraph()
        # warm-up steps are real optimizer steps: snapshot and restore so construction has no side effect
        snap_model = {k: v.clone() for k, v in model.state_dict().items()}
        snap_opt = (optimizer.flat_m.clone(), optimizer.flat_v.clone(), optimizer.step_dev.clone(),
                    {k: (m.clone(), v.clone()) for k, (m, v) in optimizer.table_state.items()})
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        model.load_state_dict(snap_model)
        optimizer.flat_m.copy_(snap_opt[0])
        optimizer.flat_v.copy_(snap_opt[1])
        optimizer.step_dev.copy_(snap_opt[2])
        for k, (m, v) in snap_opt[3].items():
            optimizer.table_state[k][0].copy_(m)
            optimizer.table_state[k][1].copy_(v)
        c0 = ops.launch_counter["calls"]
        with torch.cuda.graph(self.graph):
            self.static_loss = self._step_eager()
        # hand-written kernels recorded in the graph (torch/cuBLAS/cub kernels are not counted)
        self.launches_per_step = ops.launch_counter["calls"] - c0
        torch.cuda.synchronize()

    def _step_eager(self):
        nvtx = torch.cuda.nvtx      # ranges show up in Nsight Systems / ncu --nvtx (tracing)
        self.opt.zero_grad()
        nvtx.range_push("tt.forward")
        u, i, hn = self.model(self.static_batch)
        nvtx.range_pop()
        ids = self.static_batch["item_tower"]["sparse"][:, self.item_id_col]
        nvtx.range_push("tt.loss")
        loss = self.model.compute_loss(u, i, item_ids=ids, hard_neg_emb=hn, temperature=self.temperature)
        nvtx.range_pop()
        nvtx.range_push("tt.backward")
        loss.backward()
        nvtx.range_pop()
        nvtx.range_push("tt.optimizer")
        self.opt.step()
        nvtx.range_pop()
        return loss.detach()

    def load_batch(self, batch: dict, non_blocking: bool = True):
        _copy_tree(self.static_batch, batch, non_blocking)

    def __call__(self, batch: Optional[dict] = None):
        if batch is not None:
            self.load_batch(batch)
        self.opt.sync_lr()
        self.graph.replay()
        return self.static_loss


def _clone_tree(obj):
    if isinstance(obj, torch.Tensor):
        return obj.clone()
    if isinstance(obj, dict):
        return {k: _clone_tree(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_clone_tree(v) for v in obj]
    return obj


def _copy_tree(dst, src, non_blocking):
    if isinstance(dst, torch.Tensor):
        if dst.shape != src.shape:
            raise ops.TTError(f"GraphedTrainStep needs fixed batch shapes (got {tuple(src.shape)}, captured {tuple(dst.shape)})")
        dst.copy_(src, non_blocking=non_blocking)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_tree(dst[k], src[k], non_blocking)
    elif isinstance(dst, list):
        for a, b in zip(dst, src):
            _copy_tree(a, b, non_blocking)
