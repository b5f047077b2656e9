"""Row-sharded embedding tables behind GenericTower (BASELINE configs[2]).

The reference is single-process: every ``nn.Embedding`` of GenericTower.py:30-56 lives whole on one device.  At the
100M-user / 10M-item scale the big tables are spread over the GPUs of one NVSwitch box instead:
owner(row) = row % W, local row = row // W.  One ``ShardedTableGroup`` serves ALL sharded features of BOTH towers so
that a training step has exactly

    forward   route ids (CUDA)  ->  all-to-all #1 (int32 rows + offsets, all tables in one block per peer)
              ->  owner-side gather + pool (CUDA)  ->  all-to-all #2 (fp32 partial sums / rows)  ->  combine (CUDA)
    backward  pack d(pooled) (CUDA)  ->  all-to-all #3  ->  owner-side sorted-segment reduction (CUDA)
    step      global-norm clip (one scalar all-reduce, done by the caller) + row-wise Adam on the local shards

Block sizes are capacities fixed by (batch, len, W, capacity_factor): nothing is counted on the host, so the whole
step, collectives included, is captured in one CUDA graph.  An owner whose share of a batch exceeds the capacity drops
the excess and raises the overflow flag (checked by ``check_flags``); ids are hashed by ``row % W`` so the share is
B*L/W up to sampling noise and id skew, which ``capacity_factor`` (default 1.25) covers.

What an owner does is proportional to the entries it RECEIVES (~ B*L_valid per table whatever W is), which is what
makes the exchange scale; round 1's version sent every owner the full [B, L] id matrix with foreign ids blanked.

Device work goes through ``dev_ops`` (default: the libtt_b200 kernels of include/tt_b200.h section 7); the CPU gloo
tests inject a restatement (tests/sharded_cpu_ops.py) -- the product itself has no CPU path.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import TTError, check

_p, _stream = ops._p, ops._stream


class _CudaShardOps:
    """include/tt_b200.h section 7 + tt_emb_segment_grad_lists + tt_emb_rowwise_adam."""
    _route_ws: Dict = {}      # tile totals of the route scan, one per (device, stream): calls on ONE stream run in order

    @staticmethod
    def route(ids, pad, vocab, world, send, block_ints, off_base, rows_base, cap, n_pad, flags):
        lib = _lib.load()
        n_rows, length = ids.shape
        key = (ids.device, _stream().value)
        ws = _CudaShardOps._route_ws.get(key)
        need = 4 * world * ((n_rows + 4095) // 4096)
        if ws is None or ws.numel() < need:
            ws = _CudaShardOps._route_ws[key] = torch.empty(max(need, 1024), dtype=torch.uint8, device=ids.device)
        check(lib.tt_shard_route(_p(ids), n_rows, length, -1 if pad is None else int(pad), vocab, world, _p(send), block_ints,
                                 off_base, rows_base, cap, _p(n_pad), _p(flags), _p(ws), ws.numel(), _stream()), "tt_shard_route")
        ops._count(4)

    @staticmethod
    def owner_gather(table, local_rows, world, recv, block_ints, off_base, rows_base, cap, n_rows, pooled, out, block_floats,
                     vec_base, pos_src):
        lib = _lib.load()
        check(lib.tt_shard_owner_gather(_p(table), ops._DTYPES[table.dtype], local_rows, table.shape[1], world, _p(recv),
                                        block_ints, off_base, rows_base, cap, n_rows, 1 if pooled else 0, _p(out),
                                        block_floats, vec_base, _p(pos_src), _stream()), "tt_shard_owner_gather")
        ops._count()

    @staticmethod
    def combine(recv_vec, block_floats, vec_base, world, ids, pad, vocab, mode, send, block_ints, off_base, cap, n_pad,
                pad_row, dim, out):
        lib = _lib.load()
        n_rows, length = ids.shape
        check(lib.tt_shard_combine(_p(recv_vec), block_floats, vec_base, world, _p(ids), n_rows, length,
                                   -1 if pad is None else int(pad), vocab, mode, _p(send), block_ints, off_base, cap,
                                   _p(n_pad), _p(pad_row), dim, _p(out), out.stride(0), _stream()), "tt_shard_combine")
        ops._count()

    @staticmethod
    def grad_pack(grad, mode, dim, world, ids, pad, vocab, send, block_ints, off_base, cap, send_vec, block_floats, vec_base):
        lib = _lib.load()
        n_rows, length = ids.shape
        check(lib.tt_shard_grad_pack(_p(grad), grad.stride(0), n_rows, length, mode, dim, world, _p(ids),
                                     -1 if pad is None else int(pad), vocab, _p(send), block_ints, off_base, cap,
                                     _p(send_vec), block_floats, vec_base, _stream()), "tt_shard_grad_pack")
        ops._count()

    @staticmethod
    def segment_grad_lists(recv, world, block_ints, rows_base, cap, pos_src, local_rows, grad, piece_rows, block_floats,
                           vec_base, dim, rows_out, row_grad, n_unique, sq_norm, ws):
        lib = _lib.load()
        check(lib.tt_emb_segment_grad_lists(ctypes.c_void_p(recv.data_ptr() + 4 * rows_base), world, cap, block_ints,
                                            _p(pos_src), max(local_rows, 1),
                                            # pos_src holds float4 offsets from the START of the gradient blocks
                                            _p(grad) if pos_src is not None else ctypes.c_void_p(grad.data_ptr() + 4 * vec_base),
                                            piece_rows, block_floats, dim,
                                            _p(rows_out), _p(row_grad), _p(n_unique), _p(sq_norm), _p(ws), ws.numel(),
                                            _stream()), "tt_emb_segment_grad_lists")
        ops._count(7)

    @staticmethod
    def segment_adam_lists(n_pos, pos_src, grad, piece_rows, block_floats, dim, rows, ws, table, m, v, coef, lr, b1, b2, eps,
                           step_dev, lr_dev=None):
        """Second phase of the deferred form (include/tt_b200.h, tt_emb_segment_adam_lists): the first phase was
        segment_grad_lists(..., row_grad=None, ...) over the same buffers."""
        lib = _lib.load()
        check(lib.tt_emb_segment_adam_lists(n_pos, _p(pos_src), _p(grad), piece_rows, block_floats, dim, _p(rows), rows.numel(),
                                            _p(ws), ws.numel(), _p(table), _p(m), _p(v), _p(coef), lr, b1, b2, eps,
                                            _p(step_dev), _p(lr_dev), _stream()), "tt_emb_segment_adam_lists")
        ops._count()

    @staticmethod
    def segment_ws_bytes(n_pos, dim):
        lib = _lib.load()
        nbytes = ctypes.c_size_t(0)
        check(lib.tt_emb_segment_grad_workspace(n_pos, dim, ctypes.byref(nbytes)), "tt_emb_segment_grad_workspace")
        return nbytes.value

    @staticmethod
    def adam(table, m, v, rows, row_grad, n_unique, coef, lr, b1, b2, eps, step_dev, lr_dev=None):
        ops.rowwise_adam_(table, m, v, rows, row_grad, n_unique, coef, lr, b1, b2, eps, step_dev, lr_dev)


class _Holder:
    def __init__(self, weight, pad_row):
        self.weight, self.pad_row = weight, pad_row


class ShardedTable:
    """One row-sharded table: local shard + Adam moments.  `holder` owns `.weight` ([local_rows, dim] parameter / tensor,
    fp32 or bf16) and `.pad_row` ([dim] fp32 replica of the frozen pad row, or None); they are read through the holder at
    every use because nn.Module.to() replaces buffers."""

    def __init__(self, name, vocab, dim, mode, padding_idx, holder):
        self.name, self.vocab, self.dim, self.mode, self.padding_idx = name, int(vocab), int(dim), mode, padding_idx
        self.holder = holder
        self.exp_avg = None
        self.exp_avg_sq = None
        self.pending = None                    # (rows, row_grad, n_unique) of the last backward

    @property
    def weight(self):
        return self.holder.weight

    @property
    def pad_row(self):
        return self.holder.pad_row

    @property
    def local_rows(self):
        return self.weight.shape[0]


class _Plan:
    """Wire layout for one (batch, lens) signature."""
    pass


class ShardedTableGroup:
    def __init__(self, rank: int, world: int, device=None, capacity_factor: float = 1.25, dev_ops=None, group=None):
        if not (1 <= world <= 32):
            raise TTError("ShardedTableGroup supports 1..32 ranks")
        self.rank, self.world = int(rank), int(world)
        self._device = None if device is None else torch.device(device)
        self.capacity_factor = float(capacity_factor)
        self.ops = dev_ops or _CudaShardOps
        self.pg = group
        self.tables: Dict[str, ShardedTable] = {}
        self._plans: Dict[tuple, _Plan] = {}
        self._flags = None                     # bit 0: id out of range, bit 1: capacity overflow
        self._sq_terms = None
        self._anchor_t = None
        self.a2a_bytes = 0                     # bytes this rank SENT to other ranks (counted per call)
        # per-table chains (route / owner gather / combine / segment reduction / Adam) touch disjoint buffers: on a GPU
        # each table gets its own stream between the collectives ("lanes"), forked from and joined back into the
        # current stream -- inside a CUDA-graph capture they become parallel branches, which is what hides the
        # ~15 short launches per small table behind the big table's HBM-bound kernels.  TT_SHARD_LANES=0 disables.
        self.parallel_lanes = os.environ.get("TT_SHARD_LANES", "1") != "0"
        self._lanes: Dict = {}
        # deferred segment gradient: the backward keeps only the sorted lists and the gradient norm, step() forms each
        # touched row's gradient sum again inside the Adam kernel (no [U, D] row_grad buffer through HBM, and none
        # allocated).  Per table, when the device ops offer it and the table is fp32 with dim in {64, 96, 128, 192, 256}.
        # Set before the first lookup (the choice is part of the wire plan); TT_SEG_ADAM_FUSED=0 disables.
        self.fused_adam = os.environ.get("TT_SEG_ADAM_FUSED", "1") != "0"

    @property
    def device(self):
        """Where the shards live NOW (a model is built on the CPU and moved with .to(), like the reference's)."""
        if self.tables:
            return next(iter(self.tables.values())).weight.device
        return self._device or torch.device("cpu")

    def _dev_state(self):
        dev = self.device
        if self._flags is None or self._flags.device != dev:
            self._flags = torch.zeros(1, dtype=torch.int32, device=dev)
            self._sq_terms = torch.zeros(max(1, len(self.tables)), dtype=torch.float32, device=dev)
            # autograd anchor: the lookup's inputs are integer ids, so something that requires grad must enter the node
            self._anchor_t = torch.zeros(1, dtype=torch.float32, device=dev, requires_grad=True)
            self._plans.clear()
            for t in self.tables.values():
                t.exp_avg = t.exp_avg.to(dev) if t.exp_avg is not None else None
                t.exp_avg_sq = t.exp_avg_sq.to(dev) if t.exp_avg_sq is not None else None

    @property
    def flags(self):
        self._dev_state()
        return self._flags

    @property
    def sq_terms(self):
        self._dev_state()
        return self._sq_terms

    @property
    def _anchor(self):
        self._dev_state()
        return self._anchor_t

    # ------------------------------------------------------------------ construction
    @staticmethod
    def local_row_count(vocab: int, rank: int, world: int) -> int:
        return (int(vocab) - rank + world - 1) // world

    def add_table(self, name, vocab, dim, mode, padding_idx, weight, pad_row=None, holder=None) -> ShardedTable:
        """`holder` (an object with .weight / .pad_row, e.g. modules.ShardedEmbedding) or plain tensors."""
        if holder is None:
            holder = _Holder(weight, pad_row)
        if mode not in (ops.POOL_NONE, ops.POOL_SUM, ops.POOL_MEAN):
            raise TTError(f"row-sharded feature '{name}': pooling must be mean / sum (or a single id per sample)")
        if holder.weight.shape[0] != self.local_row_count(vocab, self.rank, self.world):
            raise TTError(f"row-sharded feature '{name}': shard has {holder.weight.shape[0]} rows, expected "
                          f"{self.local_row_count(vocab, self.rank, self.world)} (vocab {vocab}, rank {self.rank}/{self.world})")
        t = ShardedTable(name, vocab, dim, mode, padding_idx, holder)
        self.tables[name] = t
        self._plans.clear()
        self._flags = None          # device state is rebuilt (one norm slot per table)
        return t

    def init_state(self):
        for t in self.tables.values():
            if t.exp_avg is None:
                t.exp_avg = torch.zeros(t.weight.shape, dtype=torch.float32, device=t.weight.device)
                t.exp_avg_sq = torch.zeros(t.weight.shape, dtype=torch.float32, device=t.weight.device)

    # ------------------------------------------------------------------ wire layout
    def _plan(self, names: Sequence[str], shapes: Sequence[tuple]) -> _Plan:
        key = (tuple(names), tuple(shapes))
        if key in self._plans:
            return self._plans[key]
        W = self.world
        pl = _Plan()
        pl.names = list(names)
        pl.slots = {}
        ints = floats = 0
        B = shapes[0][0]
        for name, (b, L) in zip(names, shapes):
            if b != B:
                raise TTError("all sharded features of a step must share the batch size")
            t = self.tables[name]
            n_pos = B * L
            cap = n_pos if W == 1 else min(n_pos, int(n_pos / W * self.capacity_factor) + 64)
            cap = (cap + 3) // 4 * 4
            off_base = ints
            rows_base = off_base + (B + 1 + 3) // 4 * 4
            ints = rows_base + cap
            vec_base = floats
            vec_rows = B if L > 1 else cap
            floats = vec_base + vec_rows * t.dim
            pl.slots[name] = dict(len=L, cap=cap, off_base=off_base, rows_base=rows_base, vec_base=vec_base, vec_rows=vec_rows)
        pl.B = B
        pl.block_ints, pl.block_floats = ints, (floats + 3) // 4 * 4
        dev = self.device
        pl.send_ids = torch.empty(W, pl.block_ints, dtype=torch.int32, device=dev)
        pl.recv_ids = torch.empty(W, pl.block_ints, dtype=torch.int32, device=dev) if W > 1 else pl.send_ids
        pl.vec_out = torch.empty(W, pl.block_floats, dtype=torch.float32, device=dev)      # owner -> sources
        pl.vec_in = torch.empty(W, pl.block_floats, dtype=torch.float32, device=dev) if W > 1 else pl.vec_out
        pl.g_out = torch.empty(W, pl.block_floats, dtype=torch.float32, device=dev)        # sources -> owner
        pl.g_in = torch.empty(W, pl.block_floats, dtype=torch.float32, device=dev) if W > 1 else pl.g_out
        pl.n_pad, pl.pos_src, pl.seg = {}, {}, {}
        for name in names:
            s, t = pl.slots[name], self.tables[name]
            pl.n_pad[name] = torch.zeros(B, dtype=torch.int32, device=dev)
            n_pos = W * s["cap"]
            pl.pos_src[name] = torch.empty(n_pos, dtype=torch.int32, device=dev)
            ws_bytes = self.ops.segment_ws_bytes(n_pos, t.dim)
            deferred = (self.fused_adam and hasattr(self.ops, "segment_adam_lists") and t.dim in (64, 96, 128, 192, 256)
                        and t.weight.dtype == torch.float32)
            pl.seg[name] = dict(rows=torch.empty(n_pos, dtype=torch.int64, device=dev),
                                row_grad=None if deferred else torch.empty(n_pos, t.dim, dtype=torch.float32, device=dev),
                                n_pos=n_pos,
                                n_unique=torch.zeros(1, dtype=torch.int32, device=dev),
                                ws=torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev))
        self._plans[key] = pl
        return pl

    def _fan(self, jobs):
        """Run the zero-argument callables of `jobs` (one per table, disjoint outputs, NO allocations inside) each on
        its own lane stream, forked from the current stream and joined back into it before returning."""
        jobs = list(jobs)
        dev = self.device
        if not (self.parallel_lanes and dev.type == "cuda" and len(jobs) > 1):
            for job in jobs:
                job()
            return
        cur = torch.cuda.current_stream(dev)
        lanes = self._lanes.setdefault(dev.index, [])
        while len(lanes) < len(jobs) - 1:
            lanes.append(torch.cuda.Stream(device=dev))
        jobs[0]()                                   # the first table stays on the current stream
        for lane, job in zip(lanes, jobs[1:]):
            lane.wait_stream(cur)
            with torch.cuda.stream(lane):
                job()
        for lane in lanes[:len(jobs) - 1]:
            cur.wait_stream(lane)

    def _a2a(self, out, inp):
        if self.world == 1:
            return
        self.a2a_bytes += inp.numel() * inp.element_size() * (self.world - 1) // self.world
        dist.all_to_all_single(out, inp, group=self.pg)

    # ------------------------------------------------------------------ forward / backward
    def lookup(self, ids_by_name: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """ids [B] / [B, 1] / [B, L] int64 per feature -> pooled [B, dim] per feature (one batched exchange)."""
        names = [n for n in self.tables if n in ids_by_name]
        if not names:
            return {}
        ids = []
        for n in names:
            x = ids_by_name[n]
            if x.dim() == 1:
                x = x.unsqueeze(1)
            if x.dtype != torch.int64:
                x = x.long()
            ids.append(x.contiguous())
        outs = _ShardedGroupFn.apply(self, self._anchor, names, *ids)
        return dict(zip(names, outs))

    def _forward(self, names, ids):
        W = self.world
        pl = self._plan(names, [tuple(x.shape) for x in ids])
        flags = self.flags

        def route(name, x):
            s, t = pl.slots[name], self.tables[name]
            return lambda: self.ops.route(x, t.padding_idx, t.vocab, W, pl.send_ids, pl.block_ints, s["off_base"], s["rows_base"],
                                          s["cap"], pl.n_pad[name], flags)

        def owner(name):
            s, t = pl.slots[name], self.tables[name]
            weight = t.weight.detach()
            return lambda: self.ops.owner_gather(weight, t.local_rows, W, pl.recv_ids, pl.block_ints, s["off_base"],
                                                 s["rows_base"], s["cap"], pl.B, s["len"] > 1, pl.vec_out, pl.block_floats,
                                                 s["vec_base"], pl.pos_src[name])

        def combine(name, x, out):
            s, t = pl.slots[name], self.tables[name]
            pad_row = t.pad_row
            return lambda: self.ops.combine(pl.vec_in, pl.block_floats, s["vec_base"], W, x, t.padding_idx, t.vocab, t.mode,
                                            pl.send_ids, pl.block_ints, s["off_base"], s["cap"], pl.n_pad[name], pad_row,
                                            t.dim, out)

        order = self._big_first(pl, names)
        by_name = dict(zip(names, ids))
        self._fan(route(n, by_name[n]) for n in order)
        self._a2a(pl.recv_ids, pl.send_ids)                                                    # all-to-all #1
        self._fan(owner(n) for n in order)
        self._a2a(pl.vec_in, pl.vec_out)                                                       # all-to-all #2
        outs = {n: torch.empty(pl.B, self.tables[n].dim, dtype=torch.float32, device=self.device) for n in names}
        self._fan(combine(n, by_name[n], outs[n]) for n in order)
        return pl, [outs[n] for n in names]

    @staticmethod
    def _big_first(pl, names):
        """Lane order: the table with the most positions first (it stays on the current stream)."""
        return sorted(names, key=lambda n: -pl.slots[n]["len"] * pl.B)

    def _backward(self, pl, names, ids, grads):
        W = self.world
        packed = []
        for name, x, g in zip(names, ids, grads):
            t = self.tables[name]
            if g is None:
                g = torch.zeros(pl.B, t.dim, dtype=torch.float32, device=self.device)
            if not (g.stride(1) == 1 and g.stride(0) % 4 == 0 and g.data_ptr() % 16 == 0):
                g = g.contiguous()      # a column slice of the tower's concat gradient is read in place (row stride)
            packed.append((name, x, g))

        def pack(name, x, g):
            s, t = pl.slots[name], self.tables[name]
            return lambda: self.ops.grad_pack(g, t.mode, t.dim, W, x, t.padding_idx, t.vocab, pl.send_ids, pl.block_ints,
                                              s["off_base"], s["cap"], pl.g_out, pl.block_floats, s["vec_base"])

        def reduce(k, name):
            s, t, sg = pl.slots[name], self.tables[name], pl.seg[name]
            sq = self.sq_terms[k:k + 1]
            return lambda: self.ops.segment_grad_lists(pl.recv_ids, W, pl.block_ints, s["rows_base"], s["cap"], pl.pos_src[name],
                                                       t.local_rows, pl.g_in, s["vec_rows"], pl.block_floats, s["vec_base"],
                                                       t.dim, sg["rows"], sg["row_grad"], sg["n_unique"], sq, sg["ws"])

        self._fan(pack(*job) for job in packed)
        self._a2a(pl.g_in, pl.g_out)                                                           # all-to-all #3
        slot_of = {name: k for k, name in enumerate(self.tables)}
        todo = self._big_first(pl, [n for n in self.tables if n in pl.slots])
        for name in todo:
            if self.tables[name].pending is not None:
                raise TTError(f"row-sharded feature '{name}' met two backward passes in one step; call zero_grad() between steps")
        self._fan(reduce(slot_of[n], n) for n in todo)
        for name in todo:
            sg = pl.seg[name]
            self.tables[name].pending = (sg["rows"], sg["row_grad"], sg["n_unique"])
            self.tables[name].pending_plan = (pl, name)       # deferred form: step() reads the lists / gradients again
            self.tables[name].pending_positions = pl.slots[name]["len"] * pl.B

    # ------------------------------------------------------------------ optimizer side
    def zero_grad(self):
        for t in self.tables.values():
            t.pending = None
        if self.sq_terms is not None:
            self.sq_terms.zero_()

    def local_sq_norm(self) -> torch.Tensor:
        """sum of squared row gradients held by THIS rank (the caller all-reduces it into the global norm)."""
        return self.sq_terms.sum().reshape(1)

    def step(self, clip_coef, lr, step_dev, betas=(0.9, 0.999), eps=1e-8, lr_dev=None):
        self.init_state()
        live = sorted((t for t in self.tables.values() if t.pending is not None), key=lambda t: -getattr(t, "pending_positions", 0))

        def adam(t):
            rows, row_grad, n_unique = t.pending
            weight = t.weight.data if isinstance(t.weight, torch.nn.Parameter) else t.weight
            if row_grad is None:        # deferred form: segment sum + Adam in one kernel over the backward's lists
                pl, name = t.pending_plan
                s, sg = pl.slots[name], pl.seg[name]
                return lambda: self.ops.segment_adam_lists(sg["n_pos"], pl.pos_src[name], pl.g_in, s["vec_rows"], pl.block_floats,
                                                           t.dim, rows, sg["ws"], weight, t.exp_avg, t.exp_avg_sq, clip_coef, lr,
                                                           betas[0], betas[1], eps, step_dev, lr_dev)
            return lambda: self.ops.adam(weight, t.exp_avg, t.exp_avg_sq, rows, row_grad, n_unique, clip_coef, lr, betas[0],
                                         betas[1], eps, step_dev, lr_dev)

        self._fan(adam(t) for t in live)
        for t in live:
            t.pending = None

    def check_flags(self):
        """One host read: raise for ids outside a table (what nn.Embedding raises) or a capacity overflow."""
        f = int(self.flags.item())
        if f:
            self.flags.zero_()
        if f & 1:
            raise IndexError("index out of range in a row-sharded embedding feature")
        if f & 2:
            raise TTError("row-sharded exchange overflow: an owner received more ids than its capacity "
                          f"(capacity_factor={self.capacity_factor}); raise capacity_factor")

    # ------------------------------------------------------------------ checkpoints
    def gather_full_weight(self, name: str) -> torch.Tensor:
        """The whole [vocab, dim] table, assembled on every rank (gather-on-save: keeps the reference's
        state_dict keys and shapes, train_twotower.py:184-195).  Collective."""
        t = self.tables[name]
        W = self.world
        full = torch.empty(t.vocab, t.dim, dtype=t.weight.dtype, device=t.weight.device)
        max_rows = self.local_row_count(t.vocab, 0, W)
        mine = torch.zeros(max_rows, t.dim, dtype=t.weight.dtype, device=t.weight.device)
        mine[:t.local_rows] = t.weight.detach()
        if W > 1:
            parts = [torch.empty_like(mine) for _ in range(W)]
            dist.all_gather(parts, mine, group=self.pg)
        else:
            parts = [mine]
        for r in range(W):
            n = self.local_row_count(t.vocab, r, W)
            full[r::W] = parts[r][:n]
        if t.padding_idx is not None and t.pad_row is not None:
            full[t.padding_idx] = t.pad_row.to(full.dtype)
        return full

    def load_full_weight(self, name: str, full: torch.Tensor):
        t = self.tables[name]
        if full.shape[0] != t.vocab:
            raise TTError(f"row-sharded feature '{name}': full table has {full.shape[0]} rows, vocab is {t.vocab}")
        with torch.no_grad():
            t.weight.copy_(full[self.rank::self.world].to(t.weight.device, t.weight.dtype))
            if t.padding_idx is not None and t.pad_row is not None:
                t.pad_row.copy_(full[t.padding_idx].to(t.pad_row.device, torch.float32))


class _ShardedGroupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, group: ShardedTableGroup, anchor, names, *ids):
        pl, outs = group._forward(names, ids)
        ctx.group, ctx.pl, ctx.names, ctx.ids = group, pl, names, ids
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        ctx.group._backward(ctx.pl, ctx.names, ctx.ids, grads)
        return (None, None, None) + (None,) * len(ctx.ids)
