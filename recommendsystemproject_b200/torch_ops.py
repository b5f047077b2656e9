"""`torch.ops.tt_b200.*`: the hot-path kernels registered as torch custom ops (torch.library), CUDA only.

The north star asks for "a thin C-ABI torch custom-op layer".  `ops.py` is that layer for the drop-in modules
(`torch.autograd.Function`s that also carry the optimizer plumbing: sparse-gradient sinks, preallocated .grad buffers,
device flag words).  This module exposes the four north-star kernels as dispatcher-visible ops with schemas, fake
(meta) implementations and registered autograd, for callers that want to compose them with other torch code
(`torch.compile` tracing, `torch.library.opcheck`, export) without the module classes:

    torch.ops.tt_b200.gather_pool(table, ids, mode, padding_idx) -> Tensor            GenericTower.py:153-160,182
    torch.ops.tt_b200.segment_grad(ids, mode, padding_idx, vocab, grad_out)           autograd of the above
        -> (unique_rows, row_grad, n_unique)
    torch.ops.tt_b200.inbatch_ce(user, item, item_ids, temperature, precision)           TwoTowerModel.py:95-140
        -> (loss, row_lse)                                                            (autograd registered)
    torch.ops.tt_b200.score_topk(query, corpus, k, precision) -> (scores, rows)       training_utils.py:220-258

Only the CUDA key is registered: a CPU tensor fails in the dispatcher ("no CPU fallback").  Every implementation is a
call into libtt_b200.so through `ops.py` on the current stream.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch.library import custom_op

from . import ops

_PRECISIONS = {0: "fp32", 1: "bf16"}


@custom_op("tt_b200::gather_pool", mutates_args=(), device_types="cuda")
def gather_pool(table: torch.Tensor, ids: torch.Tensor, mode: int, padding_idx: int) -> torch.Tensor:
    """out[b] = pool_l table[ids[b, l]]; mode = ops.POOL_{NONE,SUM,MEAN}; padding_idx < 0: no id is special."""
    if mode == ops.POOL_MAX:
        raise ops.TTError("torch.ops.tt_b200.gather_pool: max pooling needs the argmax side output; use ops.gather_rows")
    ids = ids.contiguous().long()
    if ids.dim() == 1:
        ids = ids.unsqueeze(1)
    out = torch.empty(ids.shape[0], table.shape[1], dtype=torch.float32, device=table.device)
    oob = torch.zeros(1, dtype=torch.int32, device=table.device)
    ops.gather_pool_into(table.detach(), ids, int(mode), None if padding_idx < 0 else int(padding_idx), out, None, oob)
    return out


@gather_pool.register_fake
def _gather_pool_fake(table, ids, mode, padding_idx):
    return table.new_empty((ids.shape[0], table.shape[1]), dtype=torch.float32)


@custom_op("tt_b200::segment_grad", mutates_args=(), device_types="cuda")
def segment_grad(ids: torch.Tensor, mode: int, padding_idx: int, vocab: int,
                 grad_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Sorted-segment scatter-add: (unique_rows int64 [n_pos], row_grad [n_pos, D], n_unique int32 [1]); the first
    n_unique entries are meaningful, rows ascending, deterministic summation order."""
    ids = ids.contiguous().long()
    if ids.dim() == 1:
        ids = ids.unsqueeze(1)
    g = grad_out.contiguous().float()
    return ops.segment_grad(ids, int(mode), None if padding_idx < 0 else int(padding_idx), int(vocab), g, None, g.shape[1])


@segment_grad.register_fake
def _segment_grad_fake(ids, mode, padding_idx, vocab, grad_out):
    n = ids.numel()
    return (ids.new_empty((n,), dtype=torch.int64), grad_out.new_empty((n, grad_out.shape[1]), dtype=torch.float32),
            ids.new_empty((1,), dtype=torch.int32))


def _gather_pool_setup(ctx, inputs, output):
    table, ids, mode, padding_idx = inputs
    ctx.save_for_backward(ids)
    ctx.cfg = (int(mode), int(padding_idx), table.shape[0], table.shape[1])


def _gather_pool_backward(ctx, grad):
    (ids,) = ctx.saved_tensors
    mode, pad, vocab, dim = ctx.cfg
    if ids.dim() == 1:
        ids = ids.unsqueeze(1)
    rows, row_grad, n_unique = torch.ops.tt_b200.segment_grad(ids, mode, pad, vocab, grad)
    dense = torch.zeros(vocab, dim, dtype=torch.float32, device=grad.device)
    ops.scatter_rows_(dense, rows, row_grad, n_unique)
    return dense, None, None, None


gather_pool.register_autograd(_gather_pool_backward, setup_context=_gather_pool_setup)


@custom_op("tt_b200::inbatch_ce", mutates_args=(), device_types="cuda")
def inbatch_ce(user: torch.Tensor, item: torch.Tensor, item_ids: torch.Tensor, temperature: float,
               precision: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss, row_lse) of the fused in-batch softmax CE; item_ids: int64 [B] (pass an empty tensor for "no mask");
    precision 0 = exact fp32 kernel, 1 = tcgen05 bf16 kernel (D in {64, 128})."""
    ids = item_ids if item_ids.numel() else None
    with torch.no_grad():
        loss, lse, _ = ops.fused_inbatch_ce(user.detach(), item.detach(), ids, None, None, float(temperature),
                                            precision=_PRECISIONS[int(precision)])
    return loss.clone(), lse.clone()


@inbatch_ce.register_fake
def _inbatch_ce_fake(user, item, item_ids, temperature, precision):
    return user.new_empty((), dtype=torch.float32), user.new_empty((user.shape[0],), dtype=torch.float32)


@custom_op("tt_b200::inbatch_ce_backward", mutates_args=(), device_types="cuda")
def inbatch_ce_backward(user: torch.Tensor, item: torch.Tensor, item_ids: torch.Tensor, temperature: float, precision: int,
                        grad_loss: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(d_user, d_item) for grad_loss (0-dim).  The kernels recompute the logits tile by tile (nothing was saved)."""
    ids = item_ids if item_ids.numel() else None
    u = user.detach().clone().requires_grad_(True)
    i = item.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss, _, _ = ops.fused_inbatch_ce(u, i, ids, None, None, float(temperature), precision=_PRECISIONS[int(precision)])
        loss.backward(grad_loss.reshape(()).to(loss.dtype))
    return u.grad, i.grad


@inbatch_ce_backward.register_fake
def _inbatch_ce_backward_fake(user, item, item_ids, temperature, precision, grad_loss):
    return torch.empty_like(user, dtype=torch.float32), torch.empty_like(item, dtype=torch.float32)


def _ce_setup(ctx, inputs, output):
    user, item, item_ids, temperature, precision = inputs
    ctx.save_for_backward(user, item, item_ids)
    ctx.cfg = (float(temperature), int(precision))


def _ce_backward(ctx, grad_loss, _grad_lse):
    user, item, item_ids = ctx.saved_tensors
    temperature, precision = ctx.cfg
    du, di = torch.ops.tt_b200.inbatch_ce_backward(user, item, item_ids, temperature, precision, grad_loss)
    return du, di, None, None, None


inbatch_ce.register_autograd(_ce_backward, setup_context=_ce_setup)


@custom_op("tt_b200::score_topk", mutates_args=(), device_types="cuda")
def score_topk(query: torch.Tensor, corpus: torch.Tensor, k: int, precision: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores float64 [Q, k], corpus rows int64 [Q, k]) under (score desc, row asc); precision 1 = tcgen05 filter +
    exact fp64 re-rank (the same rows as precision 0)."""
    s, i = ops.score_topk(query, corpus, int(k), precision=_PRECISIONS[int(precision)])
    return s, i


@score_topk.register_fake
def _score_topk_fake(query, corpus, k, precision):
    return (query.new_empty((query.shape[0], k), dtype=torch.float64), query.new_empty((query.shape[0], k), dtype=torch.int64))
