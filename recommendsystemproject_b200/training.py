"""Train / eval loop with the reference's function names and argument meaning
(project/utils/training_utils.py:5-275), driving the B200 kernels.

Differences from the reference loop, all behind the same call signatures:
  * with a FusedTwoTowerOptimizer the clip + Adam are device-side (no separate
    clip_grad_norm_ call); with a plain torch optimizer the loop is the
    reference's, op for op;
  * validate() scores + selects with the fused top-K kernel once for
    max(k_list) instead of materialising [B, N_items] scores and calling
    torch.topk per k; the per-user history mask is a CSR list applied inside
    the kernel instead of a python loop over users (training_utils.py:238-252).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import ops
from .optim import FusedTwoTowerOptimizer


def to_device(data, device):
    if isinstance(data, torch.Tensor):
        return data.to(device, non_blocking=True)
    if isinstance(data, dict):
        return {k: to_device(v, device) for k, v in data.items()}
    if isinstance(data, list):
        return [to_device(v, device) for v in data]
    return data


def extract_item_id(item_batch, feature_name="movie_id_enc", feature_type="sparse", item_id_col=0):
    """training_utils.py:72-101."""
    if feature_type == "sparse":
        m = item_batch.get("sparse")
        if m is not None:
            return m[:, item_id_col]
    elif feature_type == "dense":
        m = item_batch.get("dense")
        if m is not None:
            return m[:, 0]
    elif feature_type == "sequence":
        seq = item_batch.get("sequence", {})
        if feature_name in seq:
            return seq[feature_name][:, 0]
    raise ValueError(f"Could not extract item ID '{feature_name}' from batch")


def build_user_history(train_df, user_col="user_id_enc", item_col="movie_id_enc"):
    """training_utils.py:103-119 (dict user id -> set of item ids)."""
    hist = {}
    for u, i in zip(train_df[user_col], train_df[item_col]):
        hist.setdefault(u, set()).add(i)
    return hist


def train_one_epoch(model, loader, optimizer, device, scheduler=None, log_every_n_batches=100, epoch=None,
                    max_grad_norm=1.0, temperature=0.1, item_id_feature="movie_id_enc", item_id_type="sparse"):
    """training_utils.py:19-70."""
    model.train()
    total_loss = 0.0
    fused = isinstance(optimizer, FusedTwoTowerOptimizer)
    n_batches = 0
    for batch_idx, batch in enumerate(loader):
        batch = to_device(batch, device)
        optimizer.zero_grad()
        user_emb, item_emb, hn = model(batch)
        ids = extract_item_id(batch["item_tower"], feature_name=item_id_feature, feature_type=item_id_type)
        loss = model.compute_loss(user_emb, item_emb, hard_neg_emb=hn, item_ids=ids, temperature=temperature)
        loss.backward()
        if max_grad_norm > 0 and not fused:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
        optimizer.step()
        if scheduler is not None:
            scheduler.step()
        total_loss += loss.item()
        n_batches += 1
        if batch_idx % log_every_n_batches == 0:
            print(f"epoch {epoch} batch {batch_idx} loss {loss.item():.4f} lr {optimizer.param_groups[0]['lr']:.6f}")
    avg = total_loss / max(n_batches, 1)
    print(f"Epoch {epoch} finished. Avg Loss: {avg:.4f}")
    return avg


@torch.no_grad()
def encode_corpus(model, item_loader: Iterable, device, item_id_col: int = 0):
    """training_utils.py:153-170: item tower in eval mode over the catalog."""
    embs, ids = [], []
    for item_batch in item_loader:
        item_batch = to_device(item_batch, device)
        embs.append(model.get_item_embeddings(item_batch))
        ids.append(extract_item_id(item_batch, feature_type="sparse", item_id_col=item_id_col))
    return torch.cat(embs, dim=0), torch.cat(ids, dim=0).view(-1)


def history_csr(user_ids: np.ndarray, user_history: Dict, id_to_row: np.ndarray):
    """Per-query sorted corpus rows to exclude (training_utils.py:238-252) as CSR arrays."""
    offsets = np.zeros(len(user_ids) + 1, dtype=np.int64)
    chunks = []
    max_id = len(id_to_row) - 1
    for n, u in enumerate(user_ids):
        rows = np.empty(0, dtype=np.int64)
        if u in user_history:
            items = np.fromiter((i for i in user_history[u] if 0 <= i <= max_id), dtype=np.int64)
            if items.size:
                rows = id_to_row[items]
                rows = np.unique(rows[rows >= 0])
        chunks.append(rows)
        offsets[n + 1] = offsets[n] + rows.size
    flat = np.concatenate(chunks) if chunks else np.empty(0, dtype=np.int64)
    return offsets, flat


@torch.no_grad()
def validate(model, loader, item_loader, device, epoch, k_list=[10, 20], item_id_feature="movie_id_enc",
             item_id_type="sparse", item_id_col_idx=0, meta_data_loader=None, user_id_col_idx=None,
             log_embeddings=True, user_history=None):
    """training_utils.py:121-275."""
    model.eval()
    total_loss = 0.0
    hits = {k: 0 for k in k_list}
    num_samples = 0
    all_embs, all_ids = encode_corpus(model, item_loader, device)
    max_id = int(all_ids.max().item())
    id_to_row = torch.full((max_id + 1,), -1, dtype=torch.long, device=device)
    id_to_row[all_ids] = torch.arange(len(all_ids), device=device)
    id_to_row_np = id_to_row.cpu().numpy()
    if log_embeddings and epoch is not None:
        _log_embedding_stats(all_embs, epoch)
    meta_iter = iter(meta_data_loader) if meta_data_loader is not None else None
    kmax = max(k_list)
    # tensor-core scoring (bf16 filter + exact fp64 re-rank: the SAME rows as the fp32 path, see include/tt_b200.h)
    # whenever the shapes allow it; the bf16 corpus copy is made once per catalog encode
    use_tc = all_embs.shape[1] in (64, 128) and min(kmax, all_embs.shape[0]) <= 224 and all_embs.shape[0] >= 4096
    prepared = ops.PreparedCorpus(all_embs) if use_tc else None
    n_batches = 0
    for batch in loader:
        batch = to_device(batch, device)
        user_emb, item_emb, hn = model(batch)
        item_batch = batch.get("item_tower", {})
        if not item_batch:
            raise ValueError("batch_data does not contain 'item_tower' key")
        if item_id_type == "sparse" and "sparse" in item_batch:
            targets = item_batch["sparse"][:, item_id_col_idx]
        elif item_id_type == "dense" and "dense" in item_batch:
            targets = item_batch["dense"][:, item_id_col_idx]
        elif item_id_type == "sequence" and "sequence" in item_batch:
            targets = item_batch["sequence"][item_id_feature][:, 0]
        else:
            raise ValueError(f"Cannot extract target item IDs from batch. item_batch keys: {item_batch.keys()}, "
                             f"looking for type: {item_id_type}")
        # NB: like the reference, validation uses compute_loss's DEFAULT temperature (0.1)
        loss = model.compute_loss(user_emb, item_emb, hard_neg_emb=hn, item_ids=targets)
        total_loss += loss.item()
        mask_off = mask_rows = None
        if user_history is not None:
            if meta_iter is not None:
                user_ids = next(meta_iter)["user_tower"]["sparse"][:, 0]
            elif user_id_col_idx is not None:
                user_ids = batch["user_tower"]["sparse"][:, user_id_col_idx]
            else:
                raise ValueError("Either metadata_loader or user_id_col_idx required")
            off, flat = history_csr(user_ids.cpu().numpy(), user_history, id_to_row_np)
            if flat.size:
                mask_off = torch.from_numpy(off).to(device)
                mask_rows = torch.from_numpy(flat).to(device)
        _, top_rows = ops.score_topk(user_emb, all_embs, min(kmax, all_embs.shape[0]), 0, mask_off, mask_rows,
                                     precision="bf16" if use_tc else "fp32", prepared=prepared)
        valid = top_rows >= 0
        pred_ids = all_ids[top_rows.clamp(min=0)]
        match = (pred_ids == targets.view(-1, 1)) & valid
        for k in k_list:
            hits[k] += int(match[:, :k].any(dim=1).sum().item())
        num_samples += len(targets)
        n_batches += 1
    avg_loss = total_loss / max(n_batches, 1)
    acc = {k: hits[k] / max(num_samples, 1) for k in k_list}
    print(f"\nValidation Result - Loss: {avg_loss:.4f}")
    for k, a in acc.items():
        print(f"Recall@{k}: {a:.4f}")
    return avg_loss, acc


def _log_embedding_stats(all_item_embs, epoch):
    """Print-only diagnostics (training_utils.py:277-331)."""
    std = all_item_embs.std(dim=0).mean().item()
    mean_norm = all_item_embs.mean(dim=0).norm().item()
    n = all_item_embs.shape[0]
    sample = all_item_embs[torch.randperm(n, device=all_item_embs.device)[:1000]] if n > 1000 else all_item_embs
    d = torch.cdist(sample, sample)
    off = ~torch.eye(d.shape[0], dtype=torch.bool, device=d.device)
    print(f"Epoch {epoch} - item embedding std {std:.6f}, mean-norm {mean_norm:.6f}, pairwise dist "
          f"avg {d[off].mean().item():.6f} min {d[off].min().item():.6f} max {d[off].max().item():.6f}, items {n}")


# ---------------------------------------------------------------------------- checkpoints
def save_checkpoint(path, epoch, model, optimizer, train_loss=None, val_loss=None, metrics=None, user_mapping=None,
                    item_mapping=None, config=None):
    """The reference's checkpoint dict, key for key (train_twotower.py:184-195)."""
    import os
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                "train_loss": train_loss, "val_loss": val_loss, "metrics": metrics, "user_mapping": user_mapping,
                "item_mapping": item_mapping, "config": config}, path)


def load_checkpoint(path, model, optimizer=None, map_location=None):
    """Resume (the reference only saves): restores model weights / BatchNorm statistics and, when given, the optimizer
    (torch.optim.Adam or FusedTwoTowerOptimizer; either kind of checkpoint loads into either kind of optimizer).
    Returns the checkpoint dict minus the two state dicts (epoch, losses, metrics, mappings, config)."""
    # tensors + plain dicts / lists / scalars only: never unpickle arbitrary objects from a checkpoint file
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    model.load_state_dict(ckpt["model_state_dict"])
    if optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return {k: v for k, v in ckpt.items() if k not in ("model_state_dict", "optimizer_state_dict")}
