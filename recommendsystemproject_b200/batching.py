"""GPU-resident batch builder.

The reference builds every batch on the host, sample by sample: `_combined_collate_fn` calls
`RecommendationDataset.__getitem__` per index and `collate_fn` stacks / pads the results
(`CombineTwoTower.py:62-92`, `DataLoader.py:226-288`), and leaves hard negatives as a TODO (`:86-90`).  That loop
caps the reference at ~1 k samples/s end to end.  Here the pre-tensorised columns live on the
device once; a batch is a handful of `index_select`s over them, and the N hard-negative slabs are built from the
`hard_neg_ids[B, N]` column by a lookup in the item catalog (the piece the reference left unimplemented;
`parsing.py:216-250` writes the ids).  The output is exactly the collate contract:

    {'user_tower': {'sparse': int64[B,n], 'dense': f32[B,n], 'sequence': {name: int64[B,L] | [B,L,Tags]}},
     'item_tower': {...}, 'hard_negatives': [item dict x N]}

Pure tensor indexing, device-agnostic (so the logic is unit-tested on the CPU against a per-sample restatement of the
reference's collate); on the GPU every op is a gather kernel launch on the current stream and nothing syncs.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


def _take(tree, idx):
    if tree is None:
        return None
    if isinstance(tree, torch.Tensor):
        return tree.index_select(0, idx)
    return {k: _take(v, idx) for k, v in tree.items()}


def _to(tree, device):
    if tree is None:
        return None
    if isinstance(tree, torch.Tensor):
        return tree.to(device)
    return {k: _to(v, device) for k, v in tree.items()}


class GpuBatchBuilder:
    """user / item: per-interaction column groups {'sparse': [N,n] int64, 'dense': [N,n] f32, 'sequence': {name: [N,L,..]}}
    (any key may be absent, as in the reference's collate).  catalog: the same groups per catalog item, row r describing
    the item whose id is catalog_ids[r]; hard_neg_ids [N, n_neg] item ids (0 = "no negative", embedded like any item,
    parsing.py:242-245: its features are the zero row)."""

    def __init__(self, user: Dict, item: Dict, catalog: Optional[Dict] = None, catalog_ids: Optional[torch.Tensor] = None,
                 hard_neg_ids: Optional[torch.Tensor] = None, device="cuda"):
        self.device = torch.device(device)
        self.user = _to(user, self.device)
        self.item = _to(item, self.device)
        self.catalog = _to(catalog, self.device)
        self.hard_neg_ids = None if hard_neg_ids is None else hard_neg_ids.to(self.device).long()
        self.n = next(t for t in self._leaves(self.user)).shape[0]
        self.id_to_row = None
        if hard_neg_ids is not None:
            if catalog is None or catalog_ids is None:
                raise ValueError("hard negatives need the item catalog and its ids")
            ids = catalog_ids.to(self.device).long()
            # row 0 of the lookup catalog is the all-zero "no negative" item; catalog row r moves to r + 1
            self.catalog = self._with_zero_row(self.catalog)
            self.id_to_row = torch.zeros(int(ids.max().item()) + 1, dtype=torch.long, device=self.device)
            self.id_to_row[ids] = torch.arange(1, ids.numel() + 1, device=self.device)
            self.id_to_row[0] = 0

    @staticmethod
    def _leaves(tree):
        if isinstance(tree, torch.Tensor):
            yield tree
        elif tree is not None:
            for v in tree.values():
                yield from GpuBatchBuilder._leaves(v)

    @staticmethod
    def _with_zero_row(tree):
        if isinstance(tree, torch.Tensor):
            return torch.cat([torch.zeros_like(tree[:1]), tree], dim=0)
        return {k: GpuBatchBuilder._with_zero_row(v) for k, v in tree.items()}

    def __len__(self):
        return self.n

    def batch(self, indices: torch.Tensor) -> Dict:
        idx = indices.to(self.device).long()
        out = {"user_tower": _take(self.user, idx), "item_tower": _take(self.item, idx)}
        if self.hard_neg_ids is not None:
            neg = self.hard_neg_ids.index_select(0, idx)                       # [B, n_neg] item ids
            n_map = self.id_to_row.numel()                                     # ids outside [0, max catalog id] and ids the
            known = (neg >= 0) & (neg < n_map)                                 # catalog does not hold -> row 0 ("no negative")
            rows = torch.where(known, self.id_to_row[neg.clamp(0, n_map - 1)], torch.zeros_like(neg))
            out["hard_negatives"] = [_take(self.catalog, rows[:, n].contiguous()) for n in range(rows.shape[1])]
        return out

    def epoch(self, batch_size: int, shuffle: bool = True, generator: Optional[torch.Generator] = None,
              drop_last: bool = False):
        """Iterate one epoch of batches (the role of CombinedTwoTowerDataLoader.__iter__)."""
        order = torch.randperm(self.n, generator=generator) if shuffle else torch.arange(self.n)
        order = order.to(self.device)
        stop = self.n - (self.n % batch_size if drop_last else 0)
        for s in range(0, stop, batch_size):
            yield self.batch(order[s:s + batch_size])
