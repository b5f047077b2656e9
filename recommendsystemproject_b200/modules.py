"""Drop-in model classes for the reference's two-tower DSSM (same class names,
constructor signatures, attribute names and therefore ``state_dict`` keys),
running on hand-written sm_100a kernels.

Reference classes mirrored (paths relative to the reference root):
  MLP_Tower                 project/models/TwoTower/Tower.py:5-41
  SequenceFeatureProcessor  project/utils/SequenceFeatureProcessor.py:5-85
  SequenceEncoder           project/models/TwoTower/SequenceEncoder.py:5-74
  GenericTower              project/models/TwoTower/GenericTower.py:7-237
  TwoTowerModel             project/models/TwoTower/TwoTowerModel.py:5-150

What runs where: embedding gathers + pooling, their sparse backward, the
in-batch softmax CE (forward + backward) and retrieval top-K are libtt_b200
kernels; the small dense GEMMs / BatchNorm / Transformer blocks are cuBLAS /
cuDNN calls through torch ("torch GEMM acceptable
initially").  Modules are constructed on the CPU exactly like the reference
(same torch init calls in the same order, so the same seed gives the same
weights) and must be moved to a CUDA device before ``forward``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import TTError


import os as _os
# developer knobs: TT_BN_FUSED=0 sends every BatchNorm through torch; TT_BN_FUSED_MIN_ROWS = smallest batch for the library path
_BN_FUSED_DEFAULT = _os.environ.get("TT_BN_FUSED", "1") == "1"
_BN_FUSED_MIN_ROWS = int(_os.environ.get("TT_BN_FUSED_MIN_ROWS", "1"))


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise TTError(f"{what}: recommendsystemproject_b200 modules run on CUDA (B200) only; "
                      "move the model and the batch to the device (there is no CPU fallback)")


def grouped_batch_norm(bn: nn.BatchNorm1d, x: torch.Tensor, groups: int, relu: bool = False, dropout_p: float = 0.0,
                       seed: Optional[torch.Tensor] = None, call_id: int = 0) -> torch.Tensor:
    """dropout(relu(BatchNorm1d(x))) applied to `groups` independent slabs in ONE call (relu / dropout optional: the
    Linear -> BatchNorm1d -> ReLU -> Dropout block of Tower.py:16-21).  Rows of x are ordered (sample, slab), so
    x.view(B, groups*C) turns "statistics per slab" into plain per-channel statistics over B rows: numerically the same
    as the reference's one-item-tower-pass-per-hard-negative-slab (TwoTowerModel.py:54-60: every pass normalises with its
    own batch statistics and updates the running statistics once).  Running statistics receive the 1+N updates in slab
    order:  r <- (1-m)^G r + m * sum_g (1-m)^(G-1-g) stat_g .
    Training mode on the GPU runs the library kernels (ops.batch_norm_act: statistics, normalise + ReLU + dropout, and
    their backward; statistics span all ranks when the layer is marked `_tt_sync`); eval mode / odd shapes use torch."""
    rows, C = x.shape
    fused = (bn.training and x.is_cuda and x.dtype == torch.float32 and bn.affine and bn.track_running_stats
             and bn.momentum is not None and C % 4 == 0 and getattr(bn, "_tt_fused", _BN_FUSED_DEFAULT)
             and (rows // groups) >= _BN_FUSED_MIN_ROWS)
    if fused:
        B = rows // groups
        y, mean, var_u = ops.batch_norm_act(x.contiguous().view(B, groups * C), bn.weight, bn.bias, bn.running_mean,
                                            bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps, C, relu,
                                            dropout_p if bn.training else 0.0, seed, call_id,
                                            sync=getattr(bn, "_tt_sync", None))
        if groups > 1:
            _grouped_running_update(bn, mean, var_u, groups, C, x)
        return y.view(rows, C)
    if groups == 1 or not bn.training:
        y = bn(x)
    else:
        B = rows // groups
        w = bn.weight.repeat(groups)
        b = bn.bias.repeat(groups)
        mean = x.new_zeros(groups * C)
        var = x.new_ones(groups * C)
        y = F.batch_norm(x.view(B, groups * C), mean, var, w, b, True, 1.0, bn.eps)   # momentum 1: mean/var = batch stats
        if bn.track_running_stats:
            _grouped_running_update(bn, mean, var, groups, C, x)
        y = y.view(rows, C)
    if relu:
        y = F.relu(y)
    if dropout_p > 0.0 and bn.training:
        y = F.dropout(y, dropout_p, True)
    return y


def _grouped_running_update(bn, mean, var, groups, C, x):
    with torch.no_grad():
        m = bn.momentum
        cache = bn.__dict__.setdefault("_tt_group_coef", {})   # built once (eagerly), reused under graph capture
        key = (groups, x.device, x.dtype)
        if key not in cache:
            cache[key] = torch.tensor([m * (1.0 - m) ** (groups - 1 - g) for g in range(groups)], dtype=x.dtype,
                                      device=x.device)
        coef = cache[key]
        decay = (1.0 - m) ** groups
        bn.running_mean.mul_(decay).add_(coef @ mean.view(groups, C))
        bn.running_var.mul_(decay).add_(coef @ var.view(groups, C))
        bn.num_batches_tracked.add_(groups)


class TTLinear(nn.Linear):
    """nn.Linear (same parameters, same init, same checkpoint keys) whose backward computes dW and db in one pass over
    the rows with the library kernel (ops.LinearFn) instead of cuBLAS' no-split-K "nt" GEMM + a bias reduction."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled():
            return ops.linear(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


class MLP_Tower(nn.Module):
    """[Linear -> BatchNorm1d -> ReLU -> Dropout] x len(hidden) -> Linear -> L2 normalise
    (Tower.py:9-41)."""

    def __init__(self, input_dim, hidden_dims, output_dim, dropout=0.1):
        super().__init__()
        layers = []
        curr = input_dim
        for h in hidden_dims:
            layers += [TTLinear(curr, h), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(dropout)]
            curr = h
        layers.append(TTLinear(curr, output_dim))
        self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        # dropout seed of the fused BatchNorm + ReLU + Dropout kernels (device resident, bumped per training forward:
        # a CUDA-graph replay draws new masks); own generator, the global stream the reference's init consumes is untouched
        gen = torch.Generator().manual_seed((torch.initial_seed() + 7919 * input_dim) % (2 ** 63))
        self.register_buffer("_drop_seed", torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, generator=gen, device="cpu"), persistent=False)

    def _init_weights(self, m):  # Tower.py:28-35
        if isinstance(m, nn.Linear):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def forward(self, x, groups: int = 1):
        layers = list(self.mlp)
        n = len(layers)
        seed = None
        if self.training and x.is_cuda and any(isinstance(l, nn.Dropout) and l.p > 0 for l in layers):
            self._drop_seed.add_(1)
            seed = self._drop_seed.clone()
        i = 0
        while i < n:
            layer = layers[i]
            if isinstance(layer, nn.BatchNorm1d):
                relu = i + 1 < n and isinstance(layers[i + 1], nn.ReLU)
                has_drop = relu and i + 2 < n and isinstance(layers[i + 2], nn.Dropout)
                p = layers[i + 2].p if (has_drop and self.training) else 0.0
                x = grouped_batch_norm(layer, x, groups, relu=relu, dropout_p=p, seed=seed, call_id=i)
                i += 1 + (1 if relu else 0) + (1 if has_drop else 0)
            else:
                x = layer(x)
                i += 1
        return ops.l2_normalize(x)


class SequenceFeatureProcessor(nn.Module):
    """Per-position sequence feature embedder (SequenceFeatureProcessor.py:6-85)."""

    def __init__(self, feature_config_list, target_dim, max_seq_len, dropout=0.1):
        super().__init__()
        self.feature_config_list = feature_config_list
        self.target_dim = target_dim
        self.dropout = dropout
        self.embeddings = nn.ModuleDict()
        total = 0
        for feat in feature_config_list:
            # NB: the reference reads 'padding_index' (sic), so YAML 'padding_idx' is ignored here
            self.embeddings[feat["name"]] = nn.Embedding(feat["vocab_size"], feat["embedding_dim"],
                                                         padding_idx=feat.get("padding_index", 0))
            total += feat["embedding_dim"]
        self.feature_projection = nn.Sequential(TTLinear(total, target_dim), nn.Dropout(dropout))
        self.pos_emb = nn.Embedding(max_seq_len, target_dim)
        self.sparse_sink: Optional[ops.SparseGradSink] = None
        # one device flag word per feature: the gather kernel ORs 1 into it when it meets an id outside [0, vocab)
        # (read at TwoTowerModel.check_nan_flags: the reference raises IndexError there, SURVEY 8b "errors")
        self._oob_names = [(f["name"], f["vocab_size"]) for f in feature_config_list]
        self.register_buffer("_oob_flags", torch.zeros(max(1, len(feature_config_list)), dtype=torch.int32), persistent=False)

    def forward(self, input_dict):
        specs, tables = [], []
        shape = None
        for fi, feat in enumerate(self.feature_config_list):
            name = feat["name"]
            if name not in input_dict:
                print(f"Configuration Error: Unable to find {name} in the input dictionary, {name} has skipped")
                continue
            x = input_dict[name]
            _need_cuda(x, "SequenceFeatureProcessor")
            pad = feat.get("padding_index", 0)
            if x.dim() == 2:
                ids, mode = x.reshape(-1, 1), ops.POOL_NONE
            elif x.dim() == 3:
                pooling = feat.get("pooling", None)
                if pooling not in ("mean", "sum"):
                    raise ValueError(f"sequence feature {name}: [B, L, Tags] input needs pooling 'mean' or 'sum'")
                ids, mode = x.reshape(-1, x.shape[2]), ops.POOL_MODES[pooling]
            else:
                raise ValueError(f"sequence feature {name}: expected [B, L] or [B, L, Tags] ids")
            shape = x.shape[:2]
            specs.append((ids.contiguous(), mode, pad, False, self._oob_flags[fi:fi + 1]))
            tables.append(self.embeddings[name].weight)
        if not specs:
            raise ValueError("Configuration Error: No valid features were processed!")
        concat = ops.MultiGatherPool.apply(specs, self.sparse_sink, *tables)  # [B*L, sum d]
        total = self.feature_projection(concat).reshape(shape[0], shape[1], self.target_dim)
        total = total + self.pos_emb.weight[: shape[1]].unsqueeze(0)
        return F.dropout(total, p=self.dropout, training=self.training)


class SequenceEncoder(nn.Module):
    """Transformer behaviour encoder + last-valid gather (SequenceEncoder.py:6-74)."""

    def __init__(self, feature_config_list, model_dim=64, dim_feedforward=4 * 64, max_seq_len=20, n_head=4,
                 n_layers=1, dropout=0.1):
        super().__init__()
        self.feature_embedder = SequenceFeatureProcessor(feature_config_list, model_dim, max_seq_len, dropout=dropout)
        layer = nn.TransformerEncoderLayer(d_model=model_dim, nhead=n_head, dim_feedforward=dim_feedforward,
                                           dropout=dropout, batch_first=True)
        self.transformer_backbone = nn.TransformerEncoder(layer, num_layers=n_layers, enable_nested_tensor=False)
        # The layers keep torch's parameter names (checkpoints of the reference load unchanged) but, on the GPU, run
        # through two fused kernels + cuBLAS GEMMs instead of torch's ~55 launches per layer (`fused_layers`).
        self.fused_layers = True
        self.dropout_p = float(dropout)
        self.n_head = int(n_head)
        # dropout seed of the fused kernels: lives on the device and is bumped once per training forward, so a CUDA
        # graph replay draws new masks every step; reproducible under torch.manual_seed
        # (own generator seeded from torch.initial_seed(): the global stream the reference's init consumes is untouched)
        gen = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 63))
        self.register_buffer("_drop_seed", torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, generator=gen, device="cpu"), persistent=False)

    def _fused_ok(self, x):
        d = x.shape[-1]
        dh = d // self.n_head
        first = self.transformer_backbone.layers[0]
        return (self.fused_layers and x.is_cuda and x.dtype == torch.float32 and x.shape[1] <= 32 and dh in (8, 16, 32)
                and d % 32 == 0 and d <= 256 and not first.norm_first and self.transformer_backbone.norm is None
                and first.activation_relu_or_gelu == 1)

    def _fused_backbone(self, x, padding_mask):
        """nn.TransformerEncoder.forward(src, src_key_padding_mask) for post-norm ReLU layers (SequenceEncoder.py:60):
        x = norm1(x + dropout1(out_proj(attn(in_proj(x)))));  x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))."""
        p = self.dropout_p if self.training else 0.0
        seed = None
        if p > 0.0:
            self._drop_seed.add_(1)
            # a private copy per forward: the kernels' backward re-reads it, and a second forward before that backward
            # (hard-negative slabs one pass at a time) must neither change it nor trip autograd's version check
            seed = self._drop_seed.clone()
        pad = padding_mask.to(torch.uint8).contiguous()
        for li, layer in enumerate(self.transformer_backbone.layers):
            sa = layer.self_attn
            lin = ops.linear if torch.is_grad_enabled() else F.linear
            qkv = lin(x, sa.in_proj_weight, sa.in_proj_bias)
            a = ops.attn_small(qkv, pad, sa.num_heads, p, seed, 3 * li)
            a = lin(a, sa.out_proj.weight, sa.out_proj.bias)
            x = ops.add_dropout_layer_norm(x, a, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, p, seed, 3 * li + 1)
            f = F.relu(lin(x, layer.linear1.weight, layer.linear1.bias))
            if p > 0.0:
                f = F.dropout(f, p, True)
            f = lin(f, layer.linear2.weight, layer.linear2.bias)
            x = ops.add_dropout_layer_norm(x, f, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, p, seed, 3 * li + 2)
        return x

    def forward(self, input_dict):
        main = self.feature_embedder.feature_config_list[0]
        main_seq = input_dict[main["name"]]
        padding_mask = main_seq == main.get("padding_index", 0)
        # rows that are all padding get their LAST position unmasked (SequenceEncoder.py:43-46);
        # written without the reference's `.any()` host sync
        all_pad = padding_mask.all(dim=1)
        padding_mask = padding_mask.clone()
        padding_mask[:, -1] &= ~all_pad
        seq_emb = self.feature_embedder(input_dict)
        if self._fused_ok(seq_emb):
            ctx = self._fused_backbone(seq_emb, padding_mask)
        else:
            ctx = self.transformer_backbone(src=seq_emb, src_key_padding_mask=padding_mask)
        return self._gather_last_valid(ctx, padding_mask)

    def _gather_last_valid(self, seq_output, padding_mask):
        B = seq_output.shape[0]
        idx = ((~padding_mask).long().sum(dim=1) - 1).clamp(min=0)  # assumes right padding, like the reference
        return seq_output[torch.arange(B, device=seq_output.device), idx]


def _dist_rank_world(cfg_model):
    """(rank, world) for row-sharded features: two_tower.sharding = {rank, world} when given, else torch.distributed."""
    sh = cfg_model.get("sharding") or {}
    if "world" in sh:
        return int(sh.get("rank", 0)), int(sh["world"])
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class ShardedEmbedding(nn.Module):
    """The local shard of a row-sharded table (owner = row % world, local row = row // world): what
    ``nn.Embedding(vocab, dim, padding_idx)`` + ``xavier_uniform_`` of GenericTower.py:45-51 becomes for a sparse feature
    whose YAML entry says ``row_sharded: true`` (BASELINE configs[2]: 100M users / 10M items x 128).

    * ``weight`` is the [ceil((vocab - rank) / world), dim] shard, initialised uniform(+-sqrt(6 / (vocab + dim))) -- the
      xavier bound of the WHOLE table, as the reference computes it; construct under ``with torch.device("cuda")`` for
      tables that should never exist on the host.
    * the pad row (the reference leaves it non-zero and frozen) is a replicated fp32 buffer; DataParallel-style setups
      broadcast it from rank 0 (dist.ShardedTrainStep does).
    * ``state_dict`` keeps the reference's key and, with ``gather_on_save`` (default), its [vocab, dim] shape: the
      shards are all-gathered on save (collective!) and the full table is sliced on load -- a checkpoint written by the
      reference loads here and vice versa.  ``gather_on_save = False`` stores / expects the shard."""

    def __init__(self, vocab, dim, padding_idx, rank, world, dtype=torch.float32):
        super().__init__()
        self.num_embeddings, self.embedding_dim, self.padding_idx = int(vocab), int(dim), padding_idx
        self.rank, self.world = int(rank), int(world)
        local = (self.num_embeddings - self.rank + self.world - 1) // self.world
        bound = (6.0 / (self.num_embeddings + self.embedding_dim)) ** 0.5
        self.weight = nn.Parameter(torch.empty(local, dim, dtype=dtype))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
        pad = torch.empty(dim, dtype=torch.float32, device=self.weight.device).uniform_(-bound, bound)
        self.register_buffer("pad_row", pad if padding_idx is not None else None, persistent=False)
        self.gather_on_save = True
        self.group = None            # sharded.ShardedTableGroup, attached by the tower / model
        self.table_name = None
        self._register_state_dict_hook(self._save_hook)
        self._register_load_state_dict_pre_hook(self._load_hook)

    @staticmethod
    def _save_hook(module, state_dict, prefix, local_metadata):
        if module.gather_on_save and module.group is not None:
            state_dict[prefix + "weight"] = module.group.gather_full_weight(module.table_name)

    def _load_hook(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is not None and w.shape[0] == self.num_embeddings and self.num_embeddings != self.weight.shape[0]:
            if self.padding_idx is not None and self.pad_row is not None:
                with torch.no_grad():
                    self.pad_row.copy_(w[self.padding_idx].to(self.pad_row.device, torch.float32))
            state_dict[key] = w[self.rank::self.world]
        elif w is not None and self.world == 1 and self.padding_idx is not None and self.pad_row is not None:
            with torch.no_grad():
                self.pad_row.copy_(w[self.padding_idx].to(self.pad_row.device, torch.float32))


class GenericTower(nn.Module):
    """YAML-driven tower (GenericTower.py:9-237): sparse / pooled / dense /
    sequence features -> concat -> BatchNorm1d -> MLP_Tower."""

    def __init__(self, cfg, tower_name):
        super().__init__()
        model_cfg = cfg.get("two_tower", {})
        if len(model_cfg.get(tower_name, {})) == 0:
            raise ValueError(f"TwoTower Model initializing failed, {tower_name} has no features")
        tcfg = model_cfg.get(tower_name)
        hidden = tcfg["mlp_hidden_dim"]
        out_dims = tcfg["output_dims"]
        dropout = tcfg["dropout"]
        self.tower_embedding_dim = tcfg["embedding_dim"]
        self.embeddings = nn.ModuleDict()
        self.pooling_config = {}
        self.sparse_features = tcfg.get("sparse_features", None)
        self.dense_features = tcfg.get("dense_features", None)
        self.seq_features = tcfg.get("sequence_features", None)
        self.sparse_sink: Optional[ops.SparseGradSink] = None  # set by optim.FusedTwoTowerOptimizer
        self.sparse_grad_tables = set()
        self.tower_name = tower_name
        self.shard_group = None     # sharded.ShardedTableGroup serving this tower's `row_sharded: true` features
        self.sharded_features = []

        sparse_total = 0
        if self.sparse_features is not None:
            for feat in self.sparse_features:
                for field in self.sparse_features:
                    missing = [k for k in ("name", "vocab_size", "embedding_dim") if k not in field]
                    if missing:
                        raise ValueError(f"Sparse feature config missing keys {missing}: {feat}")
                name = feat["name"]
                if feat.get("row_sharded", False):
                    # extension: this table is row-sharded over the ranks; only the local shard exists here
                    rank, world = _dist_rank_world(model_cfg)
                    self.embeddings[name] = ShardedEmbedding(feat["vocab_size"], feat["embedding_dim"],
                                                             feat.get("padding_idx", 0), rank, world)
                    self.sharded_features.append(name)
                else:
                    self.embeddings[name] = nn.Embedding(feat["vocab_size"], feat["embedding_dim"],
                                                         padding_idx=feat.get("padding_idx", 0))
                    nn.init.xavier_uniform_(self.embeddings[name].weight)  # overwrites the zeroed pad row (:51)
                if "pooling" in feat:
                    self.pooling_config[name] = feat["pooling"]
                sparse_total += feat["embedding_dim"]
        dense_total = 0
        if self.dense_features is not None:
            for feat in self.dense_features:
                for field in self.dense_features:
                    missing = [k for k in ("name", "dim", "embedding_dim") if k not in field]
                    if missing:
                        raise ValueError(f"Dense feature config missing keys {missing}: {field}, tower initializing failed")
                self.embeddings[feat["name"]] = nn.Sequential(TTLinear(feat["dim"], feat["embedding_dim"]))
                dense_total += feat["embedding_dim"]
        seq_total = 0
        # the reference leaves self.seq_encoder undefined for `sequence_features: []`
        # (AttributeError in forward); fixed here: None.
        self.seq_encoder = None
        if self.seq_features is not None and len(self.seq_features) > 0:
            model_dim = tcfg.get("embedding_dim", 32)
            tp = tcfg.get("transformer_parameters", {})
            n_head = tp.get("n_head", 4)
            if model_dim % n_head != 0:
                raise ValueError(f"Transformer initializing failed, embedding dim {model_dim} must be divisible by n_head {n_head}")
            self.seq_encoder = SequenceEncoder(feature_config_list=self.seq_features, model_dim=model_dim,
                                               dim_feedforward=tp.get("FFN_dim", 4 * model_dim),
                                               max_seq_len=tp.get("max_seq_len", 20), n_head=n_head,
                                               n_layers=tp.get("n_layers", 1), dropout=tp.get("dropout", 0.1))
            seq_total = model_dim
        self._oob_names = [(f["name"], f["vocab_size"]) for f in (self.sparse_features or [])]
        self.register_buffer("_oob_flags", torch.zeros(max(1, len(self._oob_names)), dtype=torch.int32), persistent=False)
        self.total_embed_dim = sparse_total + dense_total + seq_total
        self.feature_bn = nn.BatchNorm1d(self.total_embed_dim)
        self.mlp = MLP_Tower(input_dim=self.total_embed_dim, hidden_dims=hidden, output_dim=out_dims, dropout=dropout)
        if self.sharded_features:
            from . import sharded
            rank, world = _dist_rank_world(model_cfg)
            self.attach_shard_group(sharded.ShardedTableGroup(rank, world,
                                                              capacity_factor=float((model_cfg.get("sharding") or {}).get("capacity_factor", 1.25))))

    def attach_shard_group(self, group):
        """Register this tower's row-sharded tables with `group` (TwoTowerModel passes ONE group to both towers so that
        a step has a single batched exchange)."""
        self.shard_group = group
        for feat in self.sparse_features or []:
            name = feat["name"]
            if name not in self.sharded_features:
                continue
            emb = self.embeddings[name]
            if "pooling" in feat:
                if feat["pooling"] not in ("mean", "sum"):
                    raise ValueError(f"row-sharded feature {name}: pooling must be 'mean' or 'sum'")
                mode = ops.POOL_MODES[feat["pooling"]]
            else:
                mode = ops.POOL_NONE
            emb.group, emb.table_name = group, f"{self.tower_name}.{name}"
            group.add_table(emb.table_name, emb.num_embeddings, emb.embedding_dim, mode, emb.padding_idx, None, None, holder=emb)

    def sharded_ids(self, input_dict, mapping):
        """{group table name: ids [B, L]} of this tower's row-sharded features."""
        out = {}
        sparse_matrix = input_dict.get("sparse")
        seq_dict = input_dict.get("sequence", {}) or {}
        for feat in self.sparse_features or []:
            name = feat["name"]
            if name not in self.sharded_features:
                continue
            if "pooling" in feat:
                if name not in seq_dict:
                    raise ValueError(f"Pooled feature {name} missing from sequence dict")
                ids = seq_dict[name]
            else:
                if mapping and "sparse" in mapping:
                    col = mapping["sparse"].get(name)
                    if col is None:
                        raise ValueError(f"Feature '{name}' not found in column mapping")
                else:
                    col = [f["name"] for f in self.sparse_features if "pooling" not in f].index(name)
                ids = sparse_matrix[:, col]
            _need_cuda(ids, "GenericTower")
            out[f"{self.tower_name}.{name}"] = ids
        return out

    # ------------------------------------------------------------------
    def _sparse_specs(self, input_dict, mapping):
        specs, tables = [], []
        sparse_matrix = input_dict["sparse"]
        seq_dict = input_dict.get("sequence", {})
        for fi, feat in enumerate(self.sparse_features):
            name = feat["name"]
            if name in self.sharded_features:
                continue
            pad = feat.get("padding_idx", 0)
            if "pooling" in feat:
                if name not in seq_dict:
                    print(f"Warning: Pooled feature {name} missing from sequence dict")
                    continue
                ids = seq_dict[name]
                if ids.dim() == 1:
                    ids = ids.unsqueeze(1)
                pooling = self.pooling_config[name]
                if pooling not in ("mean", "sum", "max"):
                    raise ValueError(f"pooled feature {name}: unknown pooling '{pooling}'")
                mode = ops.POOL_MODES[pooling]
            else:
                if sparse_matrix is None:
                    continue
                if mapping and "sparse" in mapping:
                    col = mapping["sparse"].get(name)
                    if col is None:
                        raise ValueError(f"Feature '{name}' not found in column mapping")
                else:
                    col = [f["name"] for f in self.sparse_features if "pooling" not in f].index(name)
                ids = sparse_matrix[:, col].unsqueeze(1)
                mode = ops.POOL_NONE
            _need_cuda(ids, "GenericTower")
            specs.append((ids.contiguous().long(), mode, pad, name in self.sparse_grad_tables, self._oob_flags[fi:fi + 1]))
            tables.append(self.embeddings[name].weight)
        return specs, tables

    def forward_grouped(self, input_dicts, feature_column_mapping=None):
        """The tower over G input dicts of B rows each in ONE pass -> [B, G, D]; every dict keeps its own BatchNorm
        batch statistics (see grouped_batch_norm).  Used for [positive items] + hard-negative slabs."""
        G = len(input_dicts)
        first = input_dicts[0]

        def stack(ts):
            t = torch.stack(ts, dim=1)                      # [B, G, ...]: row order (sample, slab)
            return t.reshape(t.shape[0] * G, *t.shape[2:])

        merged = {}
        for key in ("sparse", "dense"):
            if key in first and first[key] is not None:
                merged[key] = stack([d[key] for d in input_dicts])
        if "sequence" in first and first["sequence"]:
            merged["sequence"] = {n: stack([d["sequence"][n] for d in input_dicts]) for n in first["sequence"]}
        out = self.forward(merged, feature_column_mapping, groups=G)
        return out.view(out.shape[0] // G, G, out.shape[1])

    def forward(self, input_dict, feature_column_mapping=None, groups: int = 1, sharded_vecs=None):
        feats = []
        if self.sparse_features and "sparse" in input_dict:
            specs, tables = self._sparse_specs(input_dict, feature_column_mapping)
            local = ops.MultiGatherPool.apply(specs, self.sparse_sink, *tables) if specs else None
            if not self.sharded_features:
                if local is not None:
                    feats.append(local)
            else:
                # row-sharded features come from the group's batched exchange (done once for both towers by
                # TwoTowerModel; a tower used on its own does its own); concat order = YAML order (GenericTower.py:131-233)
                if sharded_vecs is None:
                    sharded_vecs = self.shard_group.lookup(self.sharded_ids(input_dict, feature_column_mapping))
                col = 0
                k = 0
                for feat in self.sparse_features:
                    name = feat["name"]
                    if name in self.sharded_features:
                        feats.append(sharded_vecs[f"{self.tower_name}.{name}"])
                    elif k < len(tables) and tables[k] is self.embeddings[name].weight:
                        d = tables[k].shape[1]
                        feats.append(local[:, col:col + d])
                        col += d
                        k += 1
        if self.dense_features and "dense" in input_dict:
            dense = input_dict["dense"]
            for feat in self.dense_features:
                name = feat["name"]
                if feature_column_mapping and "dense" in feature_column_mapping:
                    col = feature_column_mapping["dense"].get(name)
                    if col is None:
                        raise ValueError(f"Dense feature '{name}' not found in column mapping")
                else:
                    col = [f["name"] for f in self.dense_features].index(name)
                x = dense[:, col:col + 1]
                if x.dtype != torch.float32:
                    x = x.float()
                feats.append(self.embeddings[name](x))
        if self.seq_encoder is not None and "sequence" in input_dict:
            seq = input_dict["sequence"]
            if seq:
                feats.append(self.seq_encoder(seq))
        if not feats:
            raise RuntimeError("Tower received no valid features. Check if input_dict matches config")
        x = feats[0] if len(feats) == 1 else torch.cat(feats, dim=1)
        x = grouped_batch_norm(self.feature_bn, x, groups)
        return self.mlp(x, groups)


def _on_cuda(obj) -> bool:
    """True when the first tensor found in a batch tree lives on a CUDA device (CPU batches fall through to the
    ordinary path, whose ops raise TTError: there is no CPU fallback)."""
    if isinstance(obj, torch.Tensor):
        return obj.is_cuda
    if isinstance(obj, dict):
        for v in obj.values():
            r = _on_cuda(v)
            if r is not None:
                return r
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            r = _on_cuda(v)
            if r is not None:
                return r
    return None


def _same_layout(item_dict, negs) -> bool:
    """True when every hard-negative slab has the item batch's keys and shapes (so they can be stacked)."""
    def sig(d):
        out = []
        for key in ("sparse", "dense"):
            t = d.get(key)
            out.append(None if t is None else tuple(t.shape))
        seq = d.get("sequence") or {}
        out.append(tuple(sorted((n, tuple(t.shape)) for n, t in seq.items())))
        return out
    ref = sig(item_dict)
    return all(sig(n) == ref for n in negs)


class TwoTowerModel(nn.Module):
    """Two-tower wrapper + fused in-batch softmax loss (TwoTowerModel.py:6-150)."""

    def __init__(self, user_tower, item_tower, user_feature_mapping=None, item_feature_mapping=None):
        super().__init__()
        self.user_tower = user_tower
        self.item_tower = item_tower
        self.user_feature_mapping = user_feature_mapping
        self.item_feature_mapping = item_feature_mapping
        # the reference does three isnan().any() host syncs per loss call; here the kernels
        # set a device flag that is read once (or never, when strict_nan_check is False /
        # the stream is being captured into a CUDA graph)
        self.strict_nan_check = True
        self.last_nan_flags: Optional[torch.Tensor] = None
        # positives + hard-negative slabs through the item tower in one grouped pass (same numbers, 1/(1+N) of the
        # launches); False = one pass per slab, op for op like the reference
        self.group_hard_negatives = True
        # the two towers are independent until the loss: run the item side on a second stream (forward, and therefore
        # its autograd backward too); inside a CUDA graph this becomes two concurrent branches
        self.parallel_towers = True
        self._side_streams = {}
        # row-sharded tables of BOTH towers share one group => one batched exchange per step
        self.shard_group = None
        groups = [t.shard_group for t in (user_tower, item_tower) if getattr(t, "shard_group", None) is not None]
        if groups:
            from . import sharded
            g0 = groups[0]
            merged = sharded.ShardedTableGroup(g0.rank, g0.world, capacity_factor=g0.capacity_factor, dev_ops=g0.ops, group=g0.pg)
            for t in (user_tower, item_tower):
                if getattr(t, "shard_group", None) is not None:
                    t.attach_shard_group(merged)
            self.shard_group = merged
        # loss kernel: "auto" = tcgen05 bf16 path when the shapes allow it and the batch is large, else exact fp32
        self.loss_precision = "auto"
        self.loss_tc_min_batch = 4096
        # tensor-core loss in training: forward + dU in ONE walk over the logit tiles (ops.FusedInBatchCE single_pass).
        # "auto" = whenever it applies (no per-row hard negatives, gradients on, temperature >= 0.0155: the towers
        # L2-normalise, so |logit| <= 1/T); embeddings that break the kernel's range raise its device flag and the loss
        # is recomputed by the three-pass kernels (and single-pass switched off) at the next flag check
        self.loss_single_pass = "auto"

    def set_feature_mappings(self, user_mapping, item_mapping):
        self.user_feature_mapping = user_mapping
        self.item_feature_mapping = item_mapping

    def _sharded_prefetch(self, batch_data):
        """The batched lookup of every row-sharded feature of both towers (None when there is nothing to prefetch:
        no sharded tables, or hard-negative slabs -- those go through the towers' own lookups)."""
        if self.shard_group is None or batch_data.get("hard_negatives"):
            return None
        ids = {}
        ids.update(self.user_tower.sharded_ids(batch_data["user_tower"], self.user_feature_mapping))
        ids.update(self.item_tower.sharded_ids(batch_data["item_tower"], self.item_feature_mapping))
        return self.shard_group.lookup(ids) if ids else None

    def forward(self, batch_data):
        vecs = self._sharded_prefetch(batch_data)
        if self.parallel_towers and torch.is_grad_enabled() and self.training and _on_cuda(batch_data):
            return self._forward_two_streams(batch_data, vecs)
        user_emb = self.user_tower(batch_data["user_tower"], self.user_feature_mapping, sharded_vecs=vecs)
        item_emb, hard_neg_emb = self._item_side(batch_data, vecs)
        return user_emb, item_emb, hard_neg_emb

    def _forward_two_streams(self, batch_data, vecs=None):
        cur = torch.cuda.current_stream()
        key = cur.device.index
        if key not in self._side_streams:
            self._side_streams[key] = torch.cuda.Stream(device=cur.device)
        side = self._side_streams[key]
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            item_emb, hard_neg_emb = self._item_side(batch_data, vecs)
        user_emb = self.user_tower(batch_data["user_tower"], self.user_feature_mapping, sharded_vecs=vecs)
        cur.wait_stream(side)
        for t in (item_emb, hard_neg_emb):
            if t is not None:
                t.record_stream(cur)
        return user_emb, item_emb, hard_neg_emb

    def _item_side(self, batch_data, vecs=None):
        negs = batch_data.get("hard_negatives") or []
        if negs and self.group_hard_negatives and _same_layout(batch_data["item_tower"], negs):
            both = self.item_tower.forward_grouped([batch_data["item_tower"]] + list(negs), self.item_feature_mapping)
            return both[:, 0], both[:, 1:]
        item_emb = self.item_tower(batch_data["item_tower"], self.item_feature_mapping, sharded_vecs=vecs)
        hard_neg_emb = None
        if "hard_negatives" in batch_data and batch_data["hard_negatives"]:
            # one item-tower pass per slab: each slab keeps its own BatchNorm batch
            # statistics and running-stat update, like the reference (:54-60)
            slabs = [self.item_tower(neg, self.item_feature_mapping) for neg in batch_data["hard_negatives"]]
            hard_neg_emb = torch.stack(slabs, dim=1)
        return item_emb, hard_neg_emb

    def predict(self, batch_data):
        user_emb, item_emb, _ = self.forward(batch_data)
        return (user_emb * item_emb).sum(dim=1)

    def get_item_embeddings(self, item_inputs):
        return self.item_tower(item_inputs, self.item_feature_mapping)

    def check_nan_flags(self):
        """Raise the reference's RuntimeErrors from the device flag word, and IndexError for ids outside a table
        (nn.Embedding raises it on the CPU; here the gather kernels set a per-feature flag): one host sync."""
        if self.last_nan_flags is None:
            return
        owners = [m for m in self.modules() if isinstance(getattr(m, "_oob_flags", None), torch.Tensor) and m._oob_flags.is_cuda]
        words = torch.cat([self.last_nan_flags.reshape(-1)] + [m._oob_flags for m in owners]).tolist()
        flags = int(words[0])
        pos = 1
        for m in owners:
            for k, (name, vocab) in enumerate(m._oob_names):
                if words[pos + k]:
                    m._oob_flags.zero_()
                    raise IndexError(f"index out of range in feature '{name}': ids must lie in [0, {vocab - 1}] "
                                     f"(vocab_size {vocab})")
            pos += m._oob_flags.numel()
        if flags & 1:
            raise RuntimeError("Found NaN in User Embedding")
        if flags & 2:
            raise RuntimeError("Found NaN in Item Embedding")
        if flags & 4:
            raise RuntimeError("Found NaN in Hard Negative Embedding")
        if flags & ops.CE_FLAG_LOGIT_RANGE:
            raise RuntimeError("in-batch loss: |logit| exceeds the single-pass tensor-core kernel's range "
                               "(|u||i|/T * log2(e) > 96); set model.loss_single_pass = False")
        if self.shard_group is not None and self.shard_group.tables:
            self.shard_group.check_flags()

    def compute_loss(self, user_emb, item_emb, item_ids=None, hard_neg_emb=None, temperature=0.1,
                     hard_neg_pool=None):
        """In-batch (+ per-row hard negative [B,N,D], + shared pool [H,D]) softmax CE.
        ``hard_neg_pool`` is an extension: it equals ``hard_neg_emb = pool.expand(B,H,D)``."""
        _need_cuda(user_emb, "compute_loss")
        if hard_neg_emb is not None:
            assert hard_neg_emb.dim() == 3, f"Expected shape [B, N, D], got {hard_neg_emb.shape}"
            assert hard_neg_emb.size(0) == user_emb.shape[0], "Batch size mismatch"
        precision = self.loss_precision
        if precision == "auto":
            # the tensor-core kernel needs D in {64, 128}; below a few thousand rows the exact fp32 kernel is as fast
            precision = "bf16" if (user_emb.shape[1] in (64, 128) and user_emb.shape[0] >= self.loss_tc_min_batch) else "fp32"
        single = self.loss_single_pass
        if single == "auto":
            single = (precision == "bf16" and hard_neg_emb is None and torch.is_grad_enabled()
                      and ops.single_pass_ok(temperature))
        loss, _, flags = ops.fused_inbatch_ce(user_emb, item_emb, item_ids=item_ids, hn_rows=hard_neg_emb,
                                              pool=hard_neg_pool, temperature=temperature, precision=precision,
                                              single_pass=bool(single))
        self.last_nan_flags = flags
        if self.strict_nan_check and not torch.cuda.is_current_stream_capturing():
            if single and int(flags) & ops.CE_FLAG_LOGIT_RANGE:
                # un-normalised embeddings: outside the single-pass kernel's range -> the general three-pass kernels
                self.loss_single_pass = False
                loss, _, flags = ops.fused_inbatch_ce(user_emb, item_emb, item_ids=item_ids, hn_rows=hard_neg_emb,
                                                      pool=hard_neg_pool, temperature=temperature, precision=precision)
                self.last_nan_flags = flags
            self.check_nan_flags()
        return loss
