"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the two-tower DSSM hot path.

A functional (no nn.Module) torch-fp32 / numpy restatement of what the
reference computes through PyTorch defaults.  Every function cites the
reference file:line it follows (paths relative to the reference root).

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference modules, runs them on seeded synthetic inputs and commits the
input/output vectors under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against them (plus KATs 1-6 of SURVEY.md section 4).

State is a flat ``dict[str, Tensor]`` keyed exactly like the reference
``TwoTowerModel.state_dict()``, so the same dict
loads into the reference modules, this oracle and the B200 drop-in modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
State = Dict[str, Tensor]

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LN_EPS = 1e-5
MASK_VALUE = -1e9  # TwoTowerModel.py:114 (literal, applied after /T)


# --------------------------------------------------------------------------
# a2/a3: embedding lookups (GenericTower.py:141-183)
# --------------------------------------------------------------------------
def embedding_lookup(weight: Tensor, ids: Tensor) -> Tensor:
    """W[ids] for any id shape (GenericTower.py:182, SequenceFeatureProcessor.py:60)."""
    return weight.index_select(0, ids.reshape(-1)).reshape(*ids.shape, weight.shape[1])


def pooled_lookup(weight: Tensor, ids: Tensor, mode: str) -> Tensor:
    """pool_{l<L} W[ids[b,l]] over ALL L positions, pads included
    (GenericTower.py:148-160): mean divides by L, not by the valid count."""
    if ids.dim() == 1:
        ids = ids.unsqueeze(1)  # GenericTower.py:150-151
    e = embedding_lookup(weight, ids)  # [B, L, D]
    if mode == "mean":
        return e.sum(dim=1) / ids.shape[1]
    if mode == "sum":
        return e.sum(dim=1)
    if mode == "max":
        return e.max(dim=1)[0]
    return e  # unknown pooling string: the reference leaves [B,L,D] untouched


def embedding_grad_dense(ids: Tensor, grad_rows: Tensor, vocab: int,
                         padding_idx: Optional[int]) -> Tensor:
    """dW[r] += sum over positions p with ids[p]==r, r != padding_idx
    (torch embedding_dense_backward semantics).  Accumulates in
    float64 so the oracle is order-independent."""
    flat = ids.reshape(-1)
    g = grad_rows.reshape(flat.shape[0], -1).double()
    if padding_idx is not None:
        keep = flat != padding_idx
        flat, g = flat[keep], g[keep]
    out = torch.zeros(vocab, g.shape[1], dtype=torch.float64)
    out.index_add_(0, flat, g)
    return out.float()


# --------------------------------------------------------------------------
# normalisation / linear building blocks (torch defaults selected by
# Tower.py:17-18, GenericTower.py:111, SequenceEncoder.py:17-23)
# --------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def batchnorm1d(x: Tensor, state: State, prefix: str, training: bool,
                update_running: bool = True) -> Tensor:
    """BatchNorm1d over [B, C] (GenericTower.py:234, Tower.py:18): batch stats
    with biased variance in train mode; running stats updated with the
    unbiased variance and momentum 0.1; eval uses running stats."""
    gamma, beta = state[prefix + ".weight"], state[prefix + ".bias"]
    if training:
        n = x.shape[0]
        mean = x.mean(dim=0)
        var_b = ((x - mean) ** 2).mean(dim=0)
        if update_running:
            with torch.no_grad():
                var_u = var_b * (n / max(n - 1, 1))
                state[prefix + ".running_mean"] = (
                    (1 - BN_MOMENTUM) * state[prefix + ".running_mean"] + BN_MOMENTUM * mean.detach())
                state[prefix + ".running_var"] = (
                    (1 - BN_MOMENTUM) * state[prefix + ".running_var"] + BN_MOMENTUM * var_u.detach())
                state[prefix + ".num_batches_tracked"] = state[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var_b = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    return (x - mean) / torch.sqrt(var_b + BN_EPS) * gamma + beta


def layernorm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def l2_normalize(x: Tensor) -> Tensor:
    """F.normalize(p=2, dim=1) (Tower.py:41): x / max(||x||, 1e-12)."""
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


# --------------------------------------------------------------------------
# a7: MLP tower (Tower.py:16-25,37-41).  Dropout is identity in the oracle:
# parity runs use dropout 0.0 / eval (SURVEY.md section 7 hard part 8).
# --------------------------------------------------------------------------
def mlp_tower(x: Tensor, state: State, prefix: str, n_hidden: int, training: bool,
              update_running: bool = True) -> Tensor:
    idx = 0
    for _ in range(n_hidden):
        x = linear(x, state[f"{prefix}.mlp.{idx}.weight"], state[f"{prefix}.mlp.{idx}.bias"])
        x = batchnorm1d(x, state, f"{prefix}.mlp.{idx + 1}", training, update_running)
        x = torch.relu(x)
        idx += 4  # Linear, BN, ReLU, Dropout
    x = linear(x, state[f"{prefix}.mlp.{idx}.weight"], state[f"{prefix}.mlp.{idx}.bias"])
    return l2_normalize(x)


# --------------------------------------------------------------------------
# a5: sequence feature embedder (SequenceFeatureProcessor.py:38-85)
# --------------------------------------------------------------------------
def seq_feature_processor(seq_inputs: Dict[str, Tensor], state: State, prefix: str,
                          seq_cfg: List[dict]) -> Tensor:
    parts = []
    for feat in seq_cfg:
        name = feat["name"]
        if name not in seq_inputs:
            continue  # SequenceFeatureProcessor.py:52-55
        ids = seq_inputs[name]
        e = embedding_lookup(state[f"{prefix}.embeddings.{name}.weight"], ids)
        if ids.dim() == 3:  # [B, L, Tags] -> pool over tags, pads included (:64-68)
            pooling = feat.get("pooling", None)
            if pooling == "mean":
                e = e.sum(dim=2) / ids.shape[2]
            elif pooling == "sum":
                e = e.sum(dim=2)
        parts.append(e)
    x = torch.cat(parts, dim=-1)
    x = linear(x, state[f"{prefix}.feature_projection.0.weight"],
               state[f"{prefix}.feature_projection.0.bias"])
    L = x.shape[1]
    return x + state[f"{prefix}.pos_emb.weight"][:L].unsqueeze(0)  # :79-82


# --------------------------------------------------------------------------
# a6: Transformer behaviour encoder (SequenceEncoder.py:32-74)
# --------------------------------------------------------------------------
def transformer_layer(x: Tensor, key_pad: Tensor, state: State, prefix: str, n_head: int) -> Tensor:
    """Post-norm encoder layer, ReLU FFN, key-padding mask as additive -inf."""
    B, L, d = x.shape
    hd = d // n_head
    qkv = linear(x, state[prefix + ".self_attn.in_proj_weight"], state[prefix + ".self_attn.in_proj_bias"])
    q, k, v = qkv.split(d, dim=-1)

    def heads(t):
        return t.reshape(B, L, n_head, hd).permute(0, 2, 1, 3)  # [B, h, L, hd]

    q, k, v = heads(q), heads(k), heads(v)
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)  # [B, h, Lq, Lk]
    att = att.masked_fill(key_pad[:, None, None, :], float("-inf"))
    att = torch.softmax(att, dim=-1)
    o = (att @ v).permute(0, 2, 1, 3).reshape(B, L, d)
    o = linear(o, state[prefix + ".self_attn.out_proj.weight"], state[prefix + ".self_attn.out_proj.bias"])
    x = layernorm(x + o, state[prefix + ".norm1.weight"], state[prefix + ".norm1.bias"])
    ff = linear(torch.relu(linear(x, state[prefix + ".linear1.weight"], state[prefix + ".linear1.bias"])),
                state[prefix + ".linear2.weight"], state[prefix + ".linear2.bias"])
    return layernorm(x + ff, state[prefix + ".norm2.weight"], state[prefix + ".norm2.bias"])


def sequence_padding_mask(main_ids: Tensor, padding_idx: int = 0) -> Tensor:
    """SequenceEncoder.py:40-46: mask = (first seq feature == pad); rows that
    are entirely padding get their LAST position unmasked."""
    if main_ids.dim() == 3:
        raise ValueError("the first sequence feature must be [B, L]")
    mask = main_ids == padding_idx
    all_pad = mask.all(dim=1)
    mask = mask.clone()
    mask[all_pad, -1] = False
    return mask


def gather_last_valid(seq_out: Tensor, key_pad: Tensor) -> Tensor:
    """SequenceEncoder.py:58-74: row index = (#valid - 1) clamped at 0; this
    assumes right padding (SURVEY.md section 7 hard part 6)."""
    idx = (~key_pad).long().sum(dim=1) - 1
    idx = idx.clamp_min(0)
    return seq_out[torch.arange(seq_out.shape[0]), idx]


def sequence_encoder(seq_inputs: Dict[str, Tensor], state: State, prefix: str, seq_cfg: List[dict],
                     n_head: int, n_layers: int) -> Tensor:
    main = seq_cfg[0]
    key_pad = sequence_padding_mask(seq_inputs[main["name"]], main.get("padding_index", 0))
    x = seq_feature_processor(seq_inputs, state, prefix + ".feature_embedder", seq_cfg)
    for i in range(n_layers):
        x = transformer_layer(x, key_pad, state, f"{prefix}.transformer_backbone.layers.{i}", n_head)
    return gather_last_valid(x, key_pad)


# --------------------------------------------------------------------------
# a2-a7: tower forward (GenericTower.py:120-237)
# --------------------------------------------------------------------------
def tower_forward(inputs: dict, mapping: Optional[dict], state: State, tower: str, cfg: dict,
                  training: bool, update_running: bool = True) -> Tensor:
    tcfg = cfg["two_tower"][tower]
    sparse_cfg = tcfg.get("sparse_features") or []
    dense_cfg = tcfg.get("dense_features") or []
    seq_cfg = tcfg.get("sequence_features")
    feats = []
    if sparse_cfg and "sparse" in inputs:
        seq_dict = inputs.get("sequence", {})
        for feat in sparse_cfg:
            name = feat["name"]
            w = state[f"{tower}.embeddings.{name}.weight"]
            if "pooling" in feat:
                if name not in seq_dict:
                    continue
                feats.append(pooled_lookup(w, seq_dict[name], feat["pooling"]))
            else:
                if mapping and "sparse" in mapping:
                    col = mapping["sparse"].get(name)
                    if col is None:
                        raise ValueError(f"Feature '{name}' not found in column mapping")
                else:
                    col = [f["name"] for f in sparse_cfg if "pooling" not in f].index(name)
                feats.append(embedding_lookup(w, inputs["sparse"][:, col]))
    if dense_cfg and "dense" in inputs:
        for feat in dense_cfg:
            name = feat["name"]
            if mapping and "dense" in mapping:
                col = mapping["dense"].get(name)
                if col is None:
                    raise ValueError(f"Dense feature '{name}' not found in column mapping")
            else:
                col = [f["name"] for f in dense_cfg].index(name)
            x = inputs["dense"][:, col:col + 1].float()
            feats.append(linear(x, state[f"{tower}.embeddings.{name}.0.weight"],
                                state[f"{tower}.embeddings.{name}.0.bias"]))
    if seq_cfg and "sequence" in inputs and inputs["sequence"]:
        tp = tcfg.get("transformer_parameters", {})
        feats.append(sequence_encoder(inputs["sequence"], state, f"{tower}.seq_encoder", seq_cfg,
                                      tp.get("n_head", 4), tp.get("n_layers", 1)))
    if not feats:
        raise RuntimeError("Tower received no valid features. Check if input_dict matches config")
    x = torch.cat(feats, dim=1)
    x = batchnorm1d(x, state, f"{tower}.feature_bn", training, update_running)
    return mlp_tower(x, state, f"{tower}.mlp", len(tcfg["mlp_hidden_dim"]), training, update_running)


def two_tower_forward(batch: dict, state: State, cfg: dict, user_mapping=None, item_mapping=None,
                      training: bool = True) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """TwoTowerModel.py:35-62: user pass, item pass, then one item-tower pass
    PER hard-negative slab (each with its own BatchNorm batch statistics and
    its own running-stat update), stacked on dim 1."""
    u = tower_forward(batch["user_tower"], user_mapping, state, "user_tower", cfg, training)
    i = tower_forward(batch["item_tower"], item_mapping, state, "item_tower", cfg, training)
    hn = None
    if batch.get("hard_negatives"):
        slabs = [tower_forward(s, item_mapping, state, "item_tower", cfg, training)
                 for s in batch["hard_negatives"]]
        hn = torch.stack(slabs, dim=1)
    return u, i, hn


# --------------------------------------------------------------------------
# a9: loss (TwoTowerModel.py:81-150)
# --------------------------------------------------------------------------
def build_logits(u: Tensor, i: Tensor, item_ids: Optional[Tensor], hn: Optional[Tensor],
                 temperature: float, hn_pool: Optional[Tensor] = None) -> Tensor:
    B = u.shape[0]
    z = (u @ i.t()) / temperature
    if item_ids is not None:
        ids = item_ids.reshape(-1)
        coll = (ids[:, None] == ids[None, :]) & ~torch.eye(B, dtype=torch.bool)
        z = z.masked_fill(coll, MASK_VALUE)
    cols = [z]
    if hn is not None:
        cols.append(torch.einsum("bd,bnd->bn", u, hn) / temperature)
    if hn_pool is not None:
        # shared-pool mode == per-row mode called with pool.expand(B,H,D)
        # (SURVEY.md section 7 hard part 1)
        cols.append((u @ hn_pool.t()) / temperature)
    return torch.cat(cols, dim=1) if len(cols) > 1 else z


def compute_loss(u: Tensor, i: Tensor, item_ids: Optional[Tensor] = None, hn: Optional[Tensor] = None,
                 temperature: float = 0.1, hn_pool: Optional[Tensor] = None) -> Tensor:
    if torch.isnan(u).any():
        raise RuntimeError("Found NaN in User Embedding")
    if torch.isnan(i).any():
        raise RuntimeError("Found NaN in Item Embedding")
    if hn is not None and torch.isnan(hn).any():
        raise RuntimeError("Found NaN in Hard Negative Embedding")
    z = build_logits(u, i, item_ids, hn, temperature, hn_pool)
    B = u.shape[0]
    lse = torch.logsumexp(z, dim=1)
    return (lse - z[torch.arange(B), torch.arange(B)]).mean()


def loss_grads_closed_form(u, i, item_ids, hn, temperature, hn_pool=None, grad_loss: float = 1.0):
    """KAT 4 of SURVEY.md section 4: G = (softmax(Z) - onehot)/B,
    dU = (G[:, :B] I + sum_n G[:, B+n] HN[:, n]) / T, dI = G[:, :B]^T U / T,
    dHN[b, n] = G[b, B+n] U[b] / T.  float64 accumulation."""
    B = u.shape[0]
    z = build_logits(u.double(), i.double(), item_ids, None if hn is None else hn.double(), temperature,
                     None if hn_pool is None else hn_pool.double())
    p = torch.softmax(z, dim=1)
    g = p.clone()
    g[torch.arange(B), torch.arange(B)] -= 1.0
    g = g * (grad_loss / B)
    du = g[:, :B] @ i.double()
    di = g[:, :B].t() @ u.double()
    dhn = dpool = None
    off = B
    if hn is not None:
        N = hn.shape[1]
        gh = g[:, off:off + N]
        du = du + torch.einsum("bn,bnd->bd", gh, hn.double())
        dhn = (gh[:, :, None] * u.double()[:, None, :]) / temperature
        off += N
    if hn_pool is not None:
        gp = g[:, off:]
        du = du + gp @ hn_pool.double()
        dpool = (gp.t() @ u.double()) / temperature
    return (du / temperature).float(), (di / temperature).float(), \
        None if dhn is None else dhn.float(), None if dpool is None else dpool.float()


def row_logsumexp(u, i, item_ids, hn, temperature, hn_pool=None) -> Tuple[Tensor, Tensor]:
    """(lse[B], positive logit[B]) in float64 -- what the fused kernel emits."""
    z = build_logits(u.double(), i.double(), item_ids, None if hn is None else hn.double(), temperature,
                     None if hn_pool is None else hn_pool.double())
    B = u.shape[0]
    return torch.logsumexp(z, dim=1), z[torch.arange(B), torch.arange(B)]


# --------------------------------------------------------------------------
# a11/a12: clip + Adam (training_utils.py:53-56, train_twotower.py:111; A5)
# --------------------------------------------------------------------------
def clip_coef(grads: List[Tensor], max_norm: float = 1.0) -> Tuple[float, float]:
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
    return min(1.0, max_norm / (total + 1e-6)), total


def adam_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
                beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """Dense torch.optim.Adam (no weight decay, no amsgrad); returns new (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


def sparse_rows_adam(table: Tensor, m: Tensor, v: Tensor, rows: Tensor, row_grad: Tensor, coef: float,
                     step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """Touched-rows-only ("lazy") Adam: equals dense Adam on the first step a
    row is touched (SURVEY.md section 7 hard part 3)."""
    table, m, v = table.clone(), m.clone(), v.clone()
    p2, m2, v2 = adam_update(table[rows], row_grad * coef, m[rows], v[rows], step, lr, beta1, beta2, eps)
    table[rows], m[rows], v[rows] = p2, m2, v2
    return table, m, v


# --------------------------------------------------------------------------
# integer side: sorted unique rows / segment sums (numpy), top-K tie-break
# --------------------------------------------------------------------------
def segment_rows(ids: np.ndarray, grad_rows: np.ndarray, padding_idx: Optional[int]):
    """(unique_rows ascending, row_grad[U, D]) with float64 accumulation --
    the compact form of embedding_grad_dense."""
    flat = ids.reshape(-1).astype(np.int64)
    g = grad_rows.reshape(flat.shape[0], -1).astype(np.float64)
    if padding_idx is not None:
        keep = flat != padding_idx
        flat, g = flat[keep], g[keep]
    rows, inv = np.unique(flat, return_inverse=True)
    out = np.zeros((rows.shape[0], g.shape[1]), dtype=np.float64)
    np.add.at(out, inv, g)
    return rows, out.astype(np.float32)


def score_topk(q: np.ndarray, e: np.ndarray, k: int, row_offset: int = 0,
               hist_mask: Optional[List[np.ndarray]] = None):
    """Retrieval scoring + top-K (training_utils.py:220,255-258) under the
    STATED tie-break: order by (score descending, corpus row ascending).
    Scores are the float64 dot products of the stored (fp32/bf16-valued)
    embeddings; ``hist_mask[i]`` lists corpus rows set to -inf for query i
    (training_utils.py:238-252).  Returns (scores float64 [Bq,k], rows int64 [Bq,k])."""
    # fixed-order accumulation over d (NOT a BLAS call): duplicate corpus rows must give bitwise-equal
    # scores so that the row-index tie-break is well defined
    q64, e64 = q.astype(np.float64), e.astype(np.float64)
    s = np.zeros((q64.shape[0], e64.shape[0]), dtype=np.float64)
    for d in range(q64.shape[1]):
        s += q64[:, d:d + 1] * e64[None, :, d]
    if hist_mask is not None:
        for r, cols in enumerate(hist_mask):
            if len(cols):
                s[r, np.asarray(cols, dtype=np.int64)] = -np.inf
    n = s.shape[1]
    idx = np.broadcast_to(np.arange(n, dtype=np.int64), s.shape)
    order = np.lexsort((idx, -s), axis=1)[:, :k]  # primary -s, secondary idx
    return np.take_along_axis(s, order, axis=1), order + row_offset


def recall_at_k(topk_rows: np.ndarray, corpus_ids: np.ndarray, targets: np.ndarray, k: int) -> int:
    """training_utils.py:255-263: hit iff target id is among the first k mapped ids."""
    pred = corpus_ids[topk_rows[:, :k]]
    return int((pred == targets.reshape(-1, 1)).any(axis=1).sum())


# --------------------------------------------------------------------------
# whole training step on the CPU (the "port" CPU baseline of bench.py)
# --------------------------------------------------------------------------
def trainable_keys(state: State) -> List[str]:
    return [k for k, t in state.items() if t.is_floating_point()
            and not k.endswith(("running_mean", "running_var"))]


def padding_rows(cfg: dict) -> Dict[str, int]:
    """state key -> padding_idx whose gradient row is masked (GenericTower.py:43,
    SequenceFeatureProcessor.py:22)."""
    out = {}
    for tower in ("user_tower", "item_tower"):
        tcfg = cfg["two_tower"].get(tower, {})
        for feat in tcfg.get("sparse_features") or []:
            out[f"{tower}.embeddings.{feat['name']}.weight"] = feat.get("padding_idx", 0)
        for feat in tcfg.get("sequence_features") or []:
            out[f"{tower}.seq_encoder.feature_embedder.embeddings.{feat['name']}.weight"] = \
                feat.get("padding_index", 0)
    return out


def train_step(batch: dict, state: State, opt: dict, cfg: dict, user_mapping=None, item_mapping=None,
               temperature: float = 0.1, lr: float = 5e-4, max_grad_norm: float = 1.0,
               item_id_col: int = 0):
    """zero_grad -> forward -> loss -> backward -> clip_grad_norm_ -> dense Adam
    (training_utils.py:28-60).  ``opt`` = {'step': int, 'm': {...}, 'v': {...}}.
    Mutates ``state``/``opt`` in place, returns (loss, grads)."""
    keys = trainable_keys(state)
    leaves = {k: state[k].detach().clone().requires_grad_(True) for k in keys}
    work = dict(state)
    work.update(leaves)
    u, i, hn = two_tower_forward(batch, work, cfg, user_mapping, item_mapping, training=True)
    item_ids = batch["item_tower"]["sparse"][:, item_id_col]  # training_utils.py:36-40,72-91
    loss = compute_loss(u, i, item_ids, hn, temperature)
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    grads = {k: (torch.zeros_like(state[k]) if g is None else g) for k, g in zip(keys, grads)}
    for k, pad in padding_rows(cfg).items():
        if k in grads and pad is not None:
            grads[k][pad] = 0.0  # nn.Embedding(padding_idx=...) never accumulates grad for the pad row
    if max_grad_norm > 0:
        coef, _ = clip_coef(list(grads.values()), max_grad_norm)
    else:
        coef = 1.0
    opt["step"] += 1
    for k in keys:
        if k not in opt["m"]:
            opt["m"][k] = torch.zeros_like(state[k])
            opt["v"][k] = torch.zeros_like(state[k])
        g = grads[k] * coef
        state[k], opt["m"][k], opt["v"][k] = adam_update(state[k], g, opt["m"][k], opt["v"][k], opt["step"], lr)
    for k in work:  # BN running stats were updated on `work`
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            state[k] = work[k]
    return loss.detach(), grads
