"""world_size-2 gloo tests (CPU) of the multi-GPU HOST logic: id routing for row-sharded tables
(owner = row % W, all-to-all out and back) and corpus-sharded top-K with global merge.  The CUDA kernels the
routing feeds are replaced here by the CPU oracle (tests may use oracle/); the production path passes the
CUDA ops (tests/test_gpu_dist.py on a multi-GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _lookup_case(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from recommendsystemproject_b200 import dist as tdist
    from oracle import twotower_oracle as O
    gen = torch.Generator().manual_seed(123)
    V, D = 101, 8
    table = torch.randn(V, D, generator=gen)            # same on every rank (seeded)
    local = table[rank::world].contiguous()             # owner = row % W, local row = row // W
    g2 = torch.Generator().manual_seed(1000 + rank)     # each rank looks up its OWN ids
    ids = torch.randint(0, V, (7, 5), generator=g2)
    got = tdist.sharded_lookup(lambda rows: O.embedding_lookup(local, rows), ids, world)
    return bool(torch.equal(got, table[ids]))


def _topk_case(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from recommendsystemproject_b200 import dist as tdist
    from oracle import twotower_oracle as O
    gen = torch.Generator().manual_seed(7)
    Nc, D, K, Bq = 240, 16, 10, 9
    corpus = torch.nn.functional.normalize(torch.randn(Nc, D, generator=gen), dim=1)
    corpus[130] = corpus[5]                              # a cross-shard exact tie
    query = torch.nn.functional.normalize(torch.randn(Bq, D, generator=gen), dim=1)
    bounds = [0, 100, Nc]                                # uneven shards
    local = corpus[bounds[rank]:bounds[rank + 1]]

    def topk_fn(q, e, k, off):
        s, i = O.score_topk(q.numpy(), e.numpy(), k, row_offset=off)
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge_fn(ss, ii):
        W, B, k = ss.shape
        s = ss.permute(1, 0, 2).reshape(B, W * k).numpy()
        i = ii.permute(1, 0, 2).reshape(B, W * k).numpy()
        order = np.lexsort((i, -s), axis=1)[:, :k]
        return torch.from_numpy(np.take_along_axis(s, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

    s, i = tdist.sharded_topk(query, local, K, rank, world, bounds[:-1], topk_fn, merge_fn)
    s_ref, i_ref = O.score_topk(query.numpy(), corpus.numpy(), K)
    return bool(np.array_equal(i.numpy(), i_ref) and np.allclose(s.numpy(), s_ref, atol=1e-12))


def test_route_ids_is_a_permutation_grouped_by_owner():
    from recommendsystemproject_b200 import dist as tdist
    ids = torch.tensor([[5, 2, 9], [4, 4, 7]])
    send, counts, order = tdist.route_ids(ids, 2)
    assert counts.tolist() == [3, 3]
    flat = ids.reshape(-1)
    assert (flat[order] % 2).tolist() == [0, 0, 0, 1, 1, 1]
    assert torch.equal(send, flat[order] // 2)
    assert sorted(order.tolist()) == list(range(6))


def test_sharded_lookup_two_ranks_gloo():
    assert all(_run(_lookup_case).values())


def test_sharded_topk_two_ranks_gloo():
    assert all(_run(_topk_case).values())


# ---------------------------------------------------------------- ShardedEmbeddingBag (host logic, gloo, oracle ops)
class _CpuOps:
    """Oracle stand-ins for the CUDA kernels ShardedEmbeddingBag drives (tests only)."""

    @staticmethod
    def gather(table, rows):
        return table[rows.reshape(-1)].float()

    @staticmethod
    def pool(table, idx, mode):
        g = table[idx]                                     # [B, L, D]
        if mode == 0:
            return g[:, 0]
        return g.sum(1) if mode == 1 else g.mean(1)

    @staticmethod
    def segment_grad(rows, vocab, grad, sq_norm):
        u, inv = torch.unique(rows.reshape(-1), sorted=True, return_inverse=True)
        rg = torch.zeros(len(u), grad.shape[1], dtype=torch.float64).index_add_(0, inv, grad.double()).float()
        sq_norm += rg.double().pow(2).sum().float()
        return u, rg, torch.tensor([len(u)], dtype=torch.int32)

    @staticmethod
    def pool_sum(table, idx, null_row):
        return table[idx].sum(1)

    @staticmethod
    def segment_grad_pooled(idx, null_row, grad, sq_norm):
        flat = idx.reshape(-1)
        g = grad.repeat_interleave(idx.shape[1], 0)
        keep = flat != null_row
        return _CpuOps.segment_grad(flat[keep], null_row, g[keep], sq_norm)

    @staticmethod
    def adam(table, m, v, rows, row_grad, n_unique, coef, lr, b1, b2, eps, step_dev):
        from oracle import twotower_oracle as O
        n = int(n_unique)
        r = rows[:n]
        g = row_grad[:n] * float(coef)
        p, mm, vv = O.adam_update(table[r], g, m[r], v[r], int(step_dev), lr, b1, b2, eps)
        table[r], m[r], v[r] = p, mm, vv


def _sharded_bag_case(rank, world, exchange="rows"):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from recommendsystemproject_b200 import dist as tdist
    from oracle import twotower_oracle as O
    gen = torch.Generator().manual_seed(5)
    V, D, L, B = 97, 8, 6, 11
    full = torch.randn(V, D, generator=gen)
    bag = tdist.ShardedEmbeddingBag(V, D, rank, world, "mean", 0, device="cpu", dev_ops=_CpuOps, full_weight=full,
                                    exchange=exchange)
    g2 = torch.Generator().manual_seed(50 + rank)
    ids = torch.randint(1, V, (B, L), generator=g2)
    ids[torch.arange(L)[None, :] >= torch.randint(1, L + 1, (B, 1), generator=g2)] = 0
    ids[0] = 0                                              # an all-padding sample
    up = torch.randn(B, D, generator=g2)
    bag.zero_grad()
    pooled = bag(ids)
    ok = torch.allclose(pooled, full[ids].mean(1), atol=1e-6)
    (pooled * up).sum().backward()
    coef = tdist.global_clip_coef([bag.sq_norm], 1.0)
    step = torch.ones(1, dtype=torch.int64)
    bag.step(coef, 1e-2, step)
    # reference: dense gradient of ALL ranks' batches on the full table, clip by the global norm, dense Adam on the
    # touched rows (first step: lazy == dense), pad row frozen
    ids_all = [torch.empty_like(ids) for _ in range(world)]
    up_all = [torch.empty_like(up) for _ in range(world)]
    dist.all_gather(ids_all, ids)
    dist.all_gather(up_all, up)
    dense = torch.zeros(V, D, dtype=torch.float64)
    for i_r, u_r in zip(ids_all, up_all):
        dense.index_add_(0, i_r.reshape(-1), (u_r.double() / L).repeat_interleave(L, 0))
    dense[0] = 0
    c_ref = min(1.0, 1.0 / (float(dense.pow(2).sum().sqrt()) + 1e-6))
    ok &= abs(float(coef) - c_ref) < 1e-5
    touched = torch.nonzero(dense.abs().sum(1) > 0).reshape(-1)
    p_ref, _, _ = O.adam_update(full[touched], (dense[touched] * c_ref).float(), torch.zeros(len(touched), D),
                                torch.zeros(len(touched), D), 1, 1e-2)
    mine = touched[touched % world == rank]
    ok &= torch.allclose(bag.weight[mine // world], p_ref[touched % world == rank], atol=1e-6)
    untouched = torch.ones(V, dtype=torch.bool)
    untouched[touched] = False
    mine_u = torch.nonzero(untouched).reshape(-1)
    mine_u = mine_u[mine_u % world == rank]
    ok &= torch.equal(bag.weight[mine_u // world], full[mine_u])
    return bool(ok)


def test_sharded_embedding_bag_two_ranks_gloo():
    assert all(_run(_sharded_bag_case).values())


def _sharded_bag_pooled_case(rank, world):
    return _sharded_bag_case(rank, world, exchange="pooled")


def test_sharded_embedding_bag_owner_side_pooling_two_ranks_gloo():
    """exchange='pooled': the owners pool and only [B, D] partial sums travel; same lookup, same update."""
    assert all(_run(_sharded_bag_pooled_case).values())


# ---------------------------------------------------------------- global-batch in-batch softmax (host logic, gloo, oracle CE)
def _global_ce_case(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from recommendsystemproject_b200 import dist as tdist
    from oracle import twotower_oracle as O
    gen = torch.Generator().manual_seed(77)
    B, D, H, T = 6, 8, 5, 0.2
    U = torch.nn.functional.normalize(torch.randn(world * B, D, generator=gen), dim=1)
    I = torch.nn.functional.normalize(torch.randn(world * B, D, generator=gen), dim=1)
    pool = torch.nn.functional.normalize(torch.randn(H, D, generator=gen), dim=1)
    ids = torch.arange(1, world * B + 1)                        # unique ids: no cross-rank collisions
    sl = slice(rank * B, (rank + 1) * B)
    u, i, pl = U[sl].clone().requires_grad_(True), I[sl].clone().requires_grad_(True), pool.clone().requires_grad_(True)
    ce = lambda uu, ii, idd, pp, tt: O.compute_loss(uu, ii, idd, None, tt, hn_pool=pp)
    loss = tdist.global_inbatch_ce(u, i, ids[sl], pl, T, ce_fn=ce)
    loss.backward()
    # single-process reference on the global batch
    Ug, Ig, Pg = U.clone().requires_grad_(True), I.clone().requires_grad_(True), pool.clone().requires_grad_(True)
    ref = O.compute_loss(Ug, Ig, ids, None, T, hn_pool=Pg)
    ref.backward()
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    ok = abs(float(tot) / world - float(ref)) < 1e-6
    # rank losses are each a mean over B rows and the global loss their mean: d(global)/dx = (1/W) d(sum of rank losses)/dx
    ok &= torch.allclose(u.grad / world, Ug.grad[sl], atol=1e-6)
    ok &= torch.allclose(i.grad / world, Ig.grad[sl], atol=1e-6)
    gp = pl.grad.clone()
    dist.all_reduce(gp)
    ok &= torch.allclose(gp / world, Pg.grad, atol=1e-6)
    return bool(ok)


def test_global_inbatch_ce_equals_single_process_global_batch_gloo():
    assert all(_run(_global_ce_case).values())
