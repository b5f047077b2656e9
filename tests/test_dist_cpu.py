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
