"""Host logic of the row-sharded table group (sharded.ShardedTableGroup) on CPU: world 1 in-process and world 2 over
gloo, with the CPU restatement of the device ops (tests/sharded_cpu_ops.py) injected.  Checked against the oracle's
unsharded pooling / embedding backward (oracle/twotower_oracle.py follows GenericTower.py:141-183)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sharded_cpu_ops import CpuShardOps


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


def _make_tables(gen):
    V1, V2, D = 211, 97, 8
    return {"hist": (V1, D, "mean", 0, torch.randn(V1, D, generator=gen)),
            "uid": (V2, D, None, None, torch.randn(V2, D, generator=gen)),
            "iid": (V1, D, None, 0, torch.randn(V1, D, generator=gen))}


def _make_ids(rank, B=13, L=9):
    g = torch.Generator().manual_seed(500 + rank)
    hist = torch.randint(1, 211, (B, L), generator=g)
    lens = torch.randint(0, L + 1, (B, 1), generator=g)      # includes an all-padding sample now and then
    hist[torch.arange(L)[None, :] >= lens] = 0
    hist[0, :3] = 7                                          # a repeated id inside one sample
    uid = torch.randint(0, 97, (B, 1), generator=g)
    iid = torch.randint(0, 211, (B, 1), generator=g)         # id 0 = the pad row of that table
    up = {k: torch.randn(B, 8, generator=g) for k in ("hist", "uid", "iid")}
    return {"hist": hist, "uid": uid, "iid": iid}, up


def _reference(tables, all_ids, all_up):
    """Unsharded pooled lookup + embedding backward over the batches of ALL ranks (GenericTower.py:153-160,182;
    mean divides by L incl. pads, pad positions get no gradient)."""
    out, grads = [], {}
    for name, (V, D, mode, pad, w) in tables.items():
        grads[name] = torch.zeros(V, D)
    for ids, up in zip(all_ids, all_up):
        o = {}
        for name, (V, D, mode, pad, w) in tables.items():
            wt = w.clone().requires_grad_(True)
            x = ids[name]
            e = wt[x]                                         # [B, L, D]
            pooled = e.mean(1) if mode == "mean" else e.sum(1)
            (pooled * up[name]).sum().backward()
            g = wt.grad
            if pad is not None:
                g[pad] = 0
            grads[name] += g
            o[name] = pooled.detach()
        out.append(o)
    return out, grads


def _case(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from recommendsystemproject_b200 import ops, sharded
    tables = _make_tables(torch.Generator().manual_seed(9))
    grp = sharded.ShardedTableGroup(rank, world, "cpu", capacity_factor=2.0, dev_ops=CpuShardOps)
    for name, (V, D, mode, pad, w) in tables.items():
        grp.add_table(name, V, D, ops.POOL_MODES[mode], pad, w[rank::world].clone(),
                      None if pad is None else w[pad].clone())
    all_ids, all_up = zip(*[_make_ids(r) for r in range(world)])
    ref_out, ref_grad = _reference(tables, all_ids, all_up)
    ids, up = all_ids[rank], all_up[rank]
    grp.zero_grad()
    got = grp.lookup(ids)
    ok = True
    for name in tables:
        ok &= bool(torch.allclose(got[name], ref_out[rank][name], atol=1e-6, rtol=1e-5))
    sum((got[n] * up[n]).sum() for n in tables).backward()
    grp.check_flags()
    sq = grp.local_sq_norm().clone()
    if world > 1:
        dist.all_reduce(sq)
    sq_ref = sum(float(g.double().pow(2).sum()) for g in ref_grad.values())
    ok &= abs(float(sq) - sq_ref) < 1e-4 * max(1.0, sq_ref)
    for name, (V, D, mode, pad, w) in tables.items():
        rows, row_grad, nu = grp.tables[name].pending
        U = int(nu[0])
        dense = torch.zeros(grp.tables[name].local_rows, D)
        dense[rows[:U]] = row_grad[:U]
        ok &= bool(torch.allclose(dense, ref_grad[name][rank::world], atol=1e-5, rtol=1e-5))
        ok &= bool(torch.equal(rows[:U], torch.unique(rows[:U])))          # ascending, unique
    # row-wise Adam moves exactly the touched rows; full-table round trip (checkpoint path)
    before = {n: grp.tables[n].weight.clone() for n in tables}
    grp.step(torch.ones(1), 1e-2, torch.ones(1, dtype=torch.int64))
    for name in tables:
        touched = ref_grad[name][rank::world].abs().sum(1) > 0
        moved = (grp.tables[name].weight != before[name]).any(1)
        ok &= bool(torch.equal(moved, touched))
    full = grp.gather_full_weight("hist")
    ok &= bool(torch.equal(full[rank::world], grp.tables["hist"].weight)) and full.shape[0] == 211
    grp.load_full_weight("hist", tables["hist"][4])
    ok &= bool(torch.equal(grp.tables["hist"].weight, tables["hist"][4][rank::world]))
    return ok


def test_group_world1_matches_unsharded_embedding():
    assert _case(0, 1)


def test_group_two_ranks_gloo_matches_unsharded_embedding():
    assert all(_run(_case).values())


def test_capacity_overflow_is_flagged():
    from recommendsystemproject_b200 import ops, sharded
    from recommendsystemproject_b200._lib import TTError

    class Two(sharded.ShardedTableGroup):     # world 2 layout without a process group: only the routing is exercised
        def _a2a(self, out, inp):
            out.copy_(inp)
    grp = Two(0, 2, "cpu", capacity_factor=1.0, dev_ops=CpuShardOps)
    w = torch.randn(101, 4)
    grp.add_table("t", 202, 4, ops.POOL_SUM, None, w, None)
    ids = torch.full((40, 10), 2, dtype=torch.int64)         # every id belongs to owner 0: twice its capacity
    grp.lookup({"t": ids})
    with pytest.raises(TTError, match="overflow"):
        grp.check_flags()
    grp.lookup({"t": torch.full((40, 10), 500, dtype=torch.int64)})
    with pytest.raises(IndexError):
        grp.check_flags()
