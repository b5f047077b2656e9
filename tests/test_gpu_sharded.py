"""CUDA kernels of the row-sharded exchange (include/tt_b200.h section 7, tt_emb_segment_grad_lists) against the CPU
restatement of the wire format (tests/sharded_cpu_ops.py): integer buffers bit-exact, float buffers bit-exact too
(both sides add fp32 rows in position order), then the group end to end at world 1 against the unsharded kernels and
the oracle."""
import pytest
import torch

from sharded_cpu_ops import CpuShardOps

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _layout(B, L, W, D, factor=1.5):
    n_pos = B * L
    cap = n_pos if W == 1 else min(n_pos, int(n_pos / W * factor) + 64)
    cap = (cap + 3) // 4 * 4
    off_base = 8                                   # not at the start of the block on purpose
    rows_base = off_base + (B + 1 + 3) // 4 * 4
    block_ints = rows_base + cap + 12
    vec_base = 16
    vec_rows = B if L > 1 else cap
    block_floats = vec_base + vec_rows * D + 8
    return dict(cap=cap, off_base=off_base, rows_base=rows_base, block_ints=block_ints, vec_base=vec_base,
                vec_rows=vec_rows, block_floats=block_floats)


def _ids(gen, B, L, V, pad):
    ids = torch.randint(1, V, (B, L), generator=gen)
    if L > 1:
        lens = torch.randint(0, L + 1, (B, 1), generator=gen)
        ids[torch.arange(L)[None, :] >= lens] = pad if pad is not None else 1
        ids[1, : min(L, 5)] = 3                     # repeated id in one sample
    elif pad is not None:
        ids[::7] = pad
    return ids


@pytest.mark.parametrize("B,L,W,V,pad", [(37, 50, 4, 1003, 0), (64, 1, 8, 5001, 0), (33, 1, 3, 777, None),
                                         (130, 200, 8, 100003, 0), (9, 33, 2, 50, 0), (200, 7, 1, 999, 0)])
def test_route_matches_restatement_bit_exact(B, L, W, V, pad):
    from recommendsystemproject_b200 import sharded
    gen = torch.Generator().manual_seed(B * 1000 + L)
    ids = _ids(gen, B, L, V, pad)
    lay = _layout(B, L, W, 8)
    ref_send = torch.zeros(W, lay["block_ints"], dtype=torch.int32)
    ref_npad = torch.zeros(B, dtype=torch.int32)
    ref_flags = torch.zeros(1, dtype=torch.int32)
    CpuShardOps.route(ids, pad, V, W, ref_send, lay["block_ints"], lay["off_base"], lay["rows_base"], lay["cap"], ref_npad, ref_flags)
    send = torch.zeros(W, lay["block_ints"], dtype=torch.int32, device=DEV)
    npad = torch.zeros(B, dtype=torch.int32, device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    sharded._CudaShardOps.route(ids.to(DEV), pad, V, W, send, lay["block_ints"], lay["off_base"], lay["rows_base"], lay["cap"], npad, flags)
    a, b = lay["off_base"], lay["rows_base"] + lay["cap"]
    assert torch.equal(send.cpu()[:, a:b], ref_send[:, a:b])
    assert torch.equal(npad.cpu(), ref_npad)
    assert int(flags.item()) == int(ref_flags.item()) == 0


def test_route_flags_out_of_range_and_overflow():
    from recommendsystemproject_b200 import sharded
    B, L, W, V = 40, 10, 2, 202
    lay = _layout(B, L, W, 8, factor=1.0)
    send = torch.zeros(W, lay["block_ints"], dtype=torch.int32, device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    ids = torch.full((B, L), 2, dtype=torch.int64, device=DEV)
    ids[3, 4] = 9999
    ids[5, 0] = -3
    sharded._CudaShardOps.route(ids, None, V, W, send, lay["block_ints"], lay["off_base"], lay["rows_base"], lay["cap"], None, flags)
    assert int(flags.item()) == 3


@pytest.mark.parametrize("B,L,W,V,D,dtype", [(37, 50, 4, 1003, 128, torch.float32), (64, 1, 8, 5001, 64, torch.float32),
                                             (50, 20, 2, 333, 8, torch.float32), (41, 30, 3, 2000, 128, torch.bfloat16),
                                             (25, 1, 2, 400, 32, torch.bfloat16)])
def test_owner_gather_combine_grad_pack_segment_grad_match_restatement(B, L, W, V, D, dtype):
    """Simulates one owner (rank 0) receiving blocks from W sources, and one source receiving vectors from W owners."""
    from recommendsystemproject_b200 import ops, sharded
    C = sharded._CudaShardOps
    gen = torch.Generator().manual_seed(77 + B)
    pad = 0
    lay = _layout(B, L, W, D)
    mode = ops.POOL_MEAN if L > 1 else ops.POOL_NONE
    local_rows = (V + W - 1) // W
    table = torch.randn(local_rows, D, generator=gen).to(dtype)
    # blocks sent by W different sources: what owner 0 receives is block [0] of each
    src_ids = [_ids(gen, B, L, V, pad) for _ in range(W)]
    sends = []
    for s in range(W):
        snd = torch.zeros(W, lay["block_ints"], dtype=torch.int32)
        CpuShardOps.route(src_ids[s], pad, V, W, snd, lay["block_ints"], lay["off_base"], lay["rows_base"], lay["cap"],
                          torch.zeros(B, dtype=torch.int32), torch.zeros(1, dtype=torch.int32))
        sends.append(snd)
    recv = torch.stack([sends[s][0] for s in range(W)])
    # --- owner gather
    ref_out = torch.zeros(W, lay["block_floats"])
    ref_pos = torch.full((W * lay["cap"],), -7, dtype=torch.int32)
    CpuShardOps.owner_gather(table, local_rows, W, recv, lay["block_ints"], lay["off_base"], lay["rows_base"], lay["cap"], B,
                             L > 1, ref_out, lay["block_floats"], lay["vec_base"], ref_pos)
    out = torch.zeros(W, lay["block_floats"], device=DEV)
    pos = torch.full((W * lay["cap"],), -7, dtype=torch.int32, device=DEV)
    recv_d = recv.to(DEV)
    C.owner_gather(table.to(DEV), local_rows, W, recv_d, lay["block_ints"], lay["off_base"], lay["rows_base"], lay["cap"], B,
                   L > 1, out, lay["block_floats"], lay["vec_base"], pos)
    assert torch.equal(out.cpu(), ref_out)
    assert torch.equal(pos.cpu(), ref_pos)
    # --- combine on source 0 (vectors from W owners: reuse random data in the vec layout)
    vec_in = torch.randn(W, lay["block_floats"], generator=gen)
    n_pad = (src_ids[0] == pad).sum(1).to(torch.int32)
    pad_row = torch.randn(D, generator=gen)
    ref_c = torch.zeros(B, D)
    CpuShardOps.combine(vec_in, lay["block_floats"], lay["vec_base"], W, src_ids[0], pad, V, mode, sends[0], lay["block_ints"],
                        lay["off_base"], lay["cap"], n_pad, pad_row, D, ref_c)
    got_c = torch.zeros(B, D, device=DEV)
    C.combine(vec_in.to(DEV), lay["block_floats"], lay["vec_base"], W, src_ids[0].to(DEV), pad, V, mode, sends[0].to(DEV),
              lay["block_ints"], lay["off_base"], lay["cap"], n_pad.to(DEV), pad_row.to(DEV), D, got_c)
    # pads enter as n_pad * pad_row: an FMA on the GPU, multiply-then-add in the restatement => equal to one rounding
    assert torch.allclose(got_c.cpu(), ref_c, atol=1e-6, rtol=1e-6)
    # --- grad pack on source 0
    g = torch.randn(B, D, generator=gen)
    ref_g = torch.zeros(W, lay["block_floats"])
    CpuShardOps.grad_pack(g, mode, D, W, src_ids[0], pad, V, sends[0], lay["block_ints"], lay["off_base"], lay["cap"], ref_g,
                          lay["block_floats"], lay["vec_base"])
    got_g = torch.zeros(W, lay["block_floats"], device=DEV)
    C.grad_pack(g.to(DEV), mode, D, W, src_ids[0].to(DEV), pad, V, sends[0].to(DEV), lay["block_ints"], lay["off_base"],
                lay["cap"], got_g, lay["block_floats"], lay["vec_base"])
    assert torch.equal(got_g.cpu(), ref_g)
    # --- owner-side segment gradient over the received lists, gradients from W sources
    g_in = torch.randn(W, lay["block_floats"], generator=gen)
    n_pos = W * lay["cap"]
    r_rows = torch.zeros(n_pos, dtype=torch.int64)
    r_grad = torch.zeros(n_pos, D)
    r_nu = torch.zeros(1, dtype=torch.int32)
    r_sq = torch.zeros(1)
    CpuShardOps.segment_grad_lists(recv, W, lay["block_ints"], lay["rows_base"], lay["cap"], ref_pos, local_rows, g_in,
                                   lay["vec_rows"], lay["block_floats"], lay["vec_base"], D, r_rows, r_grad, r_nu, r_sq, None)
    d_rows = torch.zeros(n_pos, dtype=torch.int64, device=DEV)
    d_grad = torch.zeros(n_pos, D, device=DEV)
    d_nu = torch.zeros(1, dtype=torch.int32, device=DEV)
    d_sq = torch.zeros(1, device=DEV)
    ws = torch.empty(C.segment_ws_bytes(n_pos, D), dtype=torch.uint8, device=DEV)
    C.segment_grad_lists(recv_d, W, lay["block_ints"], lay["rows_base"], lay["cap"], pos, local_rows, g_in.to(DEV),
                         lay["vec_rows"], lay["block_floats"], lay["vec_base"], D, d_rows, d_grad, d_nu, d_sq, ws)
    U = int(r_nu.item())
    assert int(d_nu.item()) == U and U > 0
    assert torch.equal(d_rows.cpu()[:U], r_rows[:U])
    # ascending positions inside a segment on both sides; segments longer than a chunk are summed chunk-wise by the
    # kernel (fixed order, but not the restatement's left-to-right order): equal to rounding
    assert torch.allclose(d_grad.cpu()[:U], r_grad[:U], atol=1e-5, rtol=1e-5)
    assert abs(float(d_sq) - float(r_sq)) <= 1e-5 * float(r_sq)


def test_group_world1_equals_unsharded_kernels_and_oracle():
    """At W = 1 the sharded path must reproduce the direct gather+pool kernel (forward equal to rounding: the pad rows
    are added as n_pad * row instead of one by one... same as the direct kernel, so bit-exact) and the direct
    segment-gradient kernel (bit-exact rows)."""
    from recommendsystemproject_b200 import ops, sharded
    from oracle import twotower_oracle as O
    gen = torch.Generator().manual_seed(5)
    V, D, B, L = 5003, 128, 300, 40
    w = torch.randn(V, D, generator=gen)
    ids = _ids(gen, B, L, V, 0)
    up = torch.randn(B, D, generator=gen)
    grp = sharded.ShardedTableGroup(0, 1, DEV)
    grp.fused_adam = False          # keep the per-row gradients (the deferred form never writes them): inspected below
    wd = w.to(DEV).clone()
    grp.add_table("hist", V, D, ops.POOL_MEAN, 0, wd, wd[0].clone())
    grp.zero_grad()
    got = grp.lookup({"hist": ids.to(DEV)})["hist"]
    direct = ops.gather_rows(w.to(DEV), ids.to(DEV), "mean", 0)
    assert torch.allclose(got, direct, atol=1e-6, rtol=1e-6)
    ref = O.pooled_lookup(w, ids, "mean")
    assert torch.allclose(got.cpu(), ref, atol=1e-6, rtol=1e-5)
    (got * up.to(DEV)).sum().backward()
    grp.check_flags()
    rows, row_grad, nu = grp.tables["hist"].pending
    r2, g2, nu2 = ops.segment_grad(ids.to(DEV), ops.POOL_MEAN, 0, V, up.to(DEV), None, D)
    U = int(nu.item())
    assert U == int(nu2.item())
    assert torch.equal(rows[:U], r2[:U])
    assert torch.allclose(row_grad[:U], g2[:U], atol=1e-6, rtol=1e-5)


def test_group_lookup_under_cuda_graph_replays_with_new_ids():
    from recommendsystemproject_b200 import ops, sharded
    gen = torch.Generator().manual_seed(11)
    V, D, B, L = 2001, 64, 128, 20
    w = torch.randn(V, D, generator=gen).to(DEV)
    grp = sharded.ShardedTableGroup(0, 1, DEV)
    grp.add_table("t", V, D, ops.POOL_SUM, 0, w.clone(), w[0].clone())
    static = _ids(gen, B, L, V, 0).to(DEV)
    grp.lookup({"t": static})
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = grp.lookup({"t": static})["t"]
    new = _ids(gen, B, L, V, 0).to(DEV)
    static.copy_(new)
    g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(out, ops.gather_rows(w, new, "sum", 0), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("D", [64, 128, 256])
def test_deferred_segment_adam_is_bitwise_the_two_kernel_form(D):
    """tt_emb_segment_grad_lists(row_grad = NULL) + tt_emb_segment_adam_lists (the per-row gradient sum is formed again
    inside the Adam kernel) against segment reduction -> row_grad -> tt_emb_rowwise_adam: same additions in the same
    order, so weights, both moments and the gradient norm must be BITWISE equal -- pooled table with heavy-hitter rows
    (segments far longer than a chunk) and a single-id table, two steps."""
    from recommendsystemproject_b200 import ops, sharded
    gen = torch.Generator().manual_seed(17 + D)
    V, B, L = 3001, 600, 50
    w = torch.randn(V, D, generator=gen)
    w1 = torch.randn(V, D, generator=gen)
    ids = _ids(gen, B, L, V, 0)
    ids[:, :4] = 7                                   # one row with 2400 positions, more with dozens
    ids[::3, 5] = 11
    ids1 = torch.randint(1, V, (B,), generator=gen)
    ids1[::5] = 42
    up = torch.randn(2, B, D, generator=gen).to(DEV)
    results = []
    for fused in (False, True):
        grp = sharded.ShardedTableGroup(0, 1, DEV)
        grp.fused_adam = fused
        wd, w1d = w.to(DEV).clone(), w1.to(DEV).clone()
        grp.add_table("hist", V, D, ops.POOL_MEAN, 0, wd, wd[0].clone())
        grp.add_table("uid", V, D, ops.POOL_NONE, 0, w1d, w1d[0].clone())
        grp.init_state()
        step_dev = torch.zeros(1, dtype=torch.int64, device=DEV)
        coef = torch.tensor([0.37], device=DEV)
        norms = []
        for k in range(2):
            grp.zero_grad()
            out = grp.lookup({"hist": ids.to(DEV), "uid": ids1.to(DEV)})
            ((out["hist"] * up[0]).sum() + (out["uid"] * up[1]).sum() * (k + 1)).backward()
            grp.check_flags()
            assert (grp.tables["hist"].pending[1] is None) == fused
            norms.append(grp.sq_terms.clone())
            step_dev += 1
            grp.step(coef, 1e-2, step_dev)
        results.append((wd, w1d, grp.tables["hist"].exp_avg, grp.tables["hist"].exp_avg_sq, grp.tables["uid"].exp_avg,
                        grp.tables["uid"].exp_avg_sq, *norms))
    for a, b in zip(*results):
        assert torch.equal(a, b)
    assert not torch.equal(results[0][0], w.to(DEV))        # the step did move the table
