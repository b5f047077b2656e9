"""Rectangular (global-batch) form of the tcgen05 in-batch CE (tt_ce_fwd_tc_rect / tt_ce_bwd_tc_rect): what rank r of a
data-parallel run computes -- its B/W user rows against ALL W*B/W gathered item rows with the false-negative mask over
the item ids of every rank -- must add up to the single-process loss on the global batch (TwoTowerModel.py:95-140):
    global loss = mean_r loss_r,   dU = concat_r dU_r / W,   dI = sum_r dI_r / W,   dPool = sum_r dPool_r / W.
Checked against the fp64 oracle on bf16-rounded inputs (forward 2e-4; backward 1e-2 relative Frobenius error: the
recomputed probabilities re-enter the tensor core as bf16) and against the square kernel itself."""
import pytest
import torch

from oracle import twotower_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rnd(t):
    return t.bfloat16().float()


def _data(Bg, D, H, n_ids, seed):
    gen = torch.Generator().manual_seed(seed)
    u = torch.nn.functional.normalize(torch.randn(Bg, D, generator=gen), dim=1)
    it = torch.nn.functional.normalize(torch.randn(Bg, D, generator=gen), dim=1)
    pool = torch.nn.functional.normalize(torch.randn(H, D, generator=gen), dim=1) if H else None
    ids = torch.randint(1, n_ids, (Bg,), generator=gen)          # many cross-"rank" collisions
    return u, it, pool, ids


def test_rect_with_square_shapes_is_the_square_kernel():
    from recommendsystemproject_b200 import ops
    u, it, pool, ids = _data(1000, 128, 300, 150, 1)
    outs = []
    for off in (None, 0):
        ud, idv, pd = (t.to(DEV).requires_grad_(True) for t in (u, it, pool))
        kw = {} if off is None else {"item_offset": 0}
        loss, lse, _ = ops.fused_inbatch_ce(ud, idv, ids.to(DEV), None, pd, 0.05, precision="bf16", **kw)
        loss.backward()
        outs.append((loss.detach(), lse, ud.grad, idv.grad, pd.grad))
    # forward: bitwise; backward: the two 128-column halves of a tile enter the Out accumulator in the order their
    # softmax groups finish (ce_tc.cu issue_out), so gradients repeat to fp32 accumulation-order noise only
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2:], outs[1][2:]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("Bg,W,D,H,n_ids", [(1024, 4, 128, 200, 300), (1536, 2, 64, 0, 10 ** 9), (2048, 8, 128, 0, 400),
                                            (777 * 3, 3, 128, 129, 500)])
def test_rank_slices_add_up_to_the_global_batch_loss_and_gradients(Bg, W, D, H, n_ids):
    from recommendsystemproject_b200 import ops
    T = 0.05
    u, it, pool, ids = _data(Bg, D, H, n_ids, Bg + W)
    B = Bg // W
    # fp64 oracle on the bf16-rounded operands, whole global batch
    lse_ref, pos_ref = O.row_logsumexp(_rnd(u), _rnd(it), ids, None, T, None if pool is None else _rnd(pool))
    loss_ref = float((lse_ref - pos_ref).mean())
    du_ref, di_ref, _, dp_ref = O.loss_grads_closed_form(_rnd(u), _rnd(it), ids, None, T, None if pool is None else _rnd(pool))
    losses, du, di, dp = [], [], torch.zeros(Bg, D, dtype=torch.float64), torch.zeros(max(H, 1), D, dtype=torch.float64)
    for r in range(W):
        ud = u[r * B:(r + 1) * B].to(DEV).requires_grad_(True)
        idv = it.to(DEV).requires_grad_(True)
        pd = pool.to(DEV).requires_grad_(True) if pool is not None else None
        loss, lse, _ = ops.fused_inbatch_ce(ud, idv, ids.to(DEV), None, pd, T, precision="bf16", item_offset=r * B)
        assert torch.allclose(lse.cpu().double(), lse_ref[r * B:(r + 1) * B], atol=2e-4)
        loss.backward()
        losses.append(float(loss))
        du.append(ud.grad.cpu().double() / W)
        di += idv.grad.cpu().double() / W
        if pd is not None:
            dp += pd.grad.cpu().double() / W
    assert abs(sum(losses) / W - loss_ref) < 2e-4
    du = torch.cat(du)
    for got, ref, what in ((du, du_ref, "dU"), (di, di_ref, "dI")) + (((dp, dp_ref, "dPool"),) if H else ()):
        err = float((got - ref.double()).norm() / ref.double().norm())
        assert err < 1e-2, (what, err)
        assert float((got - ref.double()).abs().max()) < 3e-2 * float(ref.abs().max()), what
    # the rectangular slices against the SQUARE kernel on the global batch: same bf16 operands, so the sum of the
    # slices is equal to fp32 accumulation-order noise
    ud, idv = u.to(DEV).requires_grad_(True), it.to(DEV).requires_grad_(True)
    pd = pool.to(DEV).requires_grad_(True) if pool is not None else None
    loss_sq, _, _ = ops.fused_inbatch_ce(ud, idv, ids.to(DEV), None, pd, T, precision="bf16")
    loss_sq.backward()
    assert abs(float(loss_sq) - sum(losses) / W) < 2e-5
    assert float((ud.grad.cpu().double() - du).norm() / du.norm()) < 1e-4
    assert float((idv.grad.cpu().double() - di).norm() / di.norm()) < 1e-4


def test_cross_rank_false_negatives_are_masked():
    """An item id that a user's positive shares with an item of ANOTHER rank must not count as a negative: give every
    item the same id and only the positive survives in each row => loss = 0 (KAT 2 of SURVEY section 4, global form)."""
    from recommendsystemproject_b200 import ops
    u, it, _, _ = _data(512, 128, 0, 10, 3)
    ids = torch.full((512,), 7)
    for r in range(4):
        loss, _, _ = ops.fused_inbatch_ce(u[r * 128:(r + 1) * 128].to(DEV), it.to(DEV), ids.to(DEV), None, None, 0.05,
                                          precision="bf16", item_offset=r * 128)
        assert abs(float(loss)) < 1e-6
    # without ids nothing is masked
    loss, _, _ = ops.fused_inbatch_ce(u[:128].to(DEV), it.to(DEV), None, None, None, 0.05, precision="bf16", item_offset=0)
    assert float(loss) > 1.0


def test_declared_id_range_gives_the_same_loss_and_flags_ids_outside():
    """id_bits (tt_ce_fwd_tc_rect_bits): the id sorts run over the declared bits only.  Same stable permutation => the
    forward is bitwise the undeclared call's; an id outside the declared range raises bit 3 of the flag word."""
    from recommendsystemproject_b200 import ops
    u, it, _, ids = _data(1024, 128, 0, 300, 11)
    assert ops.id_bits_for(300) == 9 and ops.id_bits_for(2) == 1 and ops.id_bits_for(10_000_001) == 24
    for off, rows in ((0, 1024), (256, 256)):                       # square form, and one rank's rectangular slab
        ud = u[off:off + rows].to(DEV)
        kw = {} if rows == 1024 else {"item_offset": off}
        ref, lse_ref, f0 = ops.fused_inbatch_ce(ud, it.to(DEV), ids.to(DEV), None, None, 0.05, precision="bf16", **kw)
        got, lse, f1 = ops.fused_inbatch_ce(ud, it.to(DEV), ids.to(DEV), None, None, 0.05, precision="bf16",
                                            id_bits=ops.id_bits_for(300), **kw)
        assert torch.equal(ref, got) and torch.equal(lse_ref, lse)
        assert int(f0) == 0 and int(f1) == 0
    bad = ids.clone()
    bad[17] = 1 << 20
    _, _, f2 = ops.fused_inbatch_ce(u.to(DEV), it.to(DEV), bad.to(DEV), None, None, 0.05, precision="bf16", id_bits=9)
    assert int(f2) & 8
