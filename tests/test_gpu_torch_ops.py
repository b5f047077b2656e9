"""torch.ops.tt_b200.* (recommendsystemproject_b200/torch_ops.py): the dispatcher-visible ops give the same numbers as
the module-level op layer and carry working autograd registrations."""
import pytest
import torch

from oracle import twotower_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_gather_pool_op_forward_backward_against_oracle():
    import recommendsystemproject_b200  # noqa: F401
    from recommendsystemproject_b200 import ops
    gen = torch.Generator().manual_seed(0)
    w = torch.randn(300, 64, generator=gen)
    ids = torch.randint(0, 300, (50, 9), generator=gen)
    up = torch.randn(50, 64, generator=gen)
    wd = w.to(DEV).requires_grad_(True)
    out = torch.ops.tt_b200.gather_pool(wd, ids.to(DEV), ops.POOL_MEAN, 0)
    assert torch.allclose(out.cpu(), O.pooled_lookup(w, ids, "mean"), atol=1e-6)
    (out * up.to(DEV)).sum().backward()
    ref = O.embedding_grad_dense(ids, (up / 9).unsqueeze(1).expand(50, 9, 64).reshape(-1, 64), 300, 0)
    assert torch.allclose(wd.grad.cpu(), ref, atol=1e-5)
    torch.library.opcheck(torch.ops.tt_b200.gather_pool.default, (w.to(DEV), ids.to(DEV), ops.POOL_SUM, 0),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


@pytest.mark.parametrize("precision", [0, 1])
def test_inbatch_ce_op_matches_the_op_layer(precision):
    import recommendsystemproject_b200  # noqa: F401
    from recommendsystemproject_b200 import ops
    gen = torch.Generator().manual_seed(1)
    u = torch.nn.functional.normalize(torch.randn(512, 128, generator=gen), dim=1)
    i = torch.nn.functional.normalize(torch.randn(512, 128, generator=gen), dim=1)
    ids = torch.randint(1, 100, (512,), generator=gen)
    a_u, a_i = u.to(DEV).requires_grad_(True), i.to(DEV).requires_grad_(True)
    loss, lse = torch.ops.tt_b200.inbatch_ce(a_u, a_i, ids.to(DEV), 0.1, precision)
    loss.backward()
    b_u, b_i = u.to(DEV).requires_grad_(True), i.to(DEV).requires_grad_(True)
    ref, ref_lse, _ = ops.fused_inbatch_ce(b_u, b_i, ids.to(DEV), None, None, 0.1, precision="bf16" if precision else "fp32")
    ref.backward()
    assert torch.equal(loss.detach(), ref.detach()) and torch.equal(lse, ref_lse)
    assert torch.allclose(a_u.grad, b_u.grad, rtol=1e-4, atol=1e-8) and torch.allclose(a_i.grad, b_i.grad, rtol=1e-4, atol=1e-8)
    if precision == 0:
        assert abs(float(loss.detach()) - float(O.compute_loss(u, i, ids, None, 0.1))) < 5e-6


def test_score_topk_op_rows_are_the_oracles():
    import numpy as np
    import recommendsystemproject_b200  # noqa: F401
    gen = torch.Generator().manual_seed(2)
    q = torch.nn.functional.normalize(torch.randn(64, 128, generator=gen), dim=1)
    e = torch.nn.functional.normalize(torch.randn(9000, 128, generator=gen), dim=1)
    _, ref = O.score_topk(q.numpy(), e.numpy(), 20)
    for precision in (0, 1):
        _, rows = torch.ops.tt_b200.score_topk(q.to(DEV), e.to(DEV), 20, precision)
        assert np.array_equal(rows.cpu().numpy(), ref)
