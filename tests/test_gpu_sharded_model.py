"""The integrated step with row-sharded tables at world 1 (the multi-rank version of the same check is
tools/dist_check.py, run under torchrun on 2+ GPUs): a model whose big tables go through the sharded exchange must
train exactly like the same model with plain nn.Embedding tables (the reference's construction, GenericTower.py:30-56)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mv(o):
    if isinstance(o, torch.Tensor):
        return o.to(DEV)
    if isinstance(o, dict):
        return {k: _mv(v) for k, v in o.items()}
    return [_mv(v) for v in o]


def _models(dim=64, dropout=0.0):
    import recommendsystemproject_b200 as tt
    from recommendsystemproject_b200 import synth
    cfg_s = synth.config_c3(v_user=40001, v_item=20001, dim=dim, dropout=dropout, shard=True, world=1)
    cfg_u = synth.config_c3(v_user=40001, v_item=20001, dim=dim, dropout=dropout, shard=False)
    torch.manual_seed(4321)
    ref = tt.TwoTowerModel(tt.GenericTower(cfg_u, "user_tower"), tt.GenericTower(cfg_u, "item_tower"), *synth.MAPS_C3).to(DEV).train()
    sh = tt.TwoTowerModel(tt.GenericTower(cfg_s, "user_tower"), tt.GenericTower(cfg_s, "item_tower"), *synth.MAPS_C3).to(DEV).train()
    sh.load_state_dict({k: v.clone() for k, v in ref.state_dict().items()})
    return ref, sh


def test_sharded_model_state_dict_keeps_reference_keys_and_shapes():
    ref, sh = _models()
    a, b = ref.state_dict(), sh.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k


def test_sharded_world1_step_equals_plain_embedding_step():
    import recommendsystemproject_b200 as tt
    from recommendsystemproject_b200 import dist as tdist, synth
    ref, sh = _models()
    batch = _mv(synth.make_batch_c3(B=512, L=30, v_user=40001, v_item=20001, seed=5))
    before = {k: v.clone() for k, v in ref.state_dict().items()}
    # eps = 1e-3 >> |g|: Adam's first step lr * g / (|g| + eps) stays LINEAR in g, so rounding-level gradient differences
    # stay rounding-level parameter differences (with the default 1e-8 an element with |g| ~ 1e-8 moves by ~lr/2 and a
    # 1e-9 change of g shows up as 5e-4 of parameter: not comparable between two summation orders)
    opt_s = tt.FusedTwoTowerOptimizer(sh, lr=1e-2, eps=1e-3, max_grad_norm=1.0, table_mode="sparse")
    step = tdist.ShardedTrainStep(sh, opt_s, batch, 0.05, loss_precision="fp32")
    loss_s = float(step())
    step.check_flags()
    opt_u = tt.FusedTwoTowerOptimizer(ref, lr=1e-2, eps=1e-3, max_grad_norm=1.0, table_mode="sparse")
    opt_u.zero_grad()
    u, i, _ = ref(batch)
    loss_u = ref.compute_loss(u, i, item_ids=batch["item_tower"]["sparse"][:, 0], temperature=0.05)
    loss_u.backward()
    opt_u.step()
    assert abs(loss_s - float(loss_u)) < 1e-5
    assert abs(float(opt_s.total_norm) - float(opt_u.total_norm)) < 1e-5 * float(opt_u.total_norm)
    new = sh.state_dict()
    moved = 0.0
    for k, v in ref.state_dict().items():
        if v.dtype.is_floating_point:
            assert torch.allclose(new[k], v, atol=2e-6, rtol=1e-5), (k, float((new[k] - v).abs().max()))
            moved = max(moved, float((v - before[k]).abs().max()))
    assert moved > 1e-4                 # the step did move the parameters
    # second step of both: same loss again
    loss_s2 = float(step())
    opt_u.zero_grad()
    u, i, _ = ref(batch)
    loss_u2 = float(ref.compute_loss(u, i, item_ids=batch["item_tower"]["sparse"][:, 0], temperature=0.05))
    assert abs(loss_s2 - loss_u2) < 1e-4


def test_fused_batchnorm_matches_torch_fwd_bwd_and_running_stats():
    """ops.batch_norm_act against torch's BatchNorm1d + ReLU (fp64 reference), plain and grouped ([B, G*C] view)."""
    from recommendsystemproject_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(3)
    for rows, C, G, relu in ((1000, 136, 1, False), (4096, 256, 1, True), (512, 48, 11, True), (77, 128, 1, True)):
        x = (torch.randn(rows, G * C, device=DEV, generator=gen) * 2 + 0.5).requires_grad_(True)
        gamma = (torch.rand(C, device=DEV, generator=gen) + 0.5).requires_grad_(True)
        beta = (torch.randn(C, device=DEV, generator=gen) * 0.1).requires_grad_(True)
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        nb = torch.zeros((), dtype=torch.int64, device=DEV)
        up = torch.randn(rows, G * C, device=DEV, generator=gen)
        y, mean, var_u = ops.batch_norm_act(x, gamma, beta, rm if G == 1 else None, rv if G == 1 else None,
                                            nb if G == 1 else None, 0.1, 1e-5, C, relu=relu)
        (y * up).sum().backward()
        xd = x.detach().double().requires_grad_(True)
        gd, bd = gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True)
        mu = xd.mean(0)
        var = xd.var(0, unbiased=False)
        yr = (xd - mu) / torch.sqrt(var + 1e-5) * gd.repeat(G) + bd.repeat(G)
        if relu:
            # the ReLU decision of an element within fp32 rounding of zero may differ between fp32 and fp64 (|y| ~ 1e-7,
            # but the whole upstream gradient flips): take the kernel's own mask for the reference
            yr = yr * (y.detach() > 0).double()
        (yr * up.double()).sum().backward()
        assert torch.allclose(y.double(), yr, atol=2e-5, rtol=1e-5)
        assert torch.allclose(x.grad.double(), xd.grad, atol=5e-5, rtol=1e-4)
        assert torch.allclose(gamma.grad.double(), gd.grad, atol=2e-3, rtol=1e-4)
        assert torch.allclose(beta.grad.double(), bd.grad, atol=2e-3, rtol=1e-4)
        assert torch.allclose(mean.double(), mu.detach(), atol=1e-5)
        assert torch.allclose(var_u.double(), xd.detach().var(0, unbiased=True), rtol=1e-4, atol=1e-6)
        if G == 1:
            assert torch.allclose(rm.double(), 0.1 * mu.detach(), atol=1e-5)
            assert torch.allclose(rv.double(), 0.9 + 0.1 * xd.detach().var(0, unbiased=True), rtol=1e-4)
            assert int(nb) == 1


def test_fused_batchnorm_dropout_is_unbiased_and_backward_uses_the_same_mask():
    from recommendsystemproject_b200 import ops
    rows, C, p = 8192, 64, 0.3
    x = torch.randn(rows, C, device=DEV).requires_grad_(True)
    gamma, beta = torch.ones(C, device=DEV, requires_grad=True), torch.full((C,), 2.0, device=DEV, requires_grad=True)
    seed = torch.tensor([1234], dtype=torch.int64, device=DEV)
    y, _, _ = ops.batch_norm_act(x, gamma, beta, None, None, None, 0.1, 1e-5, C, relu=True, dropout_p=p, seed_dev=seed, call_id=3)
    kept = y != 0
    assert abs(float(kept.float().mean()) - (1 - p) * float((torch.randn(100000) + 2 > 0).float().mean())) < 0.02
    y2, _, _ = ops.batch_norm_act(x, gamma, beta, None, None, None, 0.1, 1e-5, C, relu=True, dropout_p=p, seed_dev=seed, call_id=3)
    assert torch.equal(y, y2)                      # same (seed, call id) => same mask
    y.sum().backward()
    # where the output was dropped the gradient wrt beta's contribution vanishes: d sum(y) / d beta_c = kept count / (1-p)
    assert torch.allclose(beta.grad, kept.float().sum(0) / (1 - p), rtol=1e-4)
