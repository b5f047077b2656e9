"""tcgen05 kind::tf32 Linear kernels (linear_tc.cu: forward, input gradient, weight + bias gradient) against an fp64
reference.  TF32 keeps 10 mantissa bits of each operand (fp32 accumulation), so the stated tolerance is a relative
Frobenius error of 2e-3 per result and 1e-2 of the largest entry elementwise; the exact fp32 kernels of linear_grad.cu
remain the parity path (tests/test_gpu_kernels.py)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

SHAPES = [(8192, 256, 640), (1000, 128, 256), (333, 128, 128), (6144, 256, 136), (129, 64, 48), (65, 8, 32),
          (4096, 384, 200), (20000, 128, 64)]


def _rel(a, b):
    return float((a.double() - b).norm() / b.norm())


@pytest.mark.parametrize("rows,n_out,n_in", SHAPES)
def test_tf32_linear_fwd_dgrad_wgrad_match_fp64(rows, n_out, n_in):
    from recommendsystemproject_b200 import _lib, ops
    lib = _lib.load()
    assert lib.tt_linear_tc_supported(rows, n_out, n_in)
    gen = torch.Generator(device=DEV).manual_seed(rows + n_out)
    x = torch.randn(rows, n_in, device=DEV, generator=gen)
    w = torch.randn(n_out, n_in, device=DEV, generator=gen) * 0.1
    b = torch.randn(n_out, device=DEV, generator=gen)
    g = torch.randn(rows, n_out, device=DEV, generator=gen)
    st = ops._stream()
    y = torch.empty(rows, n_out, device=DEV)
    ops.check(lib.tt_linear_fwd_tc(ops._p(x), ops._p(w), ops._p(b), rows, n_out, n_in, 0, ops._p(y), st), "fwd")
    yr = x.double() @ w.double().t() + b.double()
    assert _rel(y, yr) < 2e-3
    assert float((y.double() - yr).abs().max()) < 1e-2 * float(yr.abs().max())
    y2 = torch.empty_like(y)
    ops.check(lib.tt_linear_fwd_tc(ops._p(x), ops._p(w), None, rows, n_out, n_in, 1, ops._p(y2), st), "fwd relu")
    assert _rel(y2, torch.relu(x.double() @ w.double().t())) < 2e-3
    gx = torch.empty(rows, n_in, device=DEV)
    ops.check(lib.tt_linear_dgrad_tc(ops._p(g), ops._p(w), rows, n_out, n_in, ops._p(gx), st), "dgrad")
    assert _rel(gx, g.double() @ w.double()) < 2e-3
    nb = ctypes.c_size_t(0)
    ops.check(lib.tt_linear_wgrad_tc_workspace(rows, n_out, n_in, ctypes.byref(nb)), "ws")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=DEV)
    gw = torch.full((n_out, n_in), 7.0, device=DEV)
    gb = torch.full((n_out,), 7.0, device=DEV)
    ops.check(lib.tt_linear_wgrad_tc(ops._p(g), ops._p(x), rows, n_out, n_in, ops._p(gw), ops._p(gb), 0, ops._p(ws), ws.numel(), st), "wgrad")
    gwr = g.double().t() @ x.double()
    assert _rel(gw, gwr) < 2e-3
    assert torch.allclose(gb.double(), g.double().sum(0), atol=1e-3 * rows ** 0.5, rtol=1e-4)
    # accumulate = 1 adds into the buffers; bitwise repeatable
    gw2, gb2 = gw.clone(), gb.clone()
    ops.check(lib.tt_linear_wgrad_tc(ops._p(g), ops._p(x), rows, n_out, n_in, ops._p(gw2), ops._p(gb2), 1, ops._p(ws), ws.numel(), st), "wgrad acc")
    assert torch.allclose(gw2, 2 * gw, rtol=1e-6, atol=1e-6) and torch.allclose(gb2, 2 * gb, rtol=1e-6, atol=1e-5)
    gw3 = torch.empty_like(gw)
    ops.check(lib.tt_linear_wgrad_tc(ops._p(g), ops._p(x), rows, n_out, n_in, ops._p(gw3), None, 0, ops._p(ws), ws.numel(), st), "wgrad again")
    assert torch.equal(gw3, gw)


def test_linear_fn_uses_tf32_kernels_when_allowed_and_fp32_otherwise():
    from recommendsystemproject_b200 import ops
    x = torch.randn(512, 136, device=DEV, requires_grad=True)
    w = (torch.randn(256, 136, device=DEV) * 0.1).requires_grad_(True)
    b = torch.zeros(256, device=DEV, requires_grad=True)
    ref = x.detach().double() @ w.detach().double().t()
    y32 = ops.linear(x, w, b)
    assert _rel(y32.detach(), ref) < 1e-6
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        y = ops.linear(x, w, b)
        y.sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    e = _rel(y.detach(), ref)
    assert 1e-6 < e < 2e-3                          # TF32, not fp32
    assert _rel(x.grad, torch.ones(512, 256, dtype=torch.float64, device=DEV) @ w.detach().double()) < 2e-3
    assert _rel(w.grad, torch.ones(256, 512, dtype=torch.float64, device=DEV) @ x.detach().double()) < 2e-3
    assert torch.allclose(b.grad, torch.full((256,), 512.0, device=DEV))
