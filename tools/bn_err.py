import torch, sys
sys.path.insert(0, ".")
from recommendsystemproject_b200 import ops
DEV="cuda"
gen = torch.Generator(device=DEV).manual_seed(3)
for rows, C, G, relu in ((1000, 136, 1, False), (4096, 256, 1, True), (512, 48, 11, True), (77, 128, 1, True), (65536, 640, 1, True)):
    x = (torch.randn(rows, G * C, device=DEV, generator=gen) * 2 + 0.5).requires_grad_(True)
    gamma = (torch.rand(C, device=DEV, generator=gen) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=DEV, generator=gen) * 0.1).requires_grad_(True)
    up = torch.randn(rows, G * C, device=DEV, generator=gen)
    y, mean, var_u = ops.batch_norm_act(x, gamma, beta, None, None, None, 0.1, 1e-5, C, relu=relu)
    (y * up).sum().backward()
    xd = x.detach().double().requires_grad_(True)
    gd, bd = gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True)
    mu = xd.mean(0); var = xd.var(0, unbiased=False)
    yr = (xd - mu) / torch.sqrt(var + 1e-5) * gd.repeat(G) + bd.repeat(G)
    if relu: yr = torch.relu(yr)
    (yr * up.double()).sum().backward()
    print(rows, C, G, relu, "y", float((y.double()-yr).abs().max()), "dx", float((x.grad.double()-xd.grad).abs().max()), "dx scale", float(xd.grad.abs().max()),
          "dgamma", float((gamma.grad.double()-gd.grad).abs().max()), float(gd.grad.abs().max()), "dbeta", float((beta.grad.double()-bd.grad).abs().max()))
