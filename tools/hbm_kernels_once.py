#!/usr/bin/env python
"""Runs each HBM-bound kernel a few times at one large size (for ncu captures): rays x samples from argv."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ddnerf_oracle as orc
from ddnerf_b200 import mlp_tc, ops
from ddnerf_b200.rays import synth_rays
N, S = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
ro, rd, rad, near, far = synth_rays("blender", N, seed=2)
rays = orc.pack_rays(ro, rd, rad, near, far).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
u = torch.rand(N, S + 1, device=dev, generator=g)
t0 = ops.sample_first_cycle(rays[:, 7:8], rays[:, 8:9], S, False, u)
raw = torch.randn(N, S, 4, device=dev, generator=g)
noise = torch.randn(N, S, device=dev, generator=g)
mus = torch.rand(N, S, device=dev, generator=g)
sig = torch.rand(N, S, device=dev, generator=g) * 0.5 + 1e-3
lt = 0.5 * (1 + torch.erf((0 - mus) / sig / 2 ** 0.5))
pin = 0.5 * (1 + torch.erf((1 - mus) / sig / 2 ** 0.5)) - lt
for _ in range(3):
    t0 = ops.sample_first_cycle(rays[:, 7:8], rays[:, 8:9], S, False, u)
    mlp_tc.encode_img(rays, t0)
    rawg = raw.clone().requires_grad_(True)
    out = ops.composite(rawg, t0, rays[:, 3:6], noise, 1.0, None, False, True, False)
    w = out[3].detach()
    torch.autograd.grad((out[0], out[3]), rawg, (torch.ones_like(out[0]), torch.ones_like(out[3])))
    t1 = ops.sample_pdf(t0, w, S + 1, True, u)
    t2 = ops.sample_pdf_mu_sigma(t0, w, mus, sig, pin, lt, S + 1, True, near, far, u)
    w0g, mug, sgg = w.clone().requires_grad_(True), mus.clone().requires_grad_(True), sig.clone().requires_grad_(True)
    l = ops.dp_loss(t2, t0, w, w0g, mug, sgg, lt, pin, False)
    torch.autograd.grad(l, (w0g, mug, sgg))
    # round 2: the fused DDNeRF coarse glue (compositor + (mu, sigma) head + regulariser sums), forward and backward
    raw6 = (torch.randn(N, S, 6, device=dev, generator=g) * 0.5).requires_grad_(True)
    dd = ops.composite_dd(raw6, t0, rays[:, 3:6], noise, 1.0, False, True, 0.01)
    torch.autograd.grad((dd[0], dd[3], dd[4], dd[6], dd[7], dd[8]), raw6,
                        tuple(torch.ones_like(x) for x in (dd[0], dd[3], dd[4], dd[6], dd[7], dd[8])))
torch.cuda.synchronize()
print("ok")
