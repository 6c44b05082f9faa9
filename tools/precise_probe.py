#!/usr/bin/env python
"""Probe: error of the 3-pass ("precise") linear kernels against float64 at the fMRI-net shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_eeg_fmri_b200 import ops


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


torch.manual_seed(0)
for (M, N, K) in [(2048, 128, 40000), (2048, 128, 400), (2048, 64, 128), (4096, 128, 128)]:
    x = torch.randn(M, K, device="cuda") + 0.3
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    dy = torch.randn(M, N, device="cuda")
    y = ops.linear_fwd_precise(x, w, None)
    y1 = ops.linear_fwd(ops.round_tf32(x), ops.round_tf32(w), None)
    yr = x.double() @ w.double().T
    y32 = x @ w.T
    d = (y.double() - yr)
    print(f"M{M} N{N} K{K} fwd precise {rel(y, yr):.2e} (bias {float((d*yr.sign()).mean()/yr.abs().mean()):+.2e}) single {rel(y1, yr):.2e} torch-fp32 {rel(y32, yr):.2e}")
    dx = ops.linear_dgrad_precise(dy, w)
    print(f"   dgrad precise {rel(dx, dy.double() @ w.double()):.2e}")
    dw, db = ops.linear_wgrad_precise(dy, x)
    print(f"   wgrad precise {rel(dw, dy.double().T @ x.double()):.2e}  db {rel(db, dy.double().sum(0)):.2e}")
# BatchNorm forward/backward at the same shapes
import torch.nn.functional as F
from multimodal_eeg_fmri_b200 import functional as XF
for (M, C) in [(2048, 128), (2048, 64)]:
    x = (torch.randn(M, C, device="cuda") * 3 + 1).requires_grad_(True)
    g = torch.randn(M, C, device="cuda") + 0.5
    gam = torch.rand(C, device="cuda") + 0.5
    bet = torch.randn(C, device="cuda")
    xd = x.detach().double().requires_grad_(True)
    yd = torch.relu(F.batch_norm(xd, None, None, gam.double(), bet.double(), True, 0.1, 1e-5))
    (gxd,) = torch.autograd.grad(yd, xd, g.double())
    print("bn available fns:", [n for n in dir(ops) if n.startswith("bn_")])
    break
